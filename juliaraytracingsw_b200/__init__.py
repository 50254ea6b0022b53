"""swrt: B200-native pseudo-spectral flow step + wave-packet ray tracer behind the reference's API."""
from . import flow, outputs, raytracing  # noqa: F401
from ._lib import LIB_PATH, SwrtError, lib  # noqa: F401
from .flow import (Problem, enforce_reality_condition, kinetic_energy, potential_energy, set_solution,  # noqa: F401
                   stepforward, updatevars)
from .raytracing import (Packets, Velocity, VelocityGradient, create_template_ode, generate_initial_wavepackets,  # noqa: F401
                         get_velocity_info, interpolate_gradients, interpolate_velocity, raytrace)
