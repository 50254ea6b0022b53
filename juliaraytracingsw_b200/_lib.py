"""ctypes binding of libswrt.so (the C ABI in include/swrt.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWRT_LIB", os.path.join(_HERE, "lib", "libswrt.so"))   # SWRT_LIB: A/B-test another build


class SwrtError(RuntimeError):
    pass


class FlowDesc(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("stepper", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nnu", C.c_int),
        ("use_filter", C.c_int), ("filter_order", C.c_int), ("device", C.c_int),
        ("Lx", C.c_double), ("Ly", C.c_double), ("dt", C.c_double), ("nu", C.c_double), ("f", C.c_double),
        ("Cg", C.c_double), ("aliased_fraction", C.c_double), ("filter_innerK", C.c_double),
        ("filter_outerK", C.c_double), ("filter_tol", C.c_double), ("U", C.c_double), ("mu", C.c_double),
        ("F", C.c_double), ("Ro", C.c_double), ("Kd2", C.c_double), ("U2", C.c_double), ("beta", C.c_double), ("slab_rank", C.c_int), ("slab_size", C.c_int),
    ]


class PacketsDesc(C.Structure):
    _fields_ = [("n", C.c_longlong), ("interp", C.c_int), ("nsub", C.c_int), ("time_lerp", C.c_int),
                ("sort_every", C.c_int), ("integrator", C.c_int), ("f", C.c_double), ("Cg", C.c_double),
                ("band_first", C.c_longlong), ("band_capacity", C.c_longlong)]


class SeqOut(C.Structure):
    _fields_ = [("max_writes", C.c_longlong), ("current_writes", C.c_longlong), ("file_index", C.c_longlong)]


# name -> (restype, argtypes); every symbol include/swrt.h declares
_P, _I, _LL, _D = C.c_void_p, C.c_int, C.c_longlong, C.c_double
_PD, _PLL, _PI, _PF = C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_float)
SIGNATURES = {
    "swrt_last_error": (C.c_char_p, []),
    "swrt_version": (_I, []),
    "swrt_device_count": (_I, [_PI]),
    "swrt_flow_create": (_I, [C.POINTER(FlowDesc), C.POINTER(_P)]),
    "swrt_flow_destroy": (_I, [_P]),
    "swrt_flow_set_solution": (_I, [_P, _P]),
    "swrt_flow_get_solution": (_I, [_P, _P]),
    "swrt_flow_set_forcing": (_I, [_P, _P]),
    "swrt_flow_set_rsw_initial_condition": (_I, [_P, _P, _P, _D, _D, _D, _D, _D, _D, _PD]),
    "swrt_flow_enforce_reality": (_I, [_P]),
    "swrt_flow_step": (_I, [_P, _I]),
    "swrt_flow_clock": (_I, [_P, _PD, _PLL]),
    "swrt_flow_set_clock": (_I, [_P, _D, _LL]),
    "swrt_flow_get_field": (_I, [_P, _I, _P]),
    "swrt_flow_set_field_physical": (_I, [_P, _I, _P]),
    "swrt_flow_energies": (_I, [_P, _PD, _PD]),
    "swrt_flow_layer_kinetic_energy": (_I, [_P, _I, _PD]),
    "swrt_flow_wave_balanced_decomposition": (_I, [_P, _P, _P]),
    "swrt_flow_wave_balanced_weights": (_I, [_P, _P, _P, _P]),
    "swrt_flow_wave_balanced_energies": (_I, [_P, _PD]),
    "swrt_flow_barotropic_energy": (_I, [_P, _PD]),
    "swrt_flow_max_abs_uv": (_I, [_P, _PD, _PD]),
    "swrt_flow_has_nan": (_I, [_P, _PI]),
    "swrt_flow_velocity_snapshot": (_I, [_P, _I, _I]),
    "swrt_flow_swap_snapshots": (_I, [_P, _I]),
    "swrt_flow_set_interp": (_I, [_P, _I]),
    "swrt_flow_snapshot_fields": (_I, [_P, _PI]),
    "swrt_flow_set_nufft_width": (_I, [_P, _I]),
    "swrt_flow_set_snapshot_refinement": (_I, [_P, _I]),
    "swrt_flow_snapshot_dims": (_I, [_P, _PI, _PI]),
    "swrt_flow_get_snapshot": (_I, [_P, _I, _P]),
    "swrt_flow_set_snapshot": (_I, [_P, _I, _P]),
    "swrt_flow_set_stream": (_I, [_P, _P]),
    "swrt_slab_buffer": (_I, [_P, _I, C.POINTER(_P), _PLL]),
    "swrt_slab_info": (_I, [_P, _PI, _PI, _PI, _PI]),
    "swrt_slab_ipc_handle": (_I, [_P, _I, _P]),
    "swrt_slab_ipc_open": (_I, [_P, _I, _I, _P]),
    "swrt_slab_set_mode": (_I, [_P, _I]),
    "swrt_slab_p2p": (_I, [_P, _PI]),
    "swrt_slab_stage_a": (_I, [_P]),
    "swrt_slab_stage_b": (_I, [_P]),
    "swrt_slab_stage_c": (_I, [_P]),
    "swrt_slab_psi_a": (_I, [_P, _I]),
    "swrt_slab_snap_b": (_I, [_P, _I]),
    "swrt_slab_set_barrier": (_I, [_P, _I, _P, _P]),
    "swrt_slab_barrier": (_I, [_P]),
    "swrt_slab_step": (_I, [_P, _I]),
    "swrt_slab_band_snapshot": (_I, [_P, _I, _I]),
    "swrt_slab_band_info": (_I, [_P, _PI, _PI, _PI]),
    "swrt_packets_ipc_handle": (_I, [_P, _P]),
    "swrt_packets_ipc_open": (_I, [_P, _I, _P]),
    "swrt_packets_resident": (_I, [_P, _PLL]),
    "swrt_flow_timer_start": (_I, [_P]),
    "swrt_flow_timer_stop": (_I, [_P, _PF]),
    "swrt_flow_sync": (_I, [_P]),
    "swrt_flow_launch_count": (_I, [_P, _PLL]),
    "swrt_flow_profile": (_I, [_P, _I]),
    "swrt_flow_profile_get": (_I, [_P, _I, _PD, _PLL, C.POINTER(C.c_char_p)]),
    "swrt_packets_create": (_I, [C.POINTER(PacketsDesc), _P, C.POINTER(_P)]),
    "swrt_packets_destroy": (_I, [_P]),
    "swrt_packets_set": (_I, [_P, _P, _P]),
    "swrt_packets_get": (_I, [_P, _P]),
    "swrt_packets_generate": (_I, [_P, _D, _D, _LL, _LL]),
    "swrt_packets_raytrace": (_I, [_P, _D, _D]),
    "swrt_packets_set_kernel": (_I, [_P, _I]),
    "swrt_packets_sample": (_I, [_P, _I, _P, _P]),
    "swrt_packets_kcutoff_reset": (_I, [_P, _D, _D, _PLL]),
    "swrt_series_create": (_I, [_P, _I, _I, _LL, C.POINTER(_P)]),
    "swrt_series_destroy": (_I, [_P]),
    "swrt_series_append": (_I, [_P]),
    "swrt_series_frames": (_I, [_P, _PLL]),
    "swrt_series_times": (_I, [_P, _P]),
    "swrt_series_get": (_I, [_P, _I, _P]),
    "swrt_series_spectrum": (_I, [_P, _I, _P]),
    "swrt_packets_coupled_steps": (_I, [_P, _I, _I, _D, _D]),
    "swrt_packets_use_own_stream": (_I, [_P]),
    "swrt_packets_set_async": (_I, [_P, _P, _LL, _P]),
    "swrt_packets_get_async": (_I, [_P, _P, _LL]),
    "swrt_packets_sample_async": (_I, [_P, _I, _P, _P, _LL]),
    "swrt_packets_sync": (_I, [_P]),
    "swrt_seqout_init": (_I, [C.POINTER(SeqOut), _LL]),
    "swrt_seqout_write": (_I, [C.POINTER(SeqOut), _LL, _PLL]),
    "swrt_seqout_filename": (_I, [C.c_char_p, _LL, C.c_char_p, _I]),
    "swrt_collated_filename": (_I, [C.c_char_p, _LL, C.c_char_p, _I]),
}

_lib = None


def lib():
    """Load libswrt.so once; raises if it has not been built (`python -c "import __graft_entry__ as g; g.build()"`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SwrtError(f"{LIB_PATH} not found: build it with juliaraytracingsw_b200/csrc/Makefile "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SwrtError(f"libswrt error {rc}: {lib().swrt_last_error().decode()}")
