"""CPU oracle: fp64 NumPy/SciPy restatement of the reference's two hot paths.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / CPU baseline.  The
product path (``juliaraytracingsw_b200`` + ``libswrt.so``) never imports it and
fails loudly when the CUDA library is missing.

What it restates (all citations relative to the reference checkout):

* ``grid``      FourierFlows ``TwoDGrid`` conventions, ``dealias!``, ``makefilter``,
                ``parsevalsum2`` (third-party, un-vendored: recalled, see SURVEY App. A.1/C;
                pinned through KATs K3, K4, K5 which exercise the grid arrays).
* ``rsw``       ``rsw/RotatingShallowWater.jl`` (``calcN!`` :140-230, ``populate_L!``
                :262-274, ``updatevars!`` :101-116, energies :323-336) and the
                Modified / Lindborg variants.
* ``ifmab3``    ``utils/IFMAB3.jl`` (``getexpLs`` :26-30, ``IFMAB3update!`` :129-140,
                ``stepforward!`` :157-169).
* ``raytrace``  ``raytracing/GPURaytracing.jl`` :18-65 (ray RHS, texture-coordinate
                bilinear sampling), ``raytracing/RaytracingDriver.jl`` :27-47,132-154,
                ``rsw/RSWRaytracingDriver.jl`` :15-67, fixed-step RK4 per the north star.
* ``outputs``   ``utils/SequencedOutputs.jl`` :37-63 and ``utils/Collated.jl`` :40-60
                roll-over arithmetic.

Parity status: the reference itself cannot run here (pure Julia on FourierFlows /
OrdinaryDiffEq, neither Julia nor the packages are installed, no network).  The
oracle is pinned against the known-answer values the reference's notebooks record
(K1, K3, K4, K5, K6, K7, K12, K13 in SURVEY.md section 4; see ``tests/test_oracle_kats.py``).
Third-party behaviour that no recorded value pins -- ``makefilter``, the
FilteredAB3/ETDRK4 steppers, and anything to do with OrdinaryDiffEq's adaptive
Vern7 -- is "parity unpinned" and says so where it is implemented.  The packet
integrator is the north star's fixed-step RK4, not the reference's Vern7.
"""
