"""Team mode: the flow slab-decomposed over P GPUs, the packets sharded by y-band, one process per GPU.

Rank r owns a block of retained kr columns in spectral space and rows [r ny/P, (r+1) ny/P) in physical space.  The two
transposes inside every 2-D transform are the stores of the FFT passes themselves, straight into the peers' receive buffers
over NVLink (CUDA IPC), with a device-side barrier where an all-to-all would be; the velocity snapshot for the packets is
produced band-wise (each rank gets exactly the rows its packets can touch, plus a few halo rows pulled from the two neighbours),
because the packets are sharded by y-band and handed over between ranks at every re-sort.  All of that is native
(`swrt_slab_step`, `swrt_slab_band_snapshot`, `swrt_packets_*` in band mode): this module only creates the handles and
distributes the 64-byte IPC handles once through `torch.distributed` (any backend; a Julia host would use MPI.jl or files).

`p2p=False` keeps the round-1 variant of the flow step with `torch.distributed.all_to_all_single` (NCCL) between the phases.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import flow
from ._lib import check, lib

A_SEND, A_RECV, B_SEND, B_RECV, SNAP0, SNAP1, FLAGS, BAND = range(8)
_BARRIER_CB = C.CFUNCTYPE(None, C.c_void_p)


class _DevBuf:
    """Zero-copy view of a libswrt device buffer for torch (CUDA array interface v3)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


class SlabProblem(flow.Problem):
    """`Problem` whose step is distributed over the ranks of `dist` (a torch.distributed process group).
    Every model and stepper (the multi-stage steppers run the three slab passes once per stage: swrt_slab_step needs the mapped peers);
    packets on a slab-decomposed flow use the fp64 bilinear interpolant."""

    def __init__(self, dist, dev=0, p2p=True, pull=None, barrier=None, **kw):
        """p2p=True: native team mode (peer stores + device barrier).  barrier = "device" (flag words over NVLink, default) or
        "host" (stream synchronise + dist.barrier(): for ranks that share one GPU, e.g. the single-GPU parity tests).
        p2p=False: NCCL all_to_all_single between the phases, flow step only."""
        import torch
        self.dist, self.torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        super().__init__(dev, slab=(self.rank, self.world), **kw)
        yr, ch, na, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().swrt_slab_info(self._h, C.byref(yr), C.byref(ch), C.byref(na), C.byref(nb)))
        self.yrows, self.chunk, self.njobs_a, self.njobs_b = yr.value, ch.value, na.value, nb.value
        self.kr_lo = self.rank * self.chunk
        self.p2p = False
        # first transpose: "push" (y-pass stores into the peers), "pull" (x-pass reads the peers' send buffers) or "copy"
        # (local stores + a block-copy kernel); measured in profiles/r01_g_multigpu_summary.md
        # (r02_h, 8 GPUs, 2048^2: push 0.331 / copy 0.341 ms per coupled step -- 64-byte pieces from 4-column y tiles; at 4096^2
        # the tiles are 2 columns wide and the block copy wins from 4 ranks on)
        default = "copy" if (self.world >= 4 and kw.get("nx", 128) > 2048) else "push"
        self.mode = os.environ.get("SWRT_SLAB_MODE", default) if pull is None else ("pull" if pull else "push")
        self.pull = self.mode == "pull"
        self.barrier = os.environ.get("SWRT_TEAM_BARRIER", "device") if barrier is None else barrier
        if p2p:
            self._open_peers()
        else:
            check(lib().swrt_flow_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            self._buf = {}
            for which in (A_SEND, A_RECV, B_SEND, B_RECV):
                p, n = C.c_void_p(), C.c_longlong()
                check(lib().swrt_slab_buffer(self._h, which, C.byref(p), C.byref(n)))
                self._buf[which] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{dev}")

    # ---------------------------------------------------------------- setup: one exchange of IPC handles
    def _open_peers(self):
        L = lib()
        shared = (A_RECV, B_RECV, FLAGS, BAND) + ((A_SEND,) if self.pull else ())
        mine = {}
        for which in shared:
            buf = C.create_string_buffer(64)
            check(L.swrt_slab_ipc_handle(self._h, which, buf))
            mine[which] = buf.raw
        allh = [None] * self.world
        self.dist.all_gather_object(allh, mine)
        for r, hs in enumerate(allh):
            for which in shared:
                check(L.swrt_slab_ipc_open(self._h, which, r, hs[which]))
        en = C.c_int()
        check(L.swrt_slab_p2p(self._h, C.byref(en)))
        self.p2p = bool(en.value)
        assert self.p2p
        check(L.swrt_slab_set_mode(self._h, {"push": 0, "pull": 1, "copy": 2}[self.mode] + (16 if int(os.environ.get("SWRT_SLAB_B_COPY", "0")) else 0)))
        if self.barrier == "host":
            self._cb = _BARRIER_CB(lambda _arg: self.dist.barrier())          # keep the callback object alive
            check(L.swrt_slab_set_barrier(self._h, 1, C.cast(self._cb, C.c_void_p), None))
        self.dist.barrier()

    def team_barrier(self):
        check(lib().swrt_slab_barrier(self._h))

    # ---------------------------------------------------------------- the step
    def _a2a(self, recv, send, njobs):
        n = self.world * njobs * self.yrows * self.chunk * 2        # doubles: [dest][job][row][chunk] complex128
        self.dist.all_to_all_single(self._buf[recv][:n], self._buf[send][:n])

    def stepforward(self, nsteps=1):
        L = lib()
        if self.p2p:
            check(L.swrt_slab_step(self._h, int(nsteps)))
            return
        for _ in range(int(nsteps)):
            check(L.swrt_slab_stage_a(self._h))
            self._a2a(A_RECV, A_SEND, self.njobs_a)
            check(L.swrt_slab_stage_b(self._h))
            self._a2a(B_RECV, B_SEND, self.njobs_b)
            check(L.swrt_slab_stage_c(self._h))

    def velocity_snapshot(self, slot, psi_kind):
        """get_streamfunction! + get_velocity_info into snapshot `slot`: this rank's band of rows (+ halo)."""
        if not self.p2p:
            raise RuntimeError("band snapshots need the native team mode (p2p=True)")
        check(lib().swrt_slab_band_snapshot(self._h, int(psi_kind), int(slot)))

    # ---------------------------------------------------------------- host-side assembly (tests, output frames)
    def _gather(self, local):
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local)
        return parts

    def gather_snapshot(self, slot):
        """The full (nx, ny, 5) snapshot assembled from the bands of all ranks (every rank gets it)."""
        band = np.empty((self.grid.nx, self.yrows, 5), dtype=np.float64, order="F")
        check(lib().swrt_flow_get_snapshot(self._h, int(slot), band.ctypes.data_as(C.c_void_p)))
        return np.concatenate(self._gather(band), axis=1)

    def set_snapshot(self, slot, fields):
        """Load this rank's rows of a full (nx, ny, 5) field array into `slot` (collective: halo rows come from the neighbours)."""
        a = np.asfortranarray(np.asarray(fields, dtype=np.float64)[:, self.rank * self.yrows:(self.rank + 1) * self.yrows, :])
        check(lib().swrt_flow_set_snapshot(self._h, int(slot), a.ctypes.data_as(C.c_void_p)))

    def gather_solution(self):
        """Array(prob.sol) assembled from the column slabs of all ranks (every rank gets the full array)."""
        out = np.zeros(self._sol_shape(), dtype=np.complex128, order="F")
        for part in self._gather(self._local_sol()):
            out += part                                             # slabs are disjoint, the rest is zero
        return out[:, :, 0] if self.desc.model == 4 else out

    def _sol_shape(self):
        return (self.grid.nkr, self.grid.nl, self.nvar)

    def _local_sol(self):
        out = np.empty(self._sol_shape(), dtype=np.complex128, order="F")
        check(lib().swrt_flow_get_solution(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def energies(self):
        """(ke, pe) summed over the slabs."""
        ke, pe = C.c_double(), C.c_double()
        check(lib().swrt_flow_energies(self._h, C.byref(ke), C.byref(pe)))
        parts = self._gather((ke.value, pe.value))
        return float(sum(p[0] for p in parts)), float(sum(p[1] for p in parts))
