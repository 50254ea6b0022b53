"""Slab-decomposed flow step over P GPUs (the 4096^2 configurations), one process per GPU.

Rank r owns a block of retained kr columns in spectral space and ny/P rows in physical space.  The two transposes
inside every 2-D transform are `torch.distributed.all_to_all_single` calls (NCCL over NVLink) on exchange buffers that
live inside libswrt (wrapped zero-copy through `__cuda_array_interface__`); everything else is the same CUDA kernels as
the single-GPU path, launched on torch's current stream so that kernels and collectives are ordered on the device.
A velocity snapshot for the packets ends with an all-gather: every rank holds the whole background field because its
packets may sit anywhere in the domain.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import flow
from ._lib import check, lib

A_SEND, A_RECV, B_SEND, B_RECV, SNAP0, SNAP1 = range(6)


class _DevBuf:
    """Zero-copy view of a libswrt device buffer for torch (CUDA array interface v3)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


class SlabProblem(flow.Problem):
    """`Problem` whose step is distributed over the ranks of `dist` (a torch.distributed process group).
    Supported: RotatingShallowWater, SWQG, TwoLayerQG with the IFMAB3 stepper."""

    def __init__(self, dist, dev=0, p2p=True, pull=None, **kw):
        """p2p=True: the transposes are direct NVLink stores into the peers' receive buffers (CUDA IPC) followed by a
        stream-ordered barrier; p2p=False: NCCL all_to_all_single on send/receive buffers."""
        import torch
        self.dist, self.torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        super().__init__(dev, slab=(self.rank, self.world), **kw)
        check(lib().swrt_flow_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        yr, ch, na, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().swrt_slab_info(self._h, C.byref(yr), C.byref(ch), C.byref(na), C.byref(nb)))
        self.yrows, self.chunk, self.njobs_a, self.njobs_b = yr.value, ch.value, na.value, nb.value
        self._buf = {}
        for which in (A_SEND, A_RECV, B_SEND, B_RECV, SNAP0, SNAP1):
            p, n = C.c_void_p(), C.c_longlong()
            check(lib().swrt_slab_buffer(self._h, which, C.byref(p), C.byref(n)))
            self._buf[which] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{dev}")
        self.kr_lo = self.rank * self.chunk
        self._flag = torch.zeros(1, device=f"cuda:{dev}")
        self.p2p = False
        # first transpose: "push" (y-pass stores into the peers), "pull" (x-pass reads the peers' send buffers) or "copy"
        # (local stores + a block-copy kernel); measured in profiles/r01_g_multigpu_summary.md
        default = "copy" if self.world >= 4 else "push"             # 8 GPUs: 0.489 (copy) / 0.502 (push) / 0.559 (pull) ms per step
        self.mode = os.environ.get("SWRT_SLAB_MODE", default) if pull is None else ("pull" if pull else "push")
        self.pull = self.mode == "pull"
        if p2p:
            self._open_peers()

    def _open_peers(self):
        """Exchange the CUDA IPC handles of the two receive buffers and map every peer's."""
        L = lib()
        mine = {}
        shared = (A_RECV, B_RECV, A_SEND) if self.pull else (A_RECV, B_RECV)
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
        if self.p2p:
            check(L.swrt_slab_set_mode(self._h, {"push": 0, "pull": 1, "copy": 2}[self.mode]))
        self._barrier()

    def _barrier(self):
        """Stream-ordered barrier: every rank's earlier kernels (and their peer stores) are complete before anything after it runs."""
        self.dist.all_reduce(self._flag)

    def _a2a(self, recv, send, njobs):
        if self.p2p:                                                # the pass already stored into the peers' buffers
            self._barrier()
            return
        n = self.world * njobs * self.yrows * self.chunk * 2        # doubles: [dest][job][row][chunk] complex128
        self.dist.all_to_all_single(self._buf[recv][:n], self._buf[send][:n])

    def stepforward(self, nsteps=1):
        L = lib()
        for _ in range(int(nsteps)):
            check(L.swrt_slab_stage_a(self._h))
            self._a2a(A_RECV, A_SEND, self.njobs_a)
            check(L.swrt_slab_stage_b(self._h))
            self._a2a(B_RECV, B_SEND, self.njobs_b)
            check(L.swrt_slab_stage_c(self._h))

    def velocity_snapshot(self, slot, psi_kind):
        """get_streamfunction! + get_velocity_info into snapshot `slot` on every rank."""
        L = lib()
        check(L.swrt_slab_psi_a(self._h, int(psi_kind)))
        self._a2a(A_RECV, A_SEND, 3)
        check(L.swrt_slab_snap_b(self._h, int(slot)))
        # the buffer pointer of a slot follows swap_snapshots(): look it up every time
        p, n = C.c_void_p(), C.c_longlong()
        check(L.swrt_slab_buffer(self._h, SNAP0 + int(slot), C.byref(p), C.byref(n)))
        full = self.torch.as_tensor(_DevBuf(p.value, n.value), device=self._buf[A_SEND].device)
        per = self.yrows * self.grid.nx * 6
        full = full[: per * self.world]
        self.dist.all_gather_into_tensor(full, full[self.rank * per:(self.rank + 1) * per])

    def gather_solution(self):
        """Array(prob.sol) assembled from the column slabs of all ranks (every rank gets the full array)."""
        local = self.torch.from_numpy(np.ascontiguousarray(self._local_sol().view(np.float64))).to(self._buf[A_SEND].device)
        self.dist.all_reduce(local)                                 # slabs are disjoint, the rest is zero
        out = local.cpu().numpy().view(np.complex128).reshape(self._sol_shape(), order="F")
        return out[:, :, 0] if self.desc.model == 4 else out

    def _sol_shape(self):
        return (self.grid.nkr, self.grid.nl, self.nvar)

    def _local_sol(self):
        out = np.empty(self._sol_shape(), dtype=np.complex128, order="F")
        check(lib().swrt_flow_get_solution(self._h, out.ctypes.data_as(C.c_void_p)))
        return out.ravel(order="F")

    def energies(self):
        """(ke, pe) summed over the slabs."""
        ke, pe = C.c_double(), C.c_double()
        check(lib().swrt_flow_energies(self._h, C.byref(ke), C.byref(pe)))
        t = self.torch.tensor([ke.value, pe.value], dtype=self.torch.float64, device=self._buf[A_SEND].device)
        self.dist.all_reduce(t)
        return float(t[0]), float(t[1])
