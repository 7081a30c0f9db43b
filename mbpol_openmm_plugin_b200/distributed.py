"""Owner-molecule partition across the GPUs of one box: one process per GPU.

Per step:   positions: NCCL broadcast from rank 0      (N x 9 doubles over NVLink)
                           every rank: tuple lists of its own owner molecules + their halo, 2B and 3B of exactly those
                                       tuples -- slices are balanced by 30,601 * dimers(I) + 111,818 * trimers(I)
                           forces + energies: ONE grouped NCCL reduce to rank 0 (or all-reduce)
There is no data-path collective inside the tuple evaluation: tuples are independent.

`NcclForce` is the product: both collectives run INSIDE libmbpol_b200.so (mbpol_eval_distributed, csrc/dist.cu);
torch.distributed only ships the 128-byte NCCL id at start-up.  `DistributedForce` is the same schedule written with
torch.distributed collectives around an injected engine, so that the host logic (partition + reduction) can be
exercised on CPU with the gloo backend (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def owner_bounds(cum_weight, world):
    """Slice boundaries b_0..b_world over owner molecules from the cumulative work W (len n+1): b_r = first I with
    W[I] >= r/world * W[n].  Mirrors TupleLists::partition in csrc/lists.cu (same rule, so CPU tests can predict it)."""
    W = np.asarray(cum_weight, dtype=np.float64)
    n = len(W) - 1
    out = [0]
    for r in range(1, world):
        out.append(int(np.searchsorted(W, W[n] * r / world, side="left")))
    out.append(n)
    return out


class NcclForce:
    """2B+3B of an MBPolContext on all ranks of the default torch.distributed group, collectives behind the C ABI.

        nf = NcclForce(ctx)                      # every rank; ctx built with any rank/world (they are set here)
        energy, force = nf.step(pos)             # pos: float64 CUDA tensor on rank 0 (None elsewhere)
    Rank 0 gets (E2, E3) as a host tuple and the summed forces; with allreduce=True every rank does."""

    def __init__(self, context, allreduce=False, group=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.ctx = context
        self.L = _lib.load()
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        ident = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
            _lib.check(self.L.mbpol_comm_unique_id(buf))
            ident[0] = buf.raw
        dist.broadcast_object_list(ident, src=0, group=group)      # 128 bytes over the store: start-up only
        _lib.check(self.L.mbpol_comm_init_rank(context.engine.h, ident[0], self.rank, self.world), context.engine.h)
        dev = torch.device("cuda", torch.cuda.current_device())
        n = context.n_particles
        self.pos = torch.zeros(n, 3, dtype=torch.float64, device=dev)
        self.force = torch.zeros(n, 3, dtype=torch.float64, device=dev)
        self.flags = (context._flags & 3) | (_lib.DIST_ALLREDUCE if allreduce else 0)
        self.allreduce = allreduce
        self._e = (C.c_double * 2)()

    def step(self, positions=None, want_energy=True, broadcast=True):
        """positions: tensor (device or pinned host) on rank 0.  One C-ABI call does broadcast, evaluation and reduce."""
        torch = self.torch
        if self.rank == 0 and positions is not None:
            self.pos.copy_(positions, non_blocking=True)
        receiver = self.allreduce or self.rank == 0
        h = self.ctx.engine.h
        _lib.check(self.L.mbpol_eval_distributed(h, self.pos.data_ptr(), self.force.data_ptr(),
                                                 self._e if (want_energy and receiver) else None,
                                                 self.flags | (0 if broadcast else _lib.DIST_NO_BCAST),
                                                 torch.cuda.current_stream().cuda_stream), h)
        return ((self._e[0], self._e[1]) if (want_energy and receiver) else None), self.force

    def step_host(self, pos_host=None, force_host=None, want_energy=True):
        """host buffers on rank 0 (torch CPU tensors, ideally pinned; None elsewhere): positions in, summed forces out
        (overwritten).  The reduce is sliced and the device-to-host copy of every slice overlaps the next reduce."""
        h = self.ctx.engine.h
        r0 = self.rank == 0
        _lib.check(self.L.mbpol_eval_distributed_host(h, pos_host.data_ptr() if r0 else None,
                                                      force_host.data_ptr() if r0 else None,
                                                      self._e if (want_energy and r0) else None, self.flags & 3), h)
        return (self._e[0], self._e[1]) if (want_energy and r0) else None

    def close(self):
        if self.ctx is not None:
            self.L.mbpol_comm_destroy(self.ctx.engine.h)
            self.ctx = None


class DistributedForce:
    """The same schedule with torch.distributed collectives around an injected engine: `engine.eval(pos_tensor) ->
    (e2, e3, force_tensor)` for THIS rank's owner slice; tensors live wherever the backend wants them (cpu for gloo).
    Forces and energies travel in one packed all-reduce."""

    def __init__(self, engine, n_particles, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pos = torch.zeros(n_particles, 3, dtype=torch.float64, device=device)
        self.packed = torch.zeros(3 * n_particles + 2, dtype=torch.float64, device=device)

    def step(self, positions=None):
        """positions: tensor on rank 0 (ignored elsewhere).  Returns (energies[2], forces) with the full sums on every rank."""
        dist = self.dist
        if self.rank == 0 and positions is not None:
            self.pos.copy_(positions, non_blocking=True)
        if self.world > 1:
            dist.broadcast(self.pos, src=0, group=self.group)
        e2, e3, force = self.engine.eval(self.pos)
        n3 = self.pos.numel()
        self.packed[:n3] = force.reshape(-1)
        self.packed[n3], self.packed[n3 + 1] = e2, e3
        if self.world > 1:
            dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.group)
        return self.packed[n3:], self.packed[:n3].view(-1, 3)


class CudaSliceEngine:
    """Engine for DistributedForce on GPUs without the library's own communicator: an MBPolContext restricted to this
    rank's owner slice."""

    def __init__(self, context):
        self.ctx = context

    def eval(self, pos):
        import torch
        ctx = self.ctx
        f = torch.zeros_like(pos)
        stream = torch.cuda.current_stream().cuda_stream
        many = ctx._flags & 3                      # MBPOL_DO_2B | MBPOL_DO_3B
        e2, e3 = ctx.engine.eval_device(pos, f, many, stream=stream) if many else (0.0, 0.0)
        if ctx._f1 is not None:                    # one-body term of this rank's molecule slice rides along in e2's slot
            e2 += ctx.engine.eval_onebody(pos, f, ctx._f1.getNonbondedMethod() == ctx._f1.Periodic, stream)
        return e2, e3, f
