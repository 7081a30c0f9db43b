"""Owner-molecule partition across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Per step:   positions: broadcast from rank 0          (N x 9 doubles, NCCL over NVLink)
                           every rank: build the (replicated, cheap) tuple lists, evaluate ONLY the tuples whose owner
                                       molecule I (highest index of the tuple) falls in its slice -- slices are balanced
                                       by 30,601 * dimers(I) + 111,818 * trimers(I) inside the C library
                           forces, energies: all-reduce (sum)         (N x 9 + 2 doubles)
There is no data-path collective inside the tuple evaluation: tuples are independent.  The reduction order of the
all-reduce is fixed for a given world size, so results are reproducible run to run at fixed N.

The engine is injected so that the host logic can be exercised on CPU with the gloo backend (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import numpy as np


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


class DistributedForce:
    """Evaluate 2B+3B on `world` ranks.  `engine` must offer eval(pos_tensor) -> (e2, e3, force_tensor) for THIS rank's
    owner slice; tensors live wherever the process group's backend wants them (cuda for nccl, cpu for gloo)."""

    def __init__(self, engine, n_particles, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pos = torch.zeros(n_particles, 3, dtype=torch.float64, device=device)
        self.energy = torch.zeros(2, dtype=torch.float64, device=device)

    def step(self, positions=None):
        """positions: tensor on rank 0 (ignored elsewhere).  Returns (e2, e3, forces) with the full sums on every rank."""
        dist = self.dist
        if self.rank == 0 and positions is not None:
            self.pos.copy_(positions, non_blocking=True)
        if self.world > 1:
            dist.broadcast(self.pos, src=0, group=self.group)
        e2, e3, force = self.engine.eval(self.pos)
        self.energy[0], self.energy[1] = e2, e3
        if self.world > 1:
            dist.all_reduce(force, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.energy, op=dist.ReduceOp.SUM, group=self.group)
        return self.energy, force


class CudaSliceEngine:
    """The product engine for DistributedForce: an MBPolContext restricted to this rank's owner slice."""

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
