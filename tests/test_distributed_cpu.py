"""world_size-2 gloo test of the multi-GPU host logic (broadcast positions -> owner slices -> all-reduce), on CPU.

The engine is a stand-in built on the C port oracle (tests may use oracle/): it evaluates the tuples of an owner slice
chosen with the same balancing rule the CUDA library uses.  Checks: slices are disjoint and complete, the all-reduced
result equals the single-rank result, non-root ranks receive the broadcast positions.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


class PortSliceEngine:
    def __init__(self, g, rank, world):
        from oracle.refbind import Port
        from mbpol_openmm_plugin_b200.distributed import owner_bounds
        self.port = Port()
        self.g = g
        n = len(g["pos_mol"])
        pairs, trips = g["pairs"].reshape(-1, 2), g["triplets"].reshape(-1, 3)
        # cumulative work per owner, exactly the library's model: 30601 * pairs + 111818 * triplets owned by molecules < I
        w = np.zeros(n + 1)
        np.add.at(w, pairs[:, 0] + 1, 30601.0)
        np.add.at(w, trips[:, 0] + 1, 111818.0)
        b = owner_bounds(np.cumsum(w), world)
        lo, hi = b[rank], b[rank + 1]
        self.pairs = pairs[(pairs[:, 0] >= lo) & (pairs[:, 0] < hi)]
        self.trips = trips[(trips[:, 0] >= lo) & (trips[:, 0] < hi)]
        self.idx = np.arange(3 * n).reshape(n, 3)

    def eval(self, pos):
        p = pos.numpy()
        e2, f2 = self.port.eval_2b(p, self.idx, self.pairs, self.g["box"], bool(self.g["pbc"]))
        e3, f3 = self.port.eval_3b(p, self.idx, self.trips, self.g["box"], bool(self.g["pbc"]))
        return e2, e3, torch.from_numpy(f2 + f3)


def worker(rank, world, port_no, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mbpol_openmm_plugin_b200.distributed import DistributedForce
    g = load_golden("water14_xml")
    eng = PortSliceEngine(g, rank, world)
    d = DistributedForce(eng, 3 * len(g["pos_mol"]), "cpu")
    pos = torch.from_numpy(g["pos_mol"].reshape(-1, 3).copy()) if rank == 0 else None
    energy, force = d.step(pos)
    out.put((rank, energy.numpy().copy(), force.numpy().copy(), len(eng.pairs), len(eng.trips)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_owner_partition_allreduce_gloo(world):
    g = load_golden("water14_xml")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[3] for r in res) == len(g["pairs"]) and sum(r[4] for r in res) == len(g["triplets"])
    for _, e, f, _, _ in res:                      # every rank holds the full sums
        assert abs(e[0] - g["e2"]) <= 1e-9 * abs(g["e2"]) and abs(e[1] - g["e3"]) <= 1e-9 * abs(g["e3"])
        assert np.abs(f - (g["f2"] + g["f3"])).max() < 1e-6
    assert np.array_equal(res[0][2], res[1][2])    # bit-identical after the all-reduce


def test_owner_bounds_balance():
    from mbpol_openmm_plugin_b200.distributed import owner_bounds
    rng = np.random.default_rng(0)
    w = np.concatenate([[0.0], np.cumsum(rng.uniform(0, 1, 5000) * np.linspace(0.1, 2.0, 5000))])
    for world in (1, 2, 4, 8):
        b = owner_bounds(w, world)
        assert b[0] == 0 and b[-1] == 5000 and all(b[i] <= b[i + 1] for i in range(world))
        loads = np.diff(w[b])
        assert loads.max() / loads.mean() < 1.01
