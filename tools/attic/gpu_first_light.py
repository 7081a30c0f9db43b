"""Development script: first end-to-end run on a GPU against the golden fixtures (superseded by tests/ -m gpu)."""
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mbpol_openmm_plugin_b200 import (MBPolContext, MBPolThreeBodyForce, MBPolTwoBodyForce, measure_fp64_peak)  # noqa: E402
from mbpol_openmm_plugin_b200.systems import particle_indices, synthetic_box  # noqa: E402


def make_ctx(n_mol, box, pbc, cut2, cut3, stride=3):
    f2, f3 = MBPolTwoBodyForce(), MBPolThreeBodyForce()
    for f, c in ((f2, cut2), (f3, cut3)):
        f.setCutoff(c)
        f.setNonbondedMethod(f.CutoffPeriodic if pbc else f.CutoffNonPeriodic)
        for row in particle_indices(n_mol, stride):
            f.addParticle(row.tolist())
    return MBPolContext(n_mol * stride, [f2, f3], box_nm=box if pbc else None)


print("fp64 peak (TFLOP/s, ms):", measure_fp64_peak())
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    if "poly_values" in path:
        continue
    g = np.load(path)
    n = len(g["pos_mol"])
    pbc = bool(g["pbc"])
    ctx = make_ctx(n, g["box"].tolist(), pbc, float(g["cut2"]), float(g["cut3"]))
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    e = ctx.getPotentialEnergy()
    f = ctx.getForces()
    pairs, trips = ctx.getPairs(), ctx.getTriplets()
    ok_p = np.array_equal(pairs, g["pairs"].reshape(-1, 2))
    ok_t = np.array_equal(trips, g["triplets"].reshape(-1, 3))
    e2, e3 = ctx.energies["CalcMBPolTwoBodyForce"], ctx.energies["CalcMBPolThreeBodyForce"]
    fref = g["f2"] + g["f3"]
    print(f"{os.path.basename(path):32s} pairs {ok_p} trips {ok_t}  E2 rel {abs(e2 - g['e2']) / max(abs(g['e2']), 1e-300):.2e} "
          f"E3 rel {abs(e3 - g['e3']) / max(abs(float(g['e3'])), 1e-300):.2e}  F maxabs {np.max(np.abs(f - fref)):.2e} "
          f"(|F|max {np.max(np.abs(fref)):.1f})")

for tiles in (2, 4):
    pos, box = synthetic_box(tiles)
    n = len(pos)
    ctx = make_ctx(n, box.tolist(), True, 0.65, 0.45)
    ctx.engine.set_timing(True)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    ctx.setPositions(p)
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        e = ctx.getPotentialEnergy()
        torch.cuda.synchronize()
        dt = time.time() - t0
        st = ctx.engine.stats()
    flops = st["n_pairs"] * 30601 + st["n_triplets"] * 111818
    print(f"synthetic {n}: E={e:.6f} step {dt * 1e3:.2f} ms  {n / dt:.3e} water-steps/s  {flops / dt / 1e12:.2f} TF alg  {st}")
