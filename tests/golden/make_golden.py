"""Generate the golden fixtures under tests/golden/ (run in the authoring container, where /root/reference exists).

    python tests/golden/make_golden.py

* inputs  : coordinates of the reference's own fixtures (python/water14_cluster.pdb, python/water256_bulk.pdb,
            python/pdb/H2O_1.pdb, python/tests/pdb_files/water256_integration_test.pdb) and of the dimer / trimer of
            platforms/reference/tests/TestReferenceMBPol{Two,Three}BodyForce.cpp:99-107, O/H/H only, nm
* outputs : tuple lists (canonical order), energies and forces computed by the reference's own code compiled
            unmodified (oracle/_ref/libmbpol_ref.so via oracle/refbind.py), plus random-input values of the two bare
            polynomials.
Also writes mbpol_openmm_plugin_b200/data/water512_cell.npz (the tile of the synthetic benchmark boxes).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from mbpol_openmm_plugin_b200.systems import read_pdb_waters, particle_indices  # noqa: E402
from oracle.refbind import Ref, canonical  # noqa: E402

REF = "/root/reference"
R = Ref()

# platforms/reference/tests/TestReferenceMBPolThreeBodyForce.cpp:97-107 (Angstrom); first two = the 2B test dimer
TRIMER_A = np.array([
    [-1.516074336e+00, -2.023167650e-01, 1.454672917e+00], [-6.218989773e-01, -6.009430735e-01, 1.572437625e+00],
    [-2.017613812e+00, -4.190350349e-01, 2.239642849e+00], [-1.763651687e+00, -3.816594649e-01, -1.300353949e+00],
    [-1.903851736e+00, -4.935677617e-01, -3.457810126e-01], [-2.527904158e+00, -7.613550077e-01, -1.733803676e+00],
    [-5.588472140e-01, 2.006699172e+00, -1.392786582e-01], [-9.411558180e-01, 1.541226676e+00, 6.163293071e-01],
    [-9.858551734e-01, 1.567124294e+00, -8.830970941e-01]])


def full_case(pos_mol, box, pbc, cut2, cut3):
    n = len(pos_mol)
    pos = pos_mol.reshape(-1, 3)
    idx = particle_indices(n)
    posO = pos_mol[:, 0].copy()
    pairs = canonical(R.pairs(posO, box, pbc, cut2))
    trips = canonical(R.triplets(posO, box, pbc, cut3))
    e2, f2 = R.eval_2b(pos, idx, pairs, box, pbc, cut2)
    e3, f3 = R.eval_3b(pos, idx, trips, box, pbc, cut3)
    e1, f1 = R.eval_1b(pos, idx, box, pbc)               # one-body term (MBPolReferenceOneBodyForce.cpp)
    return dict(pos_mol=pos_mol, box=np.array(box if box is not None else [0, 0, 0], dtype=np.float64), pbc=np.int32(pbc),
                cut2=np.float64(cut2), cut3=np.float64(cut3), pairs=pairs, triplets=trips, e2=np.float64(e2),
                e3=np.float64(e3), f2=f2, f3=f3, e1=np.float64(e1), f1=f1)


def main():
    out = {}
    w14 = read_pdb_waters(f"{REF}/python/water14_cluster.pdb")
    out["water14_xml"] = full_case(w14, None, False, 0.65, 0.45)
    out["water14_nocut"] = full_case(w14, None, False, 1e10, 1e10)
    L256 = 1.93996888399961804            # python/tests/TestReferenceMBPolIntegrationTest.py:71
    out["water256_bulk"] = full_case(read_pdb_waters(f"{REF}/python/water256_bulk.pdb"), [L256] * 3, True, 0.65, 0.45)
    out["water256_integration"] = full_case(
        read_pdb_waters(f"{REF}/python/tests/pdb_files/water256_integration_test.pdb"), [L256] * 3, True, 0.65, 0.45)
    L512 = 2.48343697144598               # python/utils/run_benchmark.py:37
    cell = read_pdb_waters(f"{REF}/python/pdb/H2O_1.pdb")
    out["water512_cell"] = full_case(cell, [L512] * 3, True, 0.65, 0.45)
    tri = (TRIMER_A * 0.1).reshape(3, 3, 3)
    out["test_dimer"] = full_case(tri[:2].copy(), None, False, 10.0, 10.0)
    out["test_trimer"] = full_case(tri.copy(), None, False, 10.0, 10.0)
    for shift, tag in ((False, "pbc"), (True, "pbc_shifted")):      # Test*.cpp: box 50 nm, molecule 2 moved one box in y
        t = tri.copy()
        if shift:
            t[1, :, 1] += 50.0
        out[f"test_dimer_{tag}"] = full_case(t[:2].copy(), [50.0] * 3, True, 10.0, 10.0)
        out[f"test_trimer_{tag}"] = full_case(t, [50.0] * 3, True, 10.0, 10.0)
    for name, case in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **case)
        print(f"{name}: n={len(case['pos_mol'])} pairs={len(case['pairs'])} triplets={len(case['triplets'])} "
              f"E2={float(case['e2']):.12f} E3={float(case['e3']):.12f}")

    rng = np.random.default_rng(7)
    x2 = rng.uniform(0.05, 1.5, (16, 31))
    x3 = rng.uniform(0.05, 1.5, (16, 36))
    p2 = [R.poly2b(x) for x in x2]
    p3 = [R.poly3b(x) for x in x3]
    np.savez_compressed(os.path.join(HERE, "poly_values.npz"), x2=x2, e2=np.array([p[0] for p in p2]),
                        g2=np.array([p[1] for p in p2]), x3=x3, e3=np.array([p[0] for p in p3]),
                        g3=np.array([p[1] for p in p3]))
    np.savez_compressed(os.path.join(ROOT, "mbpol_openmm_plugin_b200", "data", "water512_cell.npz"), pos_mol=cell,
                        box_nm=np.float64(L512))


if __name__ == "__main__":
    main()
