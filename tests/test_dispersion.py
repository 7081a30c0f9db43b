"""TT-damped C6 dispersion: oracle pinned to the reference's known answer, CUDA term against the oracle."""
import os

import numpy as np
import pytest

from conftest import load_golden

# the trimer of python/tests/pdb_files/water3.pdb with the digits the reference's C++ tests carry
# (platforms/reference/tests/TestReferenceMBPolThreeBodyForce.cpp:97-107), Angstrom
TRIMER_A = np.array([[-1.516074336e+00, -2.023167650e-01, 1.454672917e+00], [-6.218989773e-01, -6.009430735e-01, 1.572437625e+00],
                     [-2.017613812e+00, -4.190350349e-01, 2.239642849e+00], [-1.763651687e+00, -3.816594649e-01, -1.300353949e+00],
                     [-1.903851736e+00, -4.935677617e-01, -3.457810126e-01], [-2.527904158e+00, -7.613550077e-01, -1.733803676e+00],
                     [-5.588472140e-01, 2.006699172e+00, -1.392786582e-01], [-9.411558180e-01, 1.541226676e+00, 6.163293071e-01],
                     [-9.858551734e-01, 1.567124294e+00, -8.830970941e-01]])


def test_oracle_reproduces_the_reference_known_answer():
    """python/tests/TestCustomDispersion.py:14: water3, CutoffNonPeriodic, cutoff 1e3 nm: -6.84471477 kcal/mol"""
    from oracle.dispersion import eval_dispersion
    idx = np.arange(9).reshape(3, 3)
    e, f = eval_dispersion(TRIMER_A * 0.1, idx)
    assert abs(e / 4.184 - (-6.84471477)) < 5e-9                 # every printed digit (the reference's own bar is 0.01)
    assert np.abs(f.sum(axis=0)).max() < 1e-12
    # the 3-decimal PDB of the same trimer stays inside the reference's own tolerance
    ref_pdb = "/root/reference/python/tests/pdb_files/water3.pdb"
    if os.path.exists(ref_pdb):
        from mbpol_openmm_plugin_b200.forcefield import PDBFile
        pdb = PDBFile(ref_pdb)
        e_pdb, _ = eval_dispersion(pdb.positions, np.asarray(pdb.water_indices()))
        assert abs(e_pdb / 4.184 - (-6.84471477)) < 0.01
    # forces are the gradient of the energy
    p = TRIMER_A * 0.1
    for a, c in ((0, 0), (4, 2), (8, 1)):
        h = 1e-6
        pp, pm = p.copy(), p.copy()
        pp[a, c] += h
        pm[a, c] -= h
        assert abs(f[a, c] + (eval_dispersion(pp, idx)[0] - eval_dispersion(pm, idx)[0]) / (2 * h)) < 1e-6


def test_oracle_cutoff_and_periodic_images():
    from oracle.dispersion import eval_dispersion
    g = load_golden("water256_bulk")
    n = len(g["pos_mol"])
    pos = g["pos_mol"].reshape(-1, 3)
    idx = np.arange(3 * n).reshape(n, 3)
    e, f = eval_dispersion(pos, idx, g["box"], True, 0.9)
    assert -12000 < e < -3000 and np.abs(f.sum(axis=0)).max() < 1e-9
    shifted = pos.copy()
    shifted[::5] += np.array([g["box"][0], -g["box"][1], 0.0])            # whole-box translations of single ATOMS
    e2, f2 = eval_dispersion(shifted, idx, g["box"], True, 0.9)
    assert abs(e2 - e) < 1e-9 * abs(e) and np.abs(f2 - f).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("case,cutoff", [("water14_xml", 1e3), ("water256_bulk", 0.9), ("water256_bulk", 0.6)])
def test_gpu_dispersion_against_the_oracle(case, cutoff):
    import torch
    from oracle.dispersion import eval_dispersion
    from mbpol_openmm_plugin_b200 import MBPolContext, MBPolError, MBPolTwoBodyForce, CustomDispersionForce
    from mbpol_openmm_plugin_b200.systems import particle_indices, with_virtual_sites
    g = load_golden(case)
    n, pbc = len(g["pos_mol"]), bool(g["pbc"])
    pos4 = with_virtual_sites(g["pos_mol"]).reshape(-1, 3)
    if pbc:
        pos4 = pos4 - g["box"] * np.floor(pos4 / g["box"])          # atoms wrapped individually, like OpenMM may hand them over
    idx = particle_indices(n, 4)
    d = CustomDispersionForce()
    d.setNonbondedMethod(d.CutoffPeriodic if pbc else d.CutoffNonPeriodic)
    d.setCutoffDistance(cutoff)
    for row in idx:
        d.addMolecule(row.tolist())
    ctx = MBPolContext(4 * n, [d], box_nm=list(g["box"]) if pbc else None)
    e_ref, f_ref = eval_dispersion(pos4, idx, g["box"], pbc, cutoff)
    for positions in (pos4, torch.as_tensor(pos4).cuda()):           # host arrays and device tensors
        ctx.setPositions(positions)
        e = ctx.getPotentialEnergy()
        f = ctx.getForces()
        f = f if isinstance(f, np.ndarray) else f.cpu().numpy()
        assert abs(e - e_ref) <= 1e-9 * abs(e_ref)
        assert np.abs(f - f_ref).max() < 1e-6
        assert np.all(f.reshape(n, 4, 3)[:, 3] == 0.0)               # M sites are excluded (mbpol.xml:101)
    assert ctx.energies["CustomDispersionForce"] == e
    if pbc:
        with pytest.raises(MBPolError):
            bad = CustomDispersionForce()
            bad.setNonbondedMethod(bad.CutoffPeriodic)
            bad.setCutoffDistance(0.6 * float(g["box"][0]))
            for row in idx:
                bad.addMolecule(row.tolist())
            c2 = MBPolContext(4 * n, [bad], box_nm=list(g["box"]))
            c2.setPositions(pos4)
            c2.getPotentialEnergy()
    # together with 2B in one context: terms add
    f2 = MBPolTwoBodyForce()
    f2.setCutoff(float(g["cut2"]))
    f2.setNonbondedMethod(f2.CutoffPeriodic if pbc else f2.CutoffNonPeriodic)
    for row in idx:
        f2.addParticle(row.tolist())
    both = MBPolContext(4 * n, [f2, d], box_nm=list(g["box"]) if pbc else None)
    both.setPositions(pos4)
    assert abs(both.getPotentialEnergy() - (e_ref + float(g["e2"]))) <= 1e-9 * abs(e_ref + float(g["e2"]))
