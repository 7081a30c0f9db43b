"""One-body (Partridge-Schwenke) term, SURVEY.md 8f rank 4: oracle pinned to the reference's own test vector, host
mirror of MBPolOneBodyForce, and GPU parity through the C ABI (mbpol_eval_onebody)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, make_context

CAL = 4.184
# platforms/reference/tests/TestReferenceMBPolOneBodyForce.cpp:80-105: positions (Angstrom), gradients (kcal/mol/A),
# energy (kcal/mol); the reference's own tolerance is 1e-4
POS_A = np.array([[-1.516074336e+00, -2.023167650e-01, 1.454672917e+00],
                  [-6.218989773e-01, -6.009430735e-01, 1.572437625e+00],
                  [-2.017613812e+00, -4.190350349e-01, 2.239642849e+00]])
GRAD = np.array([[-27.48162433, 8.92495995, 2.80995323], [30.78909844, -11.48714187, -0.27204770],
                 [-3.30747410, 2.56218193, -2.53790553]])
ENERGY = 0.55975882


def _oracles(request):
    from oracle.refbind import have_ref
    return [request.getfixturevalue("port")] + ([request.getfixturevalue("ref")] if have_ref() else [])


def test_reference_onebody_vector(request):
    idx = np.array([[0, 1, 2]], dtype=np.int32)
    for o in _oracles(request):
        e, f = o.eval_1b(POS_A * 0.1, idx, None, False)
        assert abs(e / CAL - ENERGY) < 5e-9
        np.testing.assert_allclose(-f / (CAL * 10), GRAD, atol=6e-9)
        # periodic: a hydrogen moved by one box edge is imaged back to its oxygen (MBPolReferenceOneBodyForce.cpp:199-200)
        p = POS_A * 0.1
        p[1, 0] += 3.0
        p[2, 2] -= 6.0
        e2, f2 = o.eval_1b(p, idx, [3.0, 4.0, 6.0], True)
        assert abs(e2 - e) < 1e-9 * abs(e)
        np.testing.assert_allclose(f2, f, atol=1e-7)


@pytest.mark.parametrize("case", golden_cases())
def test_port_matches_committed_onebody_fixture(case, port):
    g = load_golden(case)
    n = len(g["pos_mol"])
    e, f = port.eval_1b(g["pos_mol"].reshape(-1, 3), np.arange(3 * n).reshape(n, 3), g["box"], bool(g["pbc"]))
    assert abs(e - g["e1"]) <= 1e-11 * max(1.0, abs(float(g["e1"])))
    assert np.abs(f - g["f1"]).max() < 1e-8


def test_onebody_force_container():
    from mbpol_openmm_plugin_b200 import B200CalcMBPolOneBodyForceKernel, MBPolError, MBPolOneBodyForce
    f = MBPolOneBodyForce()
    assert f.getNumOneBodys() == 0 and f.getNonbondedMethod() == f.NonPeriodic
    assert f.addOneBody([0, 1, 2]) == 0 and f.addOneBody([4, 5, 6]) == 1
    assert f.getOneBodyParameters(1) == [4, 5, 6]
    f.setOneBodyParameters(1, [8, 9, 10])
    assert f.getOneBodyParameters(1) == [8, 9, 10]
    f.setNonbondedMethod(f.Periodic)
    assert f.getNonbondedMethod() == MBPolOneBodyForce.Periodic
    with pytest.raises(MBPolError):
        f.setNonbondedMethod(7)
    with pytest.raises(MBPolError):
        f.addOneBody([0, 1])
    assert B200CalcMBPolOneBodyForceKernel.Name() == "CalcMBPolOneBodyForce"


@pytest.mark.gpu
@pytest.mark.parametrize("case", golden_cases())
def test_gpu_onebody_golden(case):
    g = load_golden(case)
    n = len(g["pos_mol"])
    ctx = make_context(n, g["box"], bool(g["pbc"]), float(g["cut2"]), float(g["cut3"]), which=(1,))
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    e = ctx.getPotentialEnergy()
    assert abs(e - g["e1"]) <= 1e-9 * max(abs(float(g["e1"])), 1e-3)
    assert np.abs(ctx.getForces() - g["f1"]).max() < 1e-6


@pytest.mark.gpu
def test_gpu_onebody_with_2b_3b_device_tensors_and_wrapped_atoms(checker):
    """all three terms in one context on device tensors, atoms wrapped individually into the box (hydrogens split
    from their oxygen across the boundary), O/H/H/M particle stride"""
    import torch
    from mbpol_openmm_plugin_b200.systems import with_virtual_sites
    g = load_golden("water256_bulk")
    n = len(g["pos_mol"])
    box = g["box"]
    pos4 = with_virtual_sites(g["pos_mol"]).reshape(-1, 3)
    pos4 = pos4 - box * np.floor(pos4 / box)                  # every ATOM in [0, L): molecules straddle the faces
    ctx = make_context(n, box, True, stride=4, which=(1, 2, 3))
    ctx.setPositions(torch.as_tensor(pos4).cuda())
    e = ctx.getPotentialEnergy()
    idx4 = np.arange(4 * n).reshape(n, 4)[:, :3]
    e1, f1 = checker.eval_1b(pos4, idx4, box, True)
    assert abs(ctx.energies["CalcMBPolOneBodyForce"] - e1) <= 1e-9 * max(abs(e1), 1e-3)
    assert abs(e - (e1 + ctx.energies["CalcMBPolTwoBodyForce"] + ctx.energies["CalcMBPolThreeBodyForce"])) < 1e-9
    assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - g["e2"]) <= 1e-9 * abs(float(g["e2"]))
    f = ctx.getForces().cpu().numpy().reshape(n, 4, 3)
    assert np.all(f[:, 3] == 0.0)
    # the oracle on the SAME wrapped coordinates (the fixture's forces belong to the unwrapped ones)
    _, f2 = checker.eval_2b(pos4, idx4, g["pairs"].reshape(-1, 2), box, True)[:2]
    _, f3 = checker.eval_3b(pos4, idx4, g["triplets"].reshape(-1, 3), box, True)[:2]
    want = (f1 + f2 + f3).reshape(n, 4, 3)[:, :3]
    assert np.abs(f[:, :3] - want).max() < 1e-6
    # force groups: the one-body kernel alone
    ctx.forces[0].setForceGroup(3)
    assert abs(ctx.getPotentialEnergy(groups=1 << 3) - e1) <= 1e-9 * max(abs(e1), 1e-3)
    assert np.abs(ctx.getForces().cpu().numpy()[idx4.reshape(-1)] - f1[idx4.reshape(-1)]).max() < 1e-6


@pytest.mark.gpu
def test_gpu_onebody_distorted_molecules_and_partition(checker):
    """stretched / bent monomers (the fixtures are near-rigid) and the rank slices of mbpol_set_partition"""
    rng = np.random.default_rng(11)
    g = load_golden("water14_xml")
    pos = g["pos_mol"].copy()
    pos[:, 1:] += rng.normal(scale=0.008, size=pos[:, 1:].shape)         # +-0.08 A on every hydrogen
    n = len(pos)
    idx = np.arange(3 * n).reshape(n, 3)
    e_ref, f_ref = checker.eval_1b(pos.reshape(-1, 3), idx, None, False)
    ctx = make_context(n, None, False, which=(1,))
    ctx.setPositions(pos.reshape(-1, 3))
    assert abs(ctx.getPotentialEnergy() - e_ref) <= 1e-9 * abs(e_ref)
    assert np.abs(ctx.getForces() - f_ref).max() < 1e-6
    e_sum, f_sum = 0.0, np.zeros((3 * n, 3))
    for r in range(3):
        part = make_context(n, None, False, which=(1,), rank=r, world=3)
        part.setPositions(pos.reshape(-1, 3))
        e_sum += part.getPotentialEnergy()
        f_sum += part.getForces()
    assert abs(e_sum - e_ref) <= 1e-9 * abs(e_ref)
    assert np.abs(f_sum - f_ref).max() < 1e-6
