"""Pin the oracles before trusting them:

* oracle/_ref (the reference's own sources, compiled unmodified) and the C port oracle/mbpol_oracle.c must both
  reproduce the golden vectors the reference's own tests hold for this path;
* the port must agree with the reference build on the committed fixtures;
* the committed fixtures must be what the reference build produces today (regeneration check).
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden

# platforms/reference/tests/TestReferenceMBPolTwoBodyForce.cpp:120-127 (gradients, kcal/mol/A; energy kcal/mol)
DIMER_GRAD = np.array([[-4.85337479, -4.47836379, -20.08989563], [-0.31239868, 0.52518586, -1.88893830],
                       [0.00886712, 0.73323536, -1.81715325], [-0.65181727, -0.72947395, 5.88973293],
                       [4.82340981, 3.20090213, 16.49522051], [0.98531382, 0.74851439, 1.41103374]])
DIMER_E = 6.14207815
# platforms/reference/tests/TestReferenceMBPolThreeBodyForce.cpp:122-138
TRIMER_GRAD = np.array([[0.29919011, -0.34960381, -0.16238472], [0.34138467, -0.01255068, -0.00998383],
                        [-0.44376649, 0.03687577, 0.54604510], [-0.01094164, -0.36171476, -0.05130395],
                        [0.24939202, 1.29382952, 0.22930712], [-0.13250943, -0.19313418, -0.34123592],
                        [0.56722869, 0.46036139, -0.39999973], [-0.75669111, -0.76132457, -0.29799486],
                        [-0.11328682, -0.11273867, 0.48755080]])
TRIMER_E = 0.15586446
CAL = 4.184


def oracles(request):
    return [request.getfixturevalue("port")] + ([request.getfixturevalue("ref")] if _have_ref() else [])


def _have_ref():
    from oracle.refbind import have_ref
    return have_ref()


@pytest.mark.parametrize("case", ["test_dimer", "test_dimer_pbc", "test_dimer_pbc_shifted"])
def test_reference_dimer_vector(case, request):
    g = load_golden(case)
    pos = g["pos_mol"].reshape(-1, 3)
    idx = np.arange(6).reshape(2, 3)
    for o in oracles(request):
        pairs = o.pairs(g["pos_mol"][:, 0].copy(), g["box"], bool(g["pbc"]), float(g["cut2"]))
        assert pairs.tolist() == [[1, 0]]
        e, f = o.eval_2b(pos, idx, pairs, g["box"], bool(g["pbc"]))[:2]
        # the reference's own tolerance is 1e-3 (TestReferenceMBPolTwoBodyForce.cpp:146); we hold all printed digits
        assert abs(e / CAL - DIMER_E) < 5e-9
        # molecule a = higher index = second molecule of the test: forces come back per particle, so order is global
        np.testing.assert_allclose(-f / (CAL * 10), DIMER_GRAD, atol=6e-8)


@pytest.mark.parametrize("case", ["test_trimer", "test_trimer_pbc", "test_trimer_pbc_shifted"])
def test_reference_trimer_vector(case, request):
    g = load_golden(case)
    pos = g["pos_mol"].reshape(-1, 3)
    idx = np.arange(9).reshape(3, 3)
    for o in oracles(request):
        trips = o.triplets(g["pos_mol"][:, 0].copy(), g["box"], bool(g["pbc"]), float(g["cut3"]))
        assert trips.tolist() == [[2, 1, 0]]
        e, f = o.eval_3b(pos, idx, trips, g["box"], bool(g["pbc"]))[:2]
        assert abs(e / CAL - TRIMER_E) < 5e-9
        np.testing.assert_allclose(-f / (CAL * 10), TRIMER_GRAD, atol=6e-8)


def test_reference_neighbor_list_counts(request):
    """platforms/reference/tests/TestReferenceThreeNeighborList.cpp:42-78 (incl. coincident points, minDistance 0)"""
    three = np.array([[13.6, 0, 0], [0, 0, 0], [2, 0, 0]], dtype=float)
    four = np.array([[14, 0, 0], [10, 0, 0], [10, 0, 0], [10, 0, 0]], dtype=float)
    for o in oracles(request):
        assert len(o.triplets(three, None, False, 13.7)) == 1
        assert len(o.triplets(three, None, False, 13.5)) == 0
        assert len(o.triplets(four, None, False, 12.0)) == 4
        assert len(o.triplets(four, None, False, 1.0)) == 1


@pytest.mark.parametrize("case", golden_cases())
def test_port_matches_committed_fixture(case, port):
    """C port vs the values the reference build wrote into tests/golden (lists bit-exact, E 1e-9 rel, F 1e-6 abs)."""
    g = load_golden(case)
    n = len(g["pos_mol"])
    if n > 300:
        pytest.skip("port is O(n^2) / slow: covered by the 256-water cases")
    posO = g["pos_mol"][:, 0].copy()
    pbc = bool(g["pbc"])
    pairs = port.pairs(posO, g["box"], pbc, float(g["cut2"]))
    trips = port.triplets(posO, g["box"], pbc, float(g["cut3"]))
    assert np.array_equal(pairs, g["pairs"].reshape(-1, 2))
    assert np.array_equal(trips, g["triplets"].reshape(-1, 3))
    if n > 64 and case != "water256_bulk":
        return                                  # keep the CPU suite short: one full 256-water force check
    pos = g["pos_mol"].reshape(-1, 3)
    idx = np.arange(3 * n).reshape(n, 3)
    e2, f2 = port.eval_2b(pos, idx, pairs, g["box"], pbc)
    e3, f3 = port.eval_3b(pos, idx, trips, g["box"], pbc)
    assert abs(e2 - g["e2"]) <= 1e-9 * max(abs(g["e2"]), 1e-12)
    assert abs(e3 - g["e3"]) <= 1e-9 * max(abs(g["e3"]), 1e-12)
    assert np.abs(f2 - g["f2"]).max() < 1e-6 and np.abs(f3 - g["f3"]).max() < 1e-6


@pytest.mark.parametrize("case", ["water14_xml", "water14_nocut", "water256_bulk", "test_trimer_pbc_shifted"])
def test_fixture_is_what_the_reference_build_produces(case, ref):
    from oracle.refbind import canonical
    g = load_golden(case)
    n = len(g["pos_mol"])
    posO = g["pos_mol"][:, 0].copy()
    pbc = bool(g["pbc"])
    pairs = canonical(ref.pairs(posO, g["box"], pbc, float(g["cut2"])))
    trips = canonical(ref.triplets(posO, g["box"], pbc, float(g["cut3"])))
    assert np.array_equal(pairs, g["pairs"].reshape(-1, 2)) and np.array_equal(trips, g["triplets"].reshape(-1, 3))
    idx = np.arange(3 * n).reshape(n, 3)
    e2, f2 = ref.eval_2b(g["pos_mol"].reshape(-1, 3), idx, pairs, g["box"], pbc, float(g["cut2"]))
    e3, f3 = ref.eval_3b(g["pos_mol"].reshape(-1, 3), idx, trips, g["box"], pbc, float(g["cut3"]))
    assert e2 == g["e2"] and e3 == g["e3"]
    assert np.array_equal(f2, g["f2"]) and np.array_equal(f3, g["f3"])


def test_survey_anchor_values():
    """SURVEY.md App. C anchors (independently measured during the survey)."""
    w14 = load_golden("water14_xml")
    assert (len(w14["pairs"]), len(w14["triplets"])) == (79, 57)
    assert abs(w14["e2"] - 360.672617662985) < 1e-9 and abs(w14["e3"] - 5.301730089748) < 1e-9
    nc = load_golden("water14_nocut")
    assert (len(nc["pairs"]), len(nc["triplets"])) == (91, 364) and abs(nc["e3"] - 9.206553638625) < 1e-9
    b = load_golden("water256_bulk")
    assert (len(b["pairs"]), len(b["triplets"])) == (4880, 7184)
    assert abs(b["e2"] - 13846.529641941346) < 1e-8 and abs(b["e3"] - 234.565776070133) < 1e-9
