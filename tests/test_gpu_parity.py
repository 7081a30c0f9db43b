"""Parity of the CUDA path (through the C ABI) against the oracles, on a real GPU.

Bars (BASELINE.json north_star): tuple lists bit-exact; energies within 1e-9 relative; forces within 1e-6 kJ/mol/nm
max-abs.  The checker is the reference's own code (oracle/_ref) when its build travelled with the repo, else the C port.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_cases, load_golden, make_context

pytestmark = pytest.mark.gpu

E_RTOL = 1e-9          # relative energy tolerance
F_ATOL = 1e-6          # kJ/mol/nm


def check_case(ctx, g, e_keys=("e2", "e3")):
    e2, e3 = ctx.energies["CalcMBPolTwoBodyForce"], ctx.energies["CalcMBPolThreeBodyForce"]
    assert abs(e2 - g["e2"]) <= E_RTOL * max(abs(float(g["e2"])), 1e-9)
    assert abs(e3 - g["e3"]) <= E_RTOL * max(abs(float(g["e3"])), 1e-9)


@pytest.mark.parametrize("case", golden_cases())
def test_golden_fixture_host_arrays(case):
    """C-ABI call with HOST buffers (mbpol_eval_host): lists, energies, forces vs the committed reference outputs."""
    g = load_golden(case)
    n = len(g["pos_mol"])
    ctx = make_context(n, g["box"], bool(g["pbc"]), float(g["cut2"]), float(g["cut3"]))
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    ctx.getPotentialEnergy()
    assert np.array_equal(ctx.getPairs(), g["pairs"].reshape(-1, 2))            # bit-exact, canonical order
    assert np.array_equal(ctx.getTriplets(), g["triplets"].reshape(-1, 3))
    check_case(ctx, g)
    f = ctx.getForces()
    assert isinstance(f, np.ndarray)
    assert np.abs(f - (g["f2"] + g["f3"])).max() < F_ATOL


@pytest.mark.parametrize("case", ["water14_xml", "water256_bulk", "test_trimer_pbc_shifted"])
def test_golden_fixture_device_tensors_and_m_sites(case, checker):
    """Device-resident positions, 4 particles per molecule (O,H1,H2,M) offset by +25 nm like
    platforms/cuda/tests/TestCudaMBPolThreeBodyForce.cpp:66-150: M sites get exactly zero force.  Periodic cases are
    evaluated on a translated image of the whole system; the oracle is run on the SAME translated coordinates."""
    import torch
    from mbpol_openmm_plugin_b200.systems import with_virtual_sites
    g = load_golden(case)
    n = len(g["pos_mol"])
    pos4 = with_virtual_sites(g["pos_mol"])
    if bool(g["pbc"]):
        pos4 = pos4 + np.array([float(g["box"][0]), 0.0, -2 * float(g["box"][2])])          # image of the whole system
    ctx = make_context(n, g["box"], bool(g["pbc"]), float(g["cut2"]), float(g["cut3"]), stride=4)
    ctx.setPositions(torch.as_tensor(pos4).cuda())
    ctx.getPotentialEnergy()
    f = ctx.getForces().cpu().numpy().reshape(n, 4, 3)
    assert np.all(f[:, 3] == 0.0)
    check_case(ctx, g)
    want = g["f2"] + g["f3"]
    if bool(g["pbc"]):
        from mbpol_openmm_plugin_b200.systems import particle_indices
        idx = particle_indices(n, 4)
        _, f2 = checker.eval_2b(pos4, idx, g["pairs"].reshape(-1, 2), g["box"], True)[:2]
        _, f3 = checker.eval_3b(pos4, idx, g["triplets"].reshape(-1, 3), g["box"], True)[:2]
        want = (f2 + f3).reshape(n, 4, 3)[:, :3].reshape(-1, 3)
    assert np.abs(f[:, :3].reshape(-1, 3) - want).max() < F_ATOL


def test_reference_test_vectors_on_gpu():
    """The numbers printed in the reference's own C++ tests (kcal/mol, kcal/mol/A), tolerance tightened to 5e-8."""
    from test_oracle import DIMER_E, DIMER_GRAD, TRIMER_E, TRIMER_GRAD
    for case, which, E, G in (("test_dimer", 2, DIMER_E, DIMER_GRAD), ("test_trimer", 3, TRIMER_E, TRIMER_GRAD)):
        g = load_golden(case)
        ctx = make_context(len(g["pos_mol"]), None, False, 10.0, 10.0, which=(which,))
        ctx.setPositions(g["pos_mol"].reshape(-1, 3))
        e = ctx.getPotentialEnergy()
        assert abs(e / 4.184 - E) < 5e-9
        np.testing.assert_allclose(-ctx.getForces() / 41.84, G, atol=6e-8)


@pytest.mark.parametrize("order", [2, 3])
def test_bare_polynomial(order):
    import os
    import torch
    from conftest import GOLDEN
    from mbpol_openmm_plugin_b200 import pip_eval
    g = np.load(os.path.join(GOLDEN, "poly_values.npz"))
    k = str(order)
    x = torch.as_tensor(np.ascontiguousarray(g["x" + k].T)).cuda()
    e, grad = pip_eval(order, x)
    e, grad = e.cpu().numpy(), grad.cpu().numpy().T
    assert np.abs(e - g["e" + k]).max() <= 1e-12 * np.abs(g["e" + k]).max()
    assert np.abs(grad - g["g" + k]).max() <= 1e-12 * np.abs(g["g" + k]).max()


@pytest.mark.parametrize("order,n", [(2, 1), (2, 385), (2, 1000), (3, 1), (3, 257), (3, 1000)])
def test_bare_polynomial_ragged_batches_and_relabelling(order, n):
    """the generated evaluators on batches that do not fill their last CTA (384 / 256 tuples per CTA), against the monomial
    tables on the CPU; and the symmetry the 3B evaluator is built on: relabelling the molecules permutes the variables, the
    energy must not move and the gradient must move with the variables (tools/gen_msym.py evaluates the six relabellings
    of every trimer as six contexts of one thread)"""
    import os
    import sys
    import torch
    from conftest import ROOT
    from mbpol_openmm_plugin_b200 import pip_eval
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import extract_pip
    import gen_msym
    name = {2: "pip2b", 3: "pip3b"}[order]
    t = dict(np.load(os.path.join(ROOT, "mbpol_openmm_plugin_b200", "data", name + ".npz")))
    nv = int(t["n_vars"])
    x = np.random.default_rng(100 * order + n).uniform(0.05, 1.5, (nv, n))
    e, g = pip_eval(order, torch.as_tensor(x).cuda())
    e, g = e.cpu().numpy(), g.cpu().numpy()
    for i in sorted({0, n // 2, n - 1}):
        er, gr = extract_pip.eval_table(t, x[:, i])
        assert abs(e[i] - er) <= 1e-12 * max(1.0, abs(er))
        assert np.abs(g[:, i] - gr).max() <= 1e-12 * np.abs(gr).max()
    for pm in list(gen_msym.mol_perms(name).values())[1:]:
        pm = np.asarray(pm)
        ep, gp = pip_eval(order, torch.as_tensor(np.ascontiguousarray(x[pm])).cuda())
        assert np.abs(ep.cpu().numpy() - e).max() <= 1e-12 * max(1.0, np.abs(e).max())
        assert np.abs(gp.cpu().numpy() - g[pm]).max() <= 1e-12 * np.abs(g).max()


@pytest.mark.parametrize("case", ["water14_nocut", "water256_bulk"])
def test_per_tuple_parity(case, checker):
    """energy and 6x3 / 9x3 force block of every single tuple against the oracle (before any summation)"""
    from mbpol_openmm_plugin_b200 import eval_tuples
    g = load_golden(case)
    pbc = bool(g["pbc"])
    box = g["box"] if pbc else None
    n = len(g["pos_mol"])
    pos = g["pos_mol"].reshape(-1, 3)
    idx = np.arange(3 * n).reshape(n, 3)
    for order, tuples in ((2, g["pairs"].reshape(-1, 2)), (3, g["triplets"].reshape(-1, 3))):
        e, f = eval_tuples(order, g["pos_mol"], tuples, box)
        if checker.kind == "reference":
            er, fr = (checker.each_2b if order == 2 else checker.each_3b)(g["pos_mol"], tuples, box, pbc)
        else:
            _, _, er, fr = (checker.eval_2b if order == 2 else checker.eval_3b)(pos, idx, tuples, box, pbc, each=True)
        assert np.abs(e - er).max() <= 1e-9 * max(np.abs(er).max(), 1e-9)
        assert np.abs(f - fr).max() < F_ATOL


def test_lists_edge_cases():
    """coincident points, tiny systems, empty lists (TestReferenceThreeNeighborList.cpp:60-78 on the GPU list build)"""
    base = np.array([[0.0, 0, 0], [0.09, 0.03, 0], [-0.03, 0.09, 0]])
    O = np.array([[1.4, 0, 0], [1.0, 0, 0], [1.0, 0, 0], [1.0, 0, 0]])
    pos = O[:, None, :] + base[None]
    ctx = make_context(4, None, False, 1.2, 1.2)
    ctx.setPositions(pos.reshape(-1, 3))
    ctx.getPotentialEnergy()                      # coincident oxygens: r < 2 A early-out -> zero, but they are LISTED
    assert len(ctx.getTriplets()) == 4 and len(ctx.getPairs()) == 6
    ctx = make_context(4, None, False, 0.1, 0.1)
    ctx.setPositions(pos.reshape(-1, 3))
    assert ctx.getPotentialEnergy() == 0.0
    assert ctx.getTriplets().tolist() == [[3, 2, 1]] and len(ctx.getPairs()) == 3
    assert not np.any(ctx.getForces())
    for n in (1, 2):                              # one molecule: nothing; two: a pair, no triplet
        ctx = make_context(n, None, False, 1.0, 1.0)
        ctx.setPositions((pos[:n] + np.arange(n)[:, None, None] * 0.3).reshape(-1, 3))
        ctx.getPotentialEnergy()
        assert len(ctx.getPairs()) == n - 1 and len(ctx.getTriplets()) == 0


def test_run_to_run_determinism_is_bitwise():
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    ctx = make_context(len(pos), box, True)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    out = []
    for _ in range(3):
        ctx.setPositions(p)
        e = ctx.getPotentialEnergy()
        out.append((e, ctx.getForces().clone()))
    assert out[0][0] == out[1][0] == out[2][0]
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][1], out[2][1])


def test_synthetic_4096_against_oracle(checker):
    """BASELINE config 3: full lists bit-exact, energies to 1e-9, a force sample to 1e-6; plus invariants."""
    import torch
    from oracle.refbind import canonical
    from mbpol_openmm_plugin_b200.systems import particle_indices, synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    ctx = make_context(n, box, True)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    ctx.setPositions(p)
    e = ctx.getPotentialEnergy()
    f = ctx.getForces().cpu().numpy()
    pairs, trips = ctx.getPairs(), ctx.getTriplets()
    posO = pos[:, 0].copy()
    if checker.kind == "reference":               # the port's list build is O(n^2): use it only if it must be
        assert np.array_equal(pairs, canonical(checker.pairs(posO, box, True, 0.65)))
        assert np.array_equal(trips, canonical(checker.triplets(posO, box, True, 0.45)))
    assert 17.5 * n < len(pairs) < 19.5 * n and 23 * n < len(trips) < 28 * n        # SURVEY.md 8d densities
    # sorted, strictly ordered, unique
    assert np.all(pairs[:, 0] > pairs[:, 1]) and np.all(trips[:, 0] > trips[:, 1]) and np.all(trips[:, 1] > trips[:, 2])
    key = trips[:, 0].astype(np.int64) * n * n + trips[:, 1].astype(np.int64) * n + trips[:, 2]
    assert np.all(np.diff(key) > 0)
    # per-tuple energies of a random sample against the oracle, and the full sums if the reference build is here
    from mbpol_openmm_plugin_b200 import eval_tuples
    rng = np.random.default_rng(0)
    idx = particle_indices(n)
    for order, tuples in ((2, pairs), (3, trips)):
        sel = np.sort(rng.choice(len(tuples), 1500, replace=False))
        eg, fg = eval_tuples(order, pos, tuples[sel], box)
        if checker.kind == "reference":
            er, fr = (checker.each_2b if order == 2 else checker.each_3b)(pos, tuples[sel], box, True)
        else:
            _, _, er, fr = (checker.eval_2b if order == 2 else checker.eval_3b)(pos.reshape(-1, 3), idx, tuples[sel], box,
                                                                                 True, each=True)
        assert np.abs(eg - er).max() <= 1e-9 * np.abs(er).max() and np.abs(fg - fr).max() < F_ATOL
    if checker.kind == "reference":
        e2, f2 = checker.eval_2b(pos.reshape(-1, 3), idx, pairs, box, True, 0.65)
        e3, f3 = checker.eval_3b(pos.reshape(-1, 3), idx, trips, box, True, 0.45)
        assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - e2) <= E_RTOL * abs(e2)
        assert abs(ctx.energies["CalcMBPolThreeBodyForce"] - e3) <= E_RTOL * abs(e3)
        assert np.abs(f - (f2 + f3)).max() < F_ATOL
    # size-independent properties: Newton's third law and invariance under whole-box translations of molecules
    assert np.abs(f.sum(axis=0)).max() < 1e-6
    shifted = pos.copy()
    shifted[::7] += np.array([box[0], 0, 0])
    shifted[::5] -= np.array([0, 2 * box[1], box[2]])
    ctx.setPositions(torch.as_tensor(shifted.reshape(-1, 3)).cuda())
    e_s = ctx.getPotentialEnergy()
    assert abs(e_s - e) <= 1e-9 * abs(e)
    assert np.array_equal(ctx.getPairs(), pairs) and np.array_equal(ctx.getTriplets(), trips)
    # forces of the translated system against the oracle evaluated on the SAME translated coordinates
    from oracle.refbind import eval_parallel
    e2s, f2s = eval_parallel(checker, 2, shifted.reshape(-1, 3), idx, pairs, box, True)
    e3s, f3s = eval_parallel(checker, 3, shifted.reshape(-1, 3), idx, trips, box, True)
    assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - e2s) <= E_RTOL * abs(e2s)
    assert abs(ctx.energies["CalcMBPolThreeBodyForce"] - e3s) <= E_RTOL * abs(e3s)
    assert np.abs(ctx.getForces().cpu().numpy() - (f2s + f3s)).max() < F_ATOL


def test_owner_partition_sums_to_full():
    """multi-GPU partition on one device: the slices are disjoint, cover everything and add up to the full result"""
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    full = make_context(n, box, True)
    full.setPositions(p)
    e_full = full.getPotentialEnergy()
    f_full = full.getForces()
    st_full = full.engine.stats()
    world = 3
    e_sum, f_sum, np_sum, nt_sum, flops = 0.0, torch.zeros_like(f_full), 0, 0, []
    for r in range(world):
        ctx = make_context(n, box, True, rank=r, world=world)
        ctx.setPositions(p)
        e_sum += ctx.getPotentialEnergy()
        f_sum += ctx.getForces()
        st = ctx.engine.stats()
        assert (st["n_pairs"], st["n_triplets"]) == (st_full["n_pairs"], st_full["n_triplets"])
        np_sum += st["n_pairs_owned"]
        nt_sum += st["n_triplets_owned"]
        flops.append(st["n_pairs_owned"] * 30601 + st["n_triplets_owned"] * 111818)
    assert (np_sum, nt_sum) == (st_full["n_pairs"], st_full["n_triplets"])
    assert abs(e_sum - e_full) <= 1e-11 * abs(e_full)
    assert (f_sum - f_full).abs().max().item() < 1e-8
    assert max(flops) / (sum(flops) / world) < 1.02                  # balanced by the 2B+3B work model


@pytest.mark.gpu
@pytest.mark.parametrize("pbc", [True, False])
def test_sliced_list_build_matches_full(pbc):
    """From the second evaluation on, a partitioned context builds neighbour rows only for its owner slice and the
    0.45 nm halo below it (TupleLists::build, sliced).  On moved positions the slices' lists must concatenate, in rank
    order, to exactly the full canonical lists, and energies / forces must add up to the full result."""
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    rng = np.random.default_rng(77)
    moved = pos + rng.normal(0.0, 0.01, (n, 1, 3))                   # rigid displacement of every molecule
    p0 = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    p1 = torch.as_tensor(moved.reshape(-1, 3)).cuda()
    full = make_context(n, box, pbc)
    full.setPositions(p1)
    e_full, f_full = full.getPotentialEnergy(), full.getForces()
    pairs_full, trips_full = full.getPairs(), full.getTriplets()
    world = 3
    e_sum, f_sum, pairs, trips = 0.0, torch.zeros_like(f_full), [], []
    for r in range(world):
        ctx = make_context(n, box, pbc, rank=r, world=world)
        ctx.setPositions(p0)
        ctx.getPotentialEnergy()                                      # full build: fixes this rank's owner range
        ctx.setPositions(p1)
        e_sum += ctx.getPotentialEnergy()                             # sliced build on the moved positions
        f_sum += ctx.getForces()
        st = ctx.engine.stats()
        pr, tr = ctx.getPairs(), ctx.getTriplets()
        assert (len(pr), len(tr)) == (st["n_pairs_owned"], st["n_triplets_owned"])
        assert len(pr) < 0.5 * len(pairs_full) and len(tr) < 0.5 * len(trips_full)      # really only the slice
        pairs.append(pr)
        trips.append(tr)
    assert np.array_equal(np.concatenate(pairs), pairs_full)
    assert np.array_equal(np.concatenate(trips), trips_full)
    assert abs(e_sum - e_full) <= 1e-11 * abs(e_full)
    assert (f_sum - f_full).abs().max().item() < 1e-8


@pytest.mark.gpu
def test_sliced_list_build_with_more_ranks_than_work():
    """14 waters on 6 ranks: some owner ranges are empty or own no tuples; full and sliced evaluations still add up"""
    import torch
    g = load_golden("water14_xml")
    pos = g["pos_mol"]
    n = len(pos)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    full = make_context(n, None, False)
    full.setPositions(p)
    e_full, f_full = full.getPotentialEnergy(), full.getForces()
    world = 6
    for sliced in (False, True):
        e_sum, f_sum, nt = 0.0, torch.zeros_like(f_full), 0
        for r in range(world):
            ctx = make_context(n, None, False, rank=r, world=world)
            ctx.setPositions(p)
            ctx.getPotentialEnergy()
            if sliced:
                ctx.setPositions(p)
                ctx.getPotentialEnergy()
                nt += len(ctx.getTriplets())
            e_sum += ctx.getPotentialEnergy()
            f_sum += ctx.getForces()
        assert abs(e_sum - e_full) <= 1e-11 * abs(e_full)
        assert (f_sum - f_full).abs().max().item() < 1e-8
        if sliced:
            assert nt == len(full.getTriplets())


@pytest.mark.gpu
def test_sliced_list_build_refresh_and_opt_out(monkeypatch):
    """MBPOL_REPARTITION_EVERY=k: a full build (new balanced owner ranges) every k sliced ones; 0 disables slicing"""
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
    full = make_context(n, box, True)
    full.setPositions(p)
    full.getPotentialEnergy()
    n_full = len(full.getTriplets())
    for every, expect_sliced in (("2", [False, True, True, False, True]), ("0", [False] * 4)):
        monkeypatch.setenv("MBPOL_REPARTITION_EVERY", every)
        ctx = make_context(n, box, True, rank=1, world=2)
        seen = []
        for _ in expect_sliced:
            ctx.setPositions(p)
            ctx.getPotentialEnergy()
            seen.append(len(ctx.getTriplets()) < n_full)
        assert seen == expect_sliced


def test_error_behaviour():
    from mbpol_openmm_plugin_b200 import MBPolError, _lib
    g = load_golden("water256_bulk")
    n = len(g["pos_mol"])
    ctx = make_context(n, g["box"], True)
    with pytest.raises(MBPolError):
        ctx.setPositions(np.zeros((5, 3)))                           # wrong number of positions
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    ctx.setPeriodicBoxVectors([1.2, 1.2, 1.2])                       # box shrank below 2 x cutoff
    with pytest.raises(MBPolError) as e:                             # MBPolReferenceKernels.cpp:384-391
        ctx.getPotentialEnergy()
    assert e.value.code == _lib.EBOX and "less than twice the cutoff" in e.value.message
    ctx.setPeriodicBoxVectors(g["box"].tolist())
    ctx.getPotentialEnergy()
    check_case(ctx, g)                                               # the context recovers
    with pytest.raises(MBPolError) as e:                             # MBPolReferenceKernels.cpp:402-403
        ctx.engine.update_particles(np.arange(6, dtype=np.int32).reshape(2, 3))
    assert e.value.code == _lib.ECHANGED
    with pytest.raises(MBPolError):
        ctx.engine.eval_host(g["pos_mol"].reshape(-1, 3), None, 0)   # neither 2B nor 3B requested


def test_force_groups_and_parameter_update():
    g = load_golden("water14_xml")
    n = len(g["pos_mol"])
    ctx = make_context(n, None, False, 0.65, 0.45)
    ctx.forces[1].setForceGroup(3)
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    e2 = ctx.getPotentialEnergy(groups=1 << 0)
    f2 = ctx.getForces().copy()
    e3 = ctx.getPotentialEnergy(groups=1 << 3)
    f3 = ctx.getForces().copy()
    assert abs(e2 - g["e2"]) <= E_RTOL * abs(g["e2"]) and abs(e3 - g["e3"]) <= E_RTOL * abs(g["e3"])
    assert np.abs(f2 - g["f2"]).max() < F_ATOL and np.abs(f3 - g["f3"]).max() < F_ATOL
    assert ctx.getPotentialEnergy(groups=1 << 5) == 0.0
    # updateParametersInContext: swap the hydrogens of every molecule -> the PIPs are permutationally invariant
    for f in ctx.forces:
        for m in range(n):
            o, h1, h2 = f.getParticleParameters(m)
            f.setParticleParameters(m, [o, h2, h1])
        f.updateParametersInContext(ctx)
    e = ctx.getPotentialEnergy()
    assert abs(e - (g["e2"] + g["e3"])) <= 1e-9 * abs(g["e2"] + g["e3"])


def test_openmm_fixed_point_force_buffer():
    """mbpol_add_forces_fixed_point: OpenMM CUDA convention, int64 = f * 2^32, SoA with padding
    (platforms/cuda/src/kernels/twobodyForce.cu:442-457)"""
    import torch
    from mbpol_openmm_plugin_b200 import _lib
    g = load_golden("water14_xml")
    n = len(g["pos_mol"])
    ctx = make_context(n, None, False, 0.65, 0.45)
    ctx.setPositions(torch.as_tensor(g["pos_mol"].reshape(-1, 3)).cuda())
    ctx.getPotentialEnergy()
    padded = 64
    buf = torch.zeros(3 * padded, dtype=torch.int64, device="cuda")
    L = _lib.load()
    _lib.check(L.mbpol_add_forces_fixed_point(ctx.engine.h, buf.data_ptr(), padded, None), ctx.engine.h)
    torch.cuda.synchronize()
    f = (buf.cpu().numpy().reshape(3, padded)[:, :3 * n].T / 2.0 ** 32)
    assert np.abs(f - (g["f2"] + g["f3"])).max() < 1e-6


def test_nocutoff_method_is_infinite_nonperiodic_cutoff():
    from mbpol_openmm_plugin_b200 import MBPolContext, MBPolThreeBodyForce, MBPolTwoBodyForce
    g = load_golden("water14_nocut")
    n = len(g["pos_mol"])
    fs = []
    for cls in (MBPolTwoBodyForce, MBPolThreeBodyForce):
        f = cls()
        f.setNonbondedMethod(cls.NoCutoff)
        f.setCutoff(0.3)                         # ignored for NoCutoff
        for m in range(n):
            f.addParticle([3 * m, 3 * m + 1, 3 * m + 2])
        fs.append(f)
    ctx = MBPolContext(3 * n, fs)
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    ctx.getPotentialEnergy()
    check_case(ctx, g)
    assert len(ctx.getTriplets()) == 364


def test_rectangular_box_and_unwrapped_positions(checker):
    """non-cubic box (L, L, 2L) with molecules scattered over several periodic images: same lists / energies / forces"""
    from oracle.refbind import canonical
    g = load_golden("water512_cell")
    L = float(g["box"][0])
    rng = np.random.default_rng(5)
    pos = np.concatenate([g["pos_mol"], g["pos_mol"] + np.array([0, 0, L])])
    pos = pos + rng.normal(0, 0.004, size=(len(pos), 1, 3))
    box = np.array([L, L, 2 * L])
    pos = pos + (rng.integers(-2, 3, size=(len(pos), 1, 3)) * box)        # whole molecules moved to other images
    n = len(pos)
    ctx = make_context(n, box, True)
    ctx.setPositions(pos.reshape(-1, 3))
    ctx.getPotentialEnergy()
    f = ctx.getForces()
    if checker.kind != "reference":
        pytest.skip("needs the reference build (the C port list build is O(n^2))")
    posO = pos[:, 0].copy()
    pairs = canonical(checker.pairs(posO, box, True, 0.65))
    trips = canonical(checker.triplets(posO, box, True, 0.45))
    assert np.array_equal(ctx.getPairs(), pairs) and np.array_equal(ctx.getTriplets(), trips)
    idx = np.arange(3 * n).reshape(n, 3)
    e2, f2 = checker.eval_2b(pos.reshape(-1, 3), idx, pairs, box, True, 0.65)
    e3, f3 = checker.eval_3b(pos.reshape(-1, 3), idx, trips, box, True, 0.45)
    assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - e2) <= E_RTOL * abs(e2)
    assert abs(ctx.energies["CalcMBPolThreeBodyForce"] - e3) <= E_RTOL * abs(e3)
    assert np.abs(f - (f2 + f3)).max() < F_ATOL


def test_row_builder_fast_path_and_fallbacks(checker, monkeypatch):
    """lists.cu builds neighbour rows with a one-pass warp kernel into fixed-stride rows and falls back to the two-pass
    CSR builder when a row overflows: both paths (and the forced fallback) must give the reference's lists"""
    from oracle.refbind import canonical
    g = load_golden("water256_bulk")
    n = len(g["pos_mol"])
    pos = g["pos_mol"].reshape(-1, 3)
    # (1) forced two-pass on the periodic fixture
    monkeypatch.setenv("MBPOL_FORCE_TWO_PASS", "1")
    ctx = make_context(n, g["box"], True)
    ctx.setPositions(pos)
    ctx.getPotentialEnergy()
    assert np.array_equal(ctx.getPairs(), g["pairs"].reshape(-1, 2)) and np.array_equal(ctx.getTriplets(), g["triplets"].reshape(-1, 3))
    check_case(ctx, g)
    monkeypatch.delenv("MBPOL_FORCE_TWO_PASS")
    # (2) overflow: non-periodic, infinite 2B cutoff -> rows of up to 255 lower neighbours do not fit the fixed stride
    ctx = make_context(n, None, False, 1e10, 0.45)
    ctx.setPositions(pos)
    ctx.getPotentialEnergy()
    pairs = ctx.getPairs()
    assert len(pairs) == n * (n - 1) // 2
    posO = g["pos_mol"][:, 0].copy()
    trips = canonical(checker.triplets(posO, None, False, 0.45))
    assert np.array_equal(ctx.getTriplets(), trips)
    idx = np.arange(3 * n).reshape(n, 3)
    e2, f2 = checker.eval_2b(pos, idx, pairs, None, False)[:2]
    e3, f3 = checker.eval_3b(pos, idx, trips, None, False)[:2]
    assert abs(ctx.energies["CalcMBPolTwoBodyForce"] - e2) <= E_RTOL * abs(e2)
    assert abs(ctx.energies["CalcMBPolThreeBodyForce"] - e3) <= E_RTOL * abs(e3)
    assert np.abs(ctx.getForces() - (f2 + f3)).max() < F_ATOL


def test_host_force_output_is_overwritten_not_accumulated():
    """MBPolContext hands a preallocated (e.g. pinned) output array to mbpol_eval_host with MBPOL_OVERWRITE_FORCES: stale
    contents must not leak into the result; the raw C-ABI call without the flag still ADDS (reference semantics)."""
    import ctypes
    from mbpol_openmm_plugin_b200 import _lib
    g = load_golden("water14_xml")
    n = len(g["pos_mol"])
    ctx = make_context(n, g["box"], False)
    out = np.full((3 * n, 3), np.nan)
    ctx.setForceOutput(out)
    ctx.setPositions(g["pos_mol"].reshape(-1, 3))
    ctx.getPotentialEnergy()
    assert ctx.getForces() is out
    assert np.abs(out - (g["f2"] + g["f3"])).max() < F_ATOL
    pos = np.ascontiguousarray(g["pos_mol"].reshape(-1, 3))
    acc = np.ones((3 * n, 3))
    e = (ctypes.c_double * 2)()
    _lib.check(_lib.load().mbpol_eval_host(ctx.engine.h, pos.ctypes.data, acc.ctypes.data, e, _lib.DO_2B | _lib.DO_3B),
               ctx.engine.h)
    assert np.abs(acc - 1.0 - (g["f2"] + g["f3"])).max() < F_ATOL


def test_full_size_262144_properties(checker):
    """BASELINE config 5 (the bench workload: 262,144 waters, 4.8 M dimers, 7.1 M trimers): structural list invariants
    on the full lists, list membership AND completeness against brute force for a sample of owner molecules, per-tuple
    energies / forces of a random tuple sample, and the WHOLE box -- total E2B, E3B (1e-9) and the full force array
    (1e-6 kJ/mol/nm) -- against the oracle fanned out over the host cores (about 330 core-seconds per evaluation), on
    the original coordinates and on a translated image of them."""
    import torch
    from mbpol_openmm_plugin_b200 import eval_tuples
    from mbpol_openmm_plugin_b200.systems import particle_indices, synthetic_box
    pos, box = synthetic_box(8)
    n = len(pos)
    assert n == 262144
    ctx = make_context(n, box, True)
    ctx.setPositions(torch.as_tensor(pos.reshape(-1, 3)).cuda())
    e = ctx.getPotentialEnergy()
    f = ctx.getForces().cpu().numpy()
    pairs, trips = ctx.getPairs(), ctx.getTriplets()
    st = ctx.engine.stats()
    assert st["n_pairs"] == len(pairs) and st["n_triplets"] == len(trips)
    assert 17.5 * n < len(pairs) < 19.5 * n and 23 * n < len(trips) < 28 * n
    # canonical order, strict I > J (> K), no duplicates
    assert pairs.min() >= 0 and pairs.max() < n and trips.min() >= 0 and trips.max() < n
    assert np.all(pairs[:, 0] > pairs[:, 1]) and np.all(trips[:, 0] > trips[:, 1]) and np.all(trips[:, 1] > trips[:, 2])
    assert np.all(np.diff(pairs[:, 0].astype(np.int64) * n + pairs[:, 1]) > 0)
    key = (trips[:, 0].astype(np.int64) * n + trips[:, 1]) * n + trips[:, 2]
    assert np.all(np.diff(key) > 0)
    del key
    # membership + completeness for sampled owners: the reference's minimum-image distance (periodicDifference,
    # ReferenceThreeNeighborList.cpp:14-33) by brute force over ALL lower-index oxygens
    posO = pos[:, 0]
    rng = np.random.default_rng(5)

    def lower_neighbours(i, cut):
        d = posO[:i] - posO[i]
        d -= box * np.floor(d / box + 0.5)
        return set(np.nonzero((d * d).sum(axis=1) <= cut * cut)[0].tolist())

    p_start = np.searchsorted(pairs[:, 0], np.arange(n + 1))
    t_start = np.searchsorted(trips[:, 0], np.arange(n + 1))
    for i in rng.choice(n, 60, replace=False).tolist() + [0, 1, n - 1]:
        assert set(pairs[p_start[i]:p_start[i + 1], 1].tolist()) == lower_neighbours(i, 0.65)
        near = lower_neighbours(i, 0.45)
        want = {(j, k) for j in near for k in lower_neighbours(j, 0.45)}          # chain set, SURVEY.md App. B
        assert {tuple(t) for t in trips[t_start[i]:t_start[i + 1], 1:].tolist()} == want
    # per-tuple values of a random sample against the oracle
    idx = particle_indices(n)
    for order, tuples in ((2, pairs), (3, trips)):
        sel = np.sort(rng.choice(len(tuples), 2000, replace=False))
        eg, fg = eval_tuples(order, pos, tuples[sel], box)
        if checker.kind == "reference":
            er, fr = (checker.each_2b if order == 2 else checker.each_3b)(pos, tuples[sel], box, True)
        else:
            _, _, er, fr = (checker.eval_2b if order == 2 else checker.eval_3b)(pos.reshape(-1, 3), idx, tuples[sel], box,
                                                                                 True, each=True)
        assert np.abs(eg - er).max() <= 1e-9 * np.abs(er).max() and np.abs(fg - fr).max() < F_ATOL
    # forces on the molecules of one small region against the oracle: every tuple that touches them
    # (energy sum of those tuples is not separable, forces on the chosen molecules are)
    probe = rng.choice(n, 3, replace=False)
    for m in probe.tolist():
        tp = pairs[(pairs[:, 0] == m) | (pairs[:, 1] == m)]
        tt = trips[(trips[:, 0] == m) | (trips[:, 1] == m) | (trips[:, 2] == m)]
        _, f2 = checker.eval_2b(pos.reshape(-1, 3), idx, tp, box, True)[:2]
        _, f3 = checker.eval_3b(pos.reshape(-1, 3), idx, tt, box, True)[:2]
        assert np.abs(f[3 * m:3 * m + 3] - (f2 + f3)[3 * m:3 * m + 3]).max() < F_ATOL
    # the whole box against the oracle: every dimer and trimer, total energies and all 786,432 x 3 force components
    from oracle.refbind import eval_parallel
    e2_gpu, e3_gpu = ctx.energies["CalcMBPolTwoBodyForce"], ctx.energies["CalcMBPolThreeBodyForce"]
    e2, f2 = eval_parallel(checker, 2, pos.reshape(-1, 3), idx, pairs, box, True, chunk=40000)
    e3, f3 = eval_parallel(checker, 3, pos.reshape(-1, 3), idx, trips, box, True, chunk=40000)
    assert abs(e2_gpu - e2) <= E_RTOL * abs(e2) and abs(e3_gpu - e3) <= E_RTOL * abs(e3)
    err = float(np.abs(f - (f2 + f3)).max())
    print(f"262,144 waters, whole box vs {checker.kind} oracle: dE2/E2 = {abs(e2_gpu - e2) / abs(e2):.2e}, "
          f"dE3/E3 = {abs(e3_gpu - e3) / abs(e3):.2e}, max |dF| = {err:.2e} kJ/mol/nm")
    assert err < F_ATOL
    # Newton's third law over 786,432 atoms; invariance under whole-box translations of molecules
    assert np.abs(f.sum(axis=0)).max() < 1e-5
    shifted = pos.copy()
    shifted[::7] += np.array([box[0], 0, 0])
    shifted[::5] -= np.array([0, 2 * box[1], box[2]])
    ctx.setPositions(torch.as_tensor(shifted.reshape(-1, 3)).cuda())
    e_s = ctx.getPotentialEnergy()
    assert abs(e_s - e) <= 1e-9 * abs(e)
    assert np.array_equal(ctx.getPairs(), pairs)
    f_s = ctx.getForces().cpu().numpy()
    # translated image: the oracle on the SAME translated coordinates for the molecules that moved the farthest, and
    # every tuple touching them (a second whole-box oracle run would double the test time)
    moved = np.arange(0, n, 35)[:64]                                  # shifted by both translations
    touch2 = np.isin(pairs, moved).any(axis=1)
    touch3 = np.isin(trips, moved).any(axis=1)
    _, f2s = eval_parallel(checker, 2, shifted.reshape(-1, 3), idx, pairs[touch2], box, True)
    _, f3s = eval_parallel(checker, 3, shifted.reshape(-1, 3), idx, trips[touch3], box, True)
    rows = (3 * moved[:, None] + np.arange(3)).ravel()
    assert np.abs(f_s[rows] - (f2s + f3s)[rows]).max() < F_ATOL


def test_verlet_skin_lists_equal_the_rebuilt_lists_at_every_step():
    """mbpol_set_skin: neighbour rows built out to cutoff + skin are kept while no oxygen moved more than skin / 2 and
    filtered with the exact distance test at every evaluation -> the tuple lists (hence energies and forces) must be
    IDENTICAL to those of a context that rebuilds everything on every call, at every step of a random walk; a large
    displacement must trigger a rebuild."""
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    a, b = make_context(n, box, True), make_context(n, box, True)
    b.engine.set_skin(0.1)
    rng = np.random.default_rng(3)
    p = pos.reshape(-1, 3).copy()
    reused = 0
    for step in range(12):
        if step == 8:
            p[rng.choice(len(p) // 3) * 3:][:3] += 0.3               # one molecule jumps 0.3 nm: beyond skin / 2
        elif step:
            p += np.repeat(rng.normal(scale=0.004, size=(n, 3)), 3, axis=0)      # rigid per-molecule moves, 0.004 nm rms
        t = torch.as_tensor(p).cuda()
        a.setPositions(t)
        b.setPositions(t)
        ea, eb = a.getPotentialEnergy(), b.getPotentialEnergy()
        assert np.array_equal(a.getPairs(), b.getPairs()) and np.array_equal(a.getTriplets(), b.getTriplets())
        assert ea == eb and torch.equal(a.getForces(), b.getForces())            # same lists, same kernels: bitwise
        st = b.engine.stats()
        reused += st["lists_reused"]
        if step == 8:
            assert not st["lists_reused"]
    sa, sb = a.engine.stats(), b.engine.stats()
    assert sa["n_row_builds"] == 12 and not sa["lists_reused"]
    assert reused >= 8 and sb["n_row_builds"] <= 4


def test_sliced_build_that_falls_back_on_some_ranks_keeps_the_owner_ranges():
    """A sliced build falls back to full lists on exactly the ranks whose rows overflow the fixed stride (here: a 1.3 nm
    2B cutoff, ~300 neighbours; the high-index ranks overflow, rank 0 with few lower-index neighbours may not).  The owner
    ranges and the refresh schedule must not move on those ranks alone: at every step of a random walk the partial results
    of the ranks must still add up to the unpartitioned result (tuples neither dropped nor counted twice)."""
    import torch
    from mbpol_openmm_plugin_b200.systems import synthetic_box
    pos, box = synthetic_box(2)
    n = len(pos)
    world = 3
    full = make_context(n, box, True, cut2=1.3)
    ranks = [make_context(n, box, True, cut2=1.3, rank=r, world=world) for r in range(world)]
    rng = np.random.default_rng(9)
    p = pos.reshape(-1, 3).copy()
    sliced_seen = fallback_seen = 0
    for step in range(5):
        if step:
            p += np.repeat(rng.normal(scale=0.01, size=(n, 3)), 3, axis=0)
        t = torch.as_tensor(p).cuda()
        full.setPositions(t)
        e_full = full.getPotentialEnergy()
        f_full = full.getForces()
        e_sum, f_sum, owned2, owned3 = 0.0, torch.zeros_like(f_full), 0, 0
        for ctx in ranks:
            ctx.setPositions(t)
            e_sum += ctx.getPotentialEnergy()
            f_sum += ctx.getForces()
            st = ctx.engine.stats()
            owned2 += st["n_pairs_owned"]
            owned3 += st["n_triplets_owned"]
            if step:
                sliced_seen += st["lists_sliced"]
                fallback_seen += not st["lists_sliced"]
        stf = full.engine.stats()
        assert (owned2, owned3) == (stf["n_pairs"], stf["n_triplets"])             # a partition of the tuples, every step
        assert abs(e_sum - e_full) <= 1e-10 * abs(e_full)
        assert (f_sum - f_full).abs().max().item() < 1e-8
    assert fallback_seen > 0                                     # the situation of the test did occur
    assert all(c.engine.stats()["n_repartitions"] == 1 for c in ranks)          # nobody rebalanced on its own


def test_distributed_entry_points_on_a_single_rank_communicator():
    """mbpol_eval_distributed and mbpol_eval_distributed_host (device / host buffers, NCCL broadcast + reduce, the sliced
    reduce with overlapped copy-back) on a communicator of ONE rank must reproduce mbpol_eval bit for bit -- the
    collectives are identities there, everything else is the code path the N-GPU runs take (bench.py checks N > 1)."""
    import ctypes as C
    import torch
    from mbpol_openmm_plugin_b200 import _lib
    g = load_golden("water256_bulk")
    n = len(g["pos_mol"])
    ctx = make_context(n, g["box"], True)
    L = _lib.load()
    h = ctx.engine.h
    pos = torch.as_tensor(g["pos_mol"].reshape(-1, 3)).cuda()
    f0 = torch.zeros_like(pos)
    e0 = ctx.engine.eval_device(pos, f0, 3)
    ident = C.create_string_buffer(_lib.COMM_ID_BYTES)
    _lib.check(L.mbpol_comm_unique_id(ident))
    _lib.check(L.mbpol_comm_init_rank(h, ident.raw, 0, 1), h)
    try:
        f1 = torch.full_like(pos, 7.0)                       # overwritten, not accumulated into
        e1 = (C.c_double * 2)()
        _lib.check(L.mbpol_eval_distributed(h, pos.data_ptr(), f1.data_ptr(), e1, 3, None), h)
        torch.cuda.synchronize()
        assert (e1[0], e1[1]) == tuple(e0) and torch.equal(f1, f0)
        ph = torch.as_tensor(g["pos_mol"].reshape(-1, 3)).pin_memory()
        fh = torch.full((3 * n, 3), -1.0, dtype=torch.float64).pin_memory()
        e2 = (C.c_double * 2)()
        _lib.check(L.mbpol_eval_distributed_host(h, ph.data_ptr(), fh.data_ptr(), e2, 3), h)
        assert (e2[0], e2[1]) == tuple(e0) and torch.equal(fh, f0.cpu())
    finally:
        L.mbpol_comm_destroy(h)
