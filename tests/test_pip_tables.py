"""The committed polynomial data and the generated evaluator program, checked on the CPU.

    reference SLP  --tools/extract_pip.py-->  monomial tables (data/pip{2b,3b}.npz)
                   --tools/gen_interp.py -->  command stream interpreted by the sm_100a kernel
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import extract_pip  # noqa: E402
import gen_interp  # noqa: E402
import gen_msym  # noqa: E402

DATA = os.path.join(ROOT, "mbpol_openmm_plugin_b200", "data")
SPEC = {"pip2b": (31, 1153, 12725, "2"), "pip3b": (36, 1163, 33525, "3")}


@pytest.mark.parametrize("name", ["pip2b", "pip3b"])
def test_table_shape(name):
    t = dict(np.load(os.path.join(DATA, name + ".npz")))
    nv, nc, nm, _ = SPEC[name]
    assert int(t["n_vars"]) == nv and len(t["coef"]) == nc and len(t["mono_coef"]) == nm
    assert set(np.unique(t["mono_coef"]).tolist()) == set(range(nc))      # every coefficient is used
    deg = (t["mono_vars"] != 255).sum(axis=1)
    assert deg.min() >= 1 and deg.max() == 4


@pytest.mark.parametrize("name", ["pip2b", "pip3b"])
def test_table_reproduces_reference_polynomial_values(name):
    """golden values were produced by poly_2b_v6x_eval / poly_3b_v2x::eval (tests/golden/make_golden.py)"""
    t = dict(np.load(os.path.join(DATA, name + ".npz")))
    g = np.load(os.path.join(GOLDEN, "poly_values.npz"))
    k = SPEC[name][3]
    for x, e, grad in zip(g["x" + k], g["e" + k], g["g" + k]):
        E, G = extract_pip.eval_table(t, x)
        assert abs(E - e) <= 1e-12 * max(1.0, abs(e))
        assert np.abs(G - grad).max() <= 1e-12 * np.abs(grad).max()


@pytest.mark.parametrize("name,opts", [
    # the shipped configurations (mbpol_openmm_plugin_b200/build.py GEN_ARGS) and neighbouring settings
    ("pip2b", {"sym": 12, "npre": 4, "piece": 32, "block": 384}),
    ("pip3b", {"sym": 32, "npre": 8, "piece": 64}),
    ("pip2b", {"sym": 12, "npre": 4, "piece": 32, "block": 384, "dual_max": 16}),
    ("pip3b", {"sym": 32, "npre": 8, "piece": 64, "dual_max": 13}),
    ("pip2b", {"sym": 12, "npre": 4, "piece": 32, "block": 384, "dual_max": 4, "slim": True, "anchors": True}),
    ("pip2b", {"sym": 20, "npre": 4, "piece": 32, "block": 384, "dual_max": 4, "slim": True, "anchors": True, "basis": True}),
    ("pip3b", {"sym": 28, "npre": 8, "piece": 64, "dual_max": 9, "basis": True}),
    ("pip3b", {"sym": 28, "npre": 8, "piece": 64, "dual_max": 9}),     # (the 3B evaluator of round 1 / 2: kept, not shipped)
    ("pip3b", {"sym": 32, "npre": 8, "piece": 64, "dual_max": 32, "group": 2}),
    ("pip2b", {"sym": 16, "skeleton_weight": 0}), ("pip3b", {"sym": 24, "skeleton_weight": 0, "fall": 0})])
def test_command_stream_matches_table(name, opts, tmp_path):
    """python interpreter of the packed stream (the same records the GPU kernel executes) vs the monomial definition"""
    meta, words, masks, phases, dual = gen_interp.generate(name, str(tmp_path), **opts)
    t = dict(np.load(os.path.join(DATA, name + ".npz")))
    rng = np.random.default_rng(11)
    for _ in range(2):
        x = rng.uniform(0.05, 1.6, int(t["n_vars"]))
        E, G = gen_interp.interpret(name, words, x, meta["chunk_bytes"], masks, phases, dual)
        Er, Gr = extract_pip.eval_table(t, x)
        assert abs(E - Er) <= 1e-13 * max(1.0, abs(Er))
        assert np.abs(G - Gr).max() <= 1e-13 * np.abs(Gr).max()
    # stream layout invariants the kernel relies on
    cw = meta["chunk_bytes"] // 8
    assert len(words) % cw == 0 and not words[-cw:].any()           # trailing all-zero look-ahead chunk
    assert meta["smem_bytes"] <= 227 * 1024
    # every variable belongs to exactly one orbit of every phase; phase 0 is the identity
    nv = int(t["n_vars"])
    if opts.get("basis"):       # ONE phase in the symmetry-adapted basis: nv rows with distinct lowest variables, invertible
        assert len(meta["phases"]) == 1 and sorted(o[0] for o in meta["phases"][0]) == list(range(nv))
        B = np.zeros((nv, nv))
        for q, o in enumerate(meta["phases"][0]):
            B[q, o] = 1.0
        assert abs(abs(np.linalg.det(B)) - 1.0) < 1e-9
    else:
        assert meta["phases"][0] == [[j] for j in range(nv)]
        for orbs in meta["phases"]:
            assert sorted(v for o in orbs for v in o) == list(range(nv))
    assert meta["fp64_instr"] < {"pip2b": 24000, "pip3b": 67500}[name]     # vs 30,180 / 79,182 without symmetry
    if opts.get("dual_max"):
        assert meta["records"]["dual"] > 100 and meta["dual"]            # pairs of sibling leaf nodes were found and fused


@pytest.mark.parametrize("opts", [
    {"K": (24,), "npre": 16, "piece": 16, "ahead": 1, "anchors": True, "mode": "basis"},   # shipped (build.py MSYM_ARGS)
    {"K": (20,), "npre": 16, "piece": 16, "ahead": 1, "anchors": True, "mode": "single"},  # single-variable slots
    {"K": (12,), "npre": 8, "piece": 16, "ahead": 2, "anchors": False, "mode": "basis1"}]) # other orientation of the basis
def test_molecule_symmetrised_stream_matches_table(opts, tmp_path):
    """tools/gen_msym.py: P(x) = sum over the six molecule relabellings of Q(sigma x); the python interpreter of the packed
    stream (the records the 3B kernel executes, six contexts per record) vs the monomial definition of P"""
    meta, words, handlers, phases, kvars = gen_msym.generate("pip3b", str(tmp_path), **opts)
    t = dict(np.load(os.path.join(DATA, "pip3b.npz")))
    rng = np.random.default_rng(11)
    for _ in range(2):
        x = rng.uniform(0.05, 1.6, 36)
        E, G = gen_msym.interpret("pip3b", words, x, meta["chunk_bytes"], handlers, phases, meta["block"], kvars)
        Er, Gr = extract_pip.eval_table(t, x)
        assert abs(E - Er) <= 1e-13 * max(1.0, abs(Er))
        assert np.abs(G - Gr).max() <= 1e-13 * np.abs(Gr).max()
    cw = meta["chunk_bytes"] // 8
    assert len(words) % cw == 0 and not words[-cw:].any()           # trailing all-zero look-ahead chunk
    assert meta["smem_bytes"] <= 227 * 1024
    assert meta["contexts"] == [6]
    assert meta["records"]["leaf"] < 400                            # 2,050 leaf nodes before the symmetrisation
    if opts.get("mode") == "basis":
        assert meta["fp64_instr"] < 70000                           # 78,528 with single-variable slots


@pytest.mark.parametrize("name", ["pip2b", "pip3b"])
def test_polynomial_is_invariant_under_molecule_relabelling(name):
    """the premise of tools/gen_msym.py: relabelling the molecules permutes the variables and leaves the polynomial unchanged"""
    t = dict(np.load(os.path.join(DATA, name + ".npz")))
    x = np.random.default_rng(8).uniform(0.1, 1.5, int(t["n_vars"]))
    E, G = extract_pip.eval_table(t, x)
    perms = list(gen_msym.mol_perms(name).values())
    assert len(perms) == (6 if name == "pip3b" else 2) and perms[0] == list(range(len(x)))
    for pm in perms[1:]:
        pm = np.asarray(pm)
        assert sorted(pm.tolist()) == list(range(len(pm)))
        Ep, Gp = extract_pip.eval_table(t, x[pm])
        assert abs(Ep - E) <= 1e-13 * abs(E)
        assert np.abs(Gp - G[pm]).max() <= 1e-13 * np.abs(G).max()


def test_shipped_generator_settings_are_tested():
    """whatever build.py ships must be one of the configurations the stream tests above interpret"""
    from mbpol_openmm_plugin_b200 import build
    a = build.GEN_ARGS
    assert (a["sym2"], a["block2"], a["piece2"], a["npre2"], a["dual2"], a["slim2"], a["anchors2"], a["basis2"]) == \
        (20, 384, 32, 4, 4, 1, 1, 1)
    assert build.MSYM_ARGS == {"mode": "basis", "k0": 24, "piece": 16, "npre": 16, "ahead": 1, "anchors": 1}


@pytest.mark.parametrize("name", ["pip2b", "pip3b"])
def test_polynomial_is_invariant_under_the_site_swaps(name):
    """the premise of the symmetry phases: exchanging the two hydrogens (or lone-pair sites) of one molecule permutes
    the variables and leaves the polynomial unchanged"""
    t = dict(np.load(os.path.join(DATA, name + ".npz")))
    x = np.random.default_rng(7).uniform(0.1, 1.5, int(t["n_vars"]))
    E, G = extract_pip.eval_table(t, x)
    for pm in gen_interp.swap_perms(name):
        pm = np.asarray(pm)
        assert sorted(pm.tolist()) == list(range(len(pm))) and np.array_equal(pm[pm], np.arange(len(pm)))
        Ep, Gp = extract_pip.eval_table(t, x[pm])
        assert abs(Ep - E) <= 1e-13 * abs(E) and np.abs(Gp - G[pm]).max() <= 1e-13 * np.abs(G).max()


def test_port_polynomial_matches_table(port):
    for order, name in ((2, "pip2b"), (3, "pip3b")):
        t = dict(np.load(os.path.join(DATA, name + ".npz")))
        x = np.random.default_rng(5).uniform(0.1, 1.4, int(t["n_vars"]))
        E, G = port.poly(order, x)
        Er, Gr = extract_pip.eval_table(t, x)
        assert abs(E - Er) <= 1e-12 * abs(Er) and np.abs(G - Gr).max() <= 1e-12 * np.abs(Gr).max()


@pytest.mark.parametrize("name,macro,sites", [
    ("pip2b", "MBPOL_2B_VARS", ["Oa", "Ha1", "Ha2", "Ob", "Hb1", "Hb2", "Xa1", "Xa2", "Xb1", "Xb2"]),
    ("pip3b", "MBPOL_3B_VARS", ["Oa", "Ha1", "Ha2", "Ob", "Hb1", "Hb2", "Oc", "Hc1", "Hc2"])])
def test_generator_and_cuda_kernels_agree_on_the_variable_table(name, macro, sites):
    """the symmetry phases of the generator (tools/gen_interp.py::_sites) and the variable kernels (csrc/tuples.cu) must
    mean the same site pair by variable i"""
    import re
    src = open(os.path.join(ROOT, "mbpol_openmm_plugin_b200", "csrc", "tuples.cu")).read()
    body = src[src.index("#define " + macro):]
    body = body[:body.index("\n\n")]
    rows = re.findall(r"V\((\d+),(?:\s*\d+,)?\s*\w+,\s*\w+,\s*(\d+),\s*(\d+)\)", body)
    table = {int(i): frozenset((sites[int(p)], sites[int(q)])) for i, p, q in rows}
    pairs, _ = gen_interp._sites(name)
    assert sorted(table) == list(range(len(pairs)))
    assert [table[i] for i in range(len(pairs))] == [frozenset(p) for p in pairs]


@pytest.mark.parametrize("name", ["pip2b", "pip3b"])
def test_shipped_cubins_dispatch_through_the_uniform_datapath(name):
    """ptxas keeps the opcode of the command stream either in uniform registers (dispatch = BRXU, jump-table read hoisted)
    or in vector registers (BRX R: 8 % slower, `branch_resolving` stalls) -- a global decision that flips with small changes
    of the generated PTX (tools/gen_interp.py, tail()).  The shipped kernels must be in the first regime, with no spills."""
    import shutil
    import subprocess
    gen = os.path.join(ROOT, "mbpol_openmm_plugin_b200", "csrc", "gen")
    cubin = os.path.join(gen, name + ".cubin")
    if not os.path.exists(cubin) or shutil.which("cuobjdump") is None:
        pytest.skip("library not built here")
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    assert sass.count("BRXU") >= (40 if name == "pip2b" else 20) and sass.count("BRX R") <= 2
    assert "DFMA" in sass and "LDGSTS" in sass                   # FP64 FMAs, cp.async stream staging
    log = open(os.path.join(gen, name + ".ptxas.log")).read()
    assert "0 bytes spill stores" in log
