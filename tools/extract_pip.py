"""Extract the MB-pol 2B/3B permutationally invariant polynomials into *monomial tables*.

Needs /root/reference (run in the authoring container); the result is committed as
``mbpol_openmm_plugin_b200/data/pip{2b,3b}.npz`` so that nothing downstream (code generator, oracle port,
CUDA build, GPU box) ever reads the reference tree.

The reference stores each polynomial as a Maple-optimised straight-line program (poly-2b-v6x.cpp:3-13841,
poly-3b-v2x.cpp:3-28378).  We do NOT keep that program: it is expanded symbolically into the mathematical object it
encodes,
        E(x) = sum_m  a[c_m] * prod_{v in V_m} x_v          (degree <= 4, every multiplier exactly 1)
i.e. a list of monomials (c_m, V_m).  2B: 12,725 monomials / 1,153 coefficients / 31 variables; 3B: 33,525 / 1,163 / 36.
Coefficient values come from mbpol_2body_constants.h:54-1208 and mbpol_3body_constants.h:22-1186.

npz fields: coef f64[n_coef]; mono_coef u16[n_mono]; mono_vars u8[n_mono,4] (sorted, padded with 255); n_vars.
"""
from __future__ import annotations

import collections
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slp  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mbpol_openmm_plugin_b200", "data")
PAD = 255


def expand(dag, root):
    """Expand node ``root`` into {(coef_index | -1, sorted var tuple): multiplier}; E is linear in a[]."""
    sys.setrecursionlimit(200000)
    memo = {}

    def ex(n):
        r = memo.get(n)
        if r is not None:
            return r
        k = dag.kind[n]
        if k == "a":
            r = {(dag.args[n], ()): 1.0}
        elif k == "x":
            r = {(-1, (dag.args[n],)): 1.0}
        elif k == "c":
            r = {(-1, ()): dag.args[n]}
        elif k == "add":
            acc = collections.defaultdict(float)
            for c in dag.args[n]:
                for key, v in ex(c).items():
                    acc[key] += v
            r = dict(acc)
        else:
            r = None
            for c in dag.args[n]:
                p = ex(c)
                if r is None:
                    r = p
                    continue
                q = collections.defaultdict(float)
                for (a1, m1), v1 in r.items():
                    for (a2, m2), v2 in p.items():
                        if a1 != -1 and a2 != -1:
                            raise ValueError("polynomial is not linear in the coefficients")
                        q[(max(a1, a2), tuple(sorted(m1 + m2)))] += v1 * v2
                r = dict(q)
        memo[n] = r
        return r

    return ex(root)


def read_coefficients(header):
    """The ``thefit[]`` initialiser: one value per line followed by ``// index``."""
    txt = open(header).read()
    body = txt[txt.index("thefit[]"):]
    body = body[body.index("{") + 1: body.index("};")]
    vals = [float(m.group(1)) for m in re.finditer(r"([-+]?\d\.\d+e[-+]\d+)\s*,?\s*//\s*\d+", body)]
    return np.array(vals, dtype=np.float64)


def build(poly_src, coef_header):
    dag = slp.parse(poly_src)
    E = expand(dag, dag.outputs["E"])
    rows = []
    for (a, mono), mult in E.items():
        if a < 0 or mult != 1.0 or len(mono) > 4:
            raise ValueError(f"unexpected term {(a, mono, mult)}")
        rows.append((mono, a))
    rows.sort(key=lambda r: (len(r[0]), r[0]))
    mv = np.full((len(rows), 4), PAD, dtype=np.uint8)
    mc = np.zeros(len(rows), dtype=np.uint16)
    for i, (mono, a) in enumerate(rows):
        mv[i, :len(mono)] = mono
        mc[i] = a
    coef = read_coefficients(coef_header)
    assert len(coef) == dag.n_a, (len(coef), dag.n_a)
    return dict(coef=coef, mono_coef=mc, mono_vars=mv, n_vars=np.int32(dag.n_x)), dag


def eval_table(tab, x):
    """numpy evaluation of E and dE/dx from a monomial table (slow, exact definition)."""
    xv = np.concatenate([np.asarray(x, dtype=np.float64), np.ones(256 - len(x))])  # x[255] = 1 (padding)
    mv = tab["mono_vars"].astype(np.int64)
    c = tab["coef"][tab["mono_coef"]]
    f = xv[mv]                                   # n_mono x 4
    E = float(np.sum(c * f.prod(axis=1)))
    g = np.zeros(int(tab["n_vars"]))
    for s in range(4):
        others = np.delete(f, s, axis=1).prod(axis=1) * c
        ok = mv[:, s] != PAD
        np.add.at(g, mv[ok, s], others[ok])
    return E, g


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, src, hdr in (("pip2b", slp.POLY2B, f"{slp.REF_SRC}/mbpol_2body_constants.h"),
                           ("pip3b", slp.POLY3B, f"{slp.REF_SRC}/mbpol_3body_constants.h")):
        tab, dag = build(src, hdr)
        out = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(out, **tab)
        deg = collections.Counter(int((r != PAD).sum()) for r in tab["mono_vars"])
        print(f"{name}: {len(tab['mono_coef'])} monomials, {len(tab['coef'])} coefficients, "
              f"{int(tab['n_vars'])} variables, by degree {sorted(deg.items())} -> {out}")


if __name__ == "__main__":
    main()
