"""Code generator: monomial table  ->  register-resident "adjoint Horner" straight-line program.

Why not re-emit the reference's Maple program?  In source order it keeps 4,534 (2B) / 9,404 (3B) doubles live per
tuple -- it can only run out of local memory on a GPU.  We instead factor the polynomial as a
greedy multivariate Horner tree of depth <= 4 and run forward value and reverse-mode adjoint in ONE depth-first pass:

    node(M, adj):   v = const(M)
                    for the variable x_i occurring in most remaining monomials:   M_i = {m / x_i : x_i | m}
                        leaf  (M_i = {a_k}):   v   = fma(a_k, x_i, v)          g_i = fma(a_k, adj, g_i)
                        inner:                 w   = node(M_i, adj * x_i)
                                               v   = fma(x_i, w, v)             g_i = fma(adj, w, g_i)

adj is the product of the variables on the path from the root, i.e. dE/d(node value).  The live set is x[n], g[n] and
one (adj, v) pair per tree level: ~2n + 10 doubles -> everything stays in registers, every operand of every
instruction is a register or a constant-bank coefficient, and the FP64 pipe is the only unit the program touches.
Instruction count: 2 per leaf + 3 per inner node = 26.6 k (2B) / 69.9 k (3B) FP64 instructions, fewer issue slots
than the reference program's 19.9 k + 10.7 k / 89.5 k + 22.3 k adds + muls.

The emitted file is C-preprocessor neutral: it uses PIP_FMA / PIP_MUL / PIP_ADD, the coefficient accessor PIP_C(k),
variables X(i), gradient accumulators G(i) and an optional PIP_SYNC() marker every ``sync_every`` instructions
(the CUDA kernels use it to keep the warps of a CTA inside the same instruction-cache window).  The same file is
compiled by nvcc into the kernels and by gcc into a CPU check (tests/test_horner_cpu.py).
"""
from __future__ import annotations

import argparse
import collections
import os
import re

import numpy as np

PAD = 255
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "mbpol_openmm_plugin_b200", "data")
GEN = os.path.join(HERE, "..", "mbpol_openmm_plugin_b200", "csrc", "gen")


class Emitter:
    def __init__(self, nacc=2, sync_every=0):
        self.lines = []
        self.nacc = nacc
        self.sync_every = sync_every
        self.n_fp = 0          # FP64 instructions emitted
        self.since_sync = 0
        self.counts = collections.Counter()

    def op(self, text, kind):
        self.lines.append(text)
        self.n_fp += 1
        self.counts[kind] += 1
        self.since_sync += 1

    def maybe_sync(self):
        if self.sync_every and self.since_sync >= self.sync_every:
            self.lines.append("PIP_SYNC();")
            self.since_sync = 0


def split_greedy(monos):
    """Yield (var, sub-monomials) groups: repeatedly pull out the variable that divides most remaining monomials."""
    rest = monos
    while rest:
        cnt = collections.Counter()
        for vs, _ in rest:
            for v in set(vs):
                cnt[v] += 1
        best = max(cnt.items(), key=lambda kv: (kv[1], -kv[0]))[0]
        sub, other = [], []
        for vs, a in rest:
            if best in vs:
                lst = list(vs)
                lst.remove(best)
                sub.append((tuple(lst), a))
            else:
                other.append((vs, a))
        rest = other
        yield best, sub


def gen_node(em, monos, adj, depth):
    """Emit code for one Horner node; returns the name of the variable holding its value (or a PIP_C() constant)."""
    const = [a for vs, a in monos if not vs]
    assert len(const) <= 1
    rest = [(vs, a) for vs, a in monos if vs]
    groups = list(split_greedy(rest))
    leaves = [(v, sub[0][1]) for v, sub in groups if len(sub) == 1 and not sub[0][0]]
    inner = [(v, sub) for v, sub in groups if not (len(sub) == 1 and not sub[0][0])]

    nacc = max(1, min(em.nacc, (len(leaves) + len(inner) + 3) // 4))
    acc = [f"v{depth}_{k}" for k in range(nacc)]
    started = [False] * nacc
    if const:
        em.lines.append(f"{acc[0]} = PIP_C({const[0]});")
        started[0] = True
    turn = [0]

    def accumulate(a, b):
        """acc += a*b on the next accumulator (round-robin breaks the dependent FMA chain)."""
        k = turn[0] % nacc
        turn[0] += 1
        if started[k]:
            em.op(f"{acc[k]} = PIP_FMA({a}, {b}, {acc[k]});", "fma")
        else:
            em.op(f"{acc[k]} = PIP_MUL({a}, {b});", "mul")
            started[k] = True

    # inner children first (deep work), then the leaf run
    for v, sub in inner:
        if adj is None:
            child_adj = f"X({v})"
        else:
            child_adj = f"a{depth + 1}"
            em.op(f"{child_adj} = PIP_MUL({adj}, X({v}));", "mul")
        w = gen_node(em, sub, child_adj, depth + 1)
        accumulate(f"X({v})", w)
        if adj is None:
            em.op(f"G({v}) = PIP_ADD(G({v}), {w});", "add")
        else:
            em.op(f"G({v}) = PIP_FMA({adj}, {w}, G({v}));", "fma")
        em.maybe_sync()
    for v, a in leaves:
        accumulate(f"PIP_C({a})", f"X({v})")
        if adj is None:
            em.op(f"G({v}) = PIP_ADD(G({v}), PIP_C({a}));", "add")
        else:
            em.op(f"G({v}) = PIP_FMA(PIP_C({a}), {adj}, G({v}));", "fma")
    em.maybe_sync()
    live = [acc[k] for k in range(nacc) if started[k]]
    if not live:
        em.lines.append(f"{acc[0]} = 0.0;")
        live = [acc[0]]
    while len(live) > 1:                       # fixed-shape pairwise combine
        nxt = []
        for i in range(0, len(live) - 1, 2):
            em.op(f"{live[i]} = PIP_ADD({live[i]}, {live[i + 1]});", "add")
            nxt.append(live[i])
        if len(live) % 2:
            nxt.append(live[-1])
        live = nxt
    return live[0]


def generate(name, tab, nacc=2, sync_every=0):
    mv = tab["mono_vars"]
    mc = tab["mono_coef"]
    monos = [(tuple(int(v) for v in row if v != PAD), int(c)) for row, c in zip(mv, mc)]
    n_vars = int(tab["n_vars"])
    em = Emitter(nacc=nacc, sync_every=sync_every)
    root = gen_node(em, monos, None, 0)
    decl = []
    for d in range(5):
        decl.append("double " + ", ".join(f"v{d}_{k}" for k in range(max(1, nacc))) + (f", a{d}" if d else "") + ";")
    head = [
        f"/* GENERATED by tools/gen_horner.py from data/{name}.npz -- do not edit.",
        f" * {name}: {len(monos)} monomials, {n_vars} variables; adjoint-Horner straight-line program,",
        f" * {em.n_fp} FP64 instructions ({dict(em.counts)}), accumulators per node = {nacc}.",
        " * Needs: PIP_FMA(a,b,c) PIP_MUL(a,b) PIP_ADD(a,b) PIP_C(k) X(i) G(i) PIP_SYNC(); G(i) must be zeroed by the",
        " * caller; the polynomial value is left in PIP_RESULT. */",
    ]
    body = head + decl + em.lines + [f"PIP_RESULT = {root};", ""]
    em.root = root
    em.n_vars = n_vars
    return "\n".join(body), em


_CALL = re.compile(r"^(\S+) = (PIP_FMA|PIP_MUL|PIP_ADD)\((.*)\);$")
_MOVE = re.compile(r"^(\S+) = (PIP_C\(\d+\)|0\.0);$")


def to_ptx(name, lines, n_vars, n_coef_values, nacc, block=256, min_blocks=1):
    """Translate the neutral program into a complete PTX module with one kernel:

        {name}_eval(const double* x /*[n_vars][ld]*/, double* out /*[1+n_vars][ld]: E, g*/, u32 n, u32 ld)

    one thread per tuple, x/g/adjoints in virtual registers, coefficients in the module's .const bank (ptxas folds
    the ld.const into the c[bank][offset] operand of DFMA), bar.sync at every PIP_SYNC marker."""
    out = []
    w = out.append
    w("// GENERATED by tools/gen_horner.py -- adjoint-Horner PIP evaluator, see the file header of that tool.")
    w(".version 8.7")
    w(".target sm_100a")
    w(".address_size 64")
    w("")
    w(f".const .align 8 .f64 {name}_coef[{len(n_coef_values)}] = {{")
    w(",\n".join("0d%016X" % int(np.float64(c).view(np.uint64)) for c in n_coef_values))
    w("};")
    w("")
    w(f".visible .entry {name}_eval(.param .u64 p_x, .param .u64 p_out, .param .u32 p_n, .param .u32 p_ld)")
    w(f".maxntid {block}, 1, 1")
    w(f".minnctapersm {min_blocks}")
    w("{")
    w(f".reg .f64 %x<{n_vars}>, %g<{n_vars}>, %k<4>, %res;")
    for d in range(5):
        w(".reg .f64 " + ", ".join(f"%v{d}_{k}" for k in range(max(1, nacc))) + (f", %a{d}" if d else "") + ";")
    w(".reg .u64 %px, %po, %off, %ldb; .reg .u32 %n, %ld, %t, %tc, %r<3>; .reg .pred %valid;")
    w("ld.param.u64 %px, [p_x]; ld.param.u64 %po, [p_out]; ld.param.u32 %n, [p_n]; ld.param.u32 %ld, [p_ld];")
    w("cvta.to.global.u64 %px, %px; cvta.to.global.u64 %po, %po;")
    w("mov.u32 %r0, %ctaid.x; mov.u32 %r1, %ntid.x; mov.u32 %r2, %tid.x;")
    w("mad.lo.u32 %t, %r0, %r1, %r2;")
    w("setp.lt.u32 %valid, %t, %n;")
    w("sub.u32 %tc, %n, 1; min.u32 %tc, %tc, %t;")            # clamp: idle lanes recompute the last tuple
    w("mul.wide.u32 %off, %tc, 8; add.u64 %px, %px, %off;")
    w("mul.wide.u32 %ldb, %ld, 8;")
    for i in range(n_vars):
        w(f"ld.global.nc.f64 %x{i}, [%px]; add.u64 %px, %px, %ldb;")
        w(f"mov.f64 %g{i}, 0d0000000000000000;")

    def reg(tok):
        tok = tok.strip()
        m = re.fullmatch(r"X\((\d+)\)", tok)
        if m:
            return "%x" + m.group(1)
        m = re.fullmatch(r"G\((\d+)\)", tok)
        if m:
            return "%g" + m.group(1)
        return "%" + tok

    held, rr = {}, [0]
    for ln in lines:
        if ln == "PIP_SYNC();":
            w("bar.sync 0;")
            continue
        m = _MOVE.match(ln)
        if m:
            dst, src = m.groups()
            if src == "0.0":
                w(f"mov.f64 {reg(dst)}, 0d0000000000000000;")
            else:
                k = int(src[6:-1])
                w(f"ld.const.f64 {reg(dst)}, [{name}_coef+{8 * k}];")
            continue
        m = _CALL.match(ln)
        assert m, ln
        dst, fn, argstr = m.groups()
        args = [a.strip() for a in re.split(r",\s*(?![^()]*\))", argstr)]
        ops = []
        for a in args:
            if a.startswith("PIP_C("):
                k = int(a[6:-1])
                if k in held:                      # the two FMAs of a leaf share one coefficient load
                    ops.append(held[k])
                else:
                    r = f"%k{rr[0] % 4}"
                    rr[0] += 1
                    for kk in [kk for kk, v in held.items() if v == r]:
                        del held[kk]
                    held[k] = r
                    w(f"ld.const.f64 {r}, [{name}_coef+{8 * k}];")
                    ops.append(r)
            else:
                ops.append(reg(a))
        ins = {"PIP_FMA": "fma.rn.f64", "PIP_MUL": "mul.rn.f64", "PIP_ADD": "add.rn.f64"}[fn]
        w(f"{ins} {reg(dst)}, {', '.join(ops)};")
    return out


def ptx_tail(result, n_vars):
    out = []
    w = out.append
    w("@!%valid bra DONE;")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off;")
    w(f"st.global.f64 [%po], %{result}; add.u64 %po, %po, %ldb;")
    for i in range(n_vars):
        w(f"st.global.f64 [%po], %g{i}; add.u64 %po, %po, %ldb;")
    w("DONE:")
    w("ret;")
    w("}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nacc", type=int, default=2)
    ap.add_argument("--sync-every", type=int, default=1024)
    ap.add_argument("--out", default=GEN)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--min-blocks", type=int, default=1)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name in ("pip2b", "pip3b"):
        tab = dict(np.load(os.path.join(DATA, name + ".npz")))
        text, em = generate(name, tab, args.nacc, args.sync_every)
        path = os.path.join(args.out, f"{name}_horner.inc")
        with open(path, "w") as f:
            f.write(text)
        coef = os.path.join(args.out, f"{name}_coef.inc")
        with open(coef, "w") as f:
            f.write(f"/* GENERATED by tools/gen_horner.py: {len(tab['coef'])} linear coefficients of {name} (kcal/mol) */\n")
            f.write(",\n".join(float(c).hex() for c in tab["coef"]))
            f.write("\n")
        ptx = to_ptx(name, em.lines, em.n_vars, tab["coef"], max(1, args.nacc), args.block, args.min_blocks)
        ptx += ptx_tail(em.root, em.n_vars)
        with open(os.path.join(args.out, f"{name}_eval.ptx"), "w") as f:
            f.write("\n".join(ptx) + "\n")
        print(f"{name}: {em.n_fp} FP64 instr {dict(em.counts)} -> {path} (+ .ptx)")


if __name__ == "__main__":
    main()
