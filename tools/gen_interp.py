"""Code generator: monomial table -> cache-resident adjoint-Horner evaluator (command stream + small sm_100a PTX kernel).

The polynomial E(x) = sum_m a[c_m] prod_{v in m} x_v (data/pip{2b,3b}.npz, degree <= 4) is factored as a greedy
multivariate Horner tree and value + reverse-mode gradient are computed in ONE depth-first pass:

    node(M, adj):  v = const(M)
       leaf  a_k x_i :  v = fma(a_k, x_i, v)      g_i = fma(a_k, adj, g_i)
       inner x_i (..):  w = node(M / x_i, adj x_i)  v = fma(x_i, w, v)   g_i = fma(adj, w, g_i)

Fully unrolled this is 1.4 MB of SASS and instruction-fetch bound (profiles/r01_*; tools/attic/gen_horner_r01.py), so
the program is DATA: the tree is serialised depth-first into a command stream and a few-thousand-instruction kernel
interprets it, one thread per tuple, all per-tuple state in registers:

    ENTER_d k, c0  : push (v, a);  a *= x_k;  v = c0               (x_k: dynamic index -> shared-memory copy of x)
    LEAVE_d k      : w = v; pop (v, a);  v += x_k w;  g_k += a w   (g_k: dynamic -> shared-memory accumulator)
    SIBLING_d      : LEAVE_d k_old + ENTER_d k_new, c0 in one record
    MASK_h coefs   : for every register slot j of handler h:  v += c_j x_j ;  g_j += c_j a
    LEAFNODE_h     : ENTER + MASK_h + LEAVE fused (nodes without inner children: 83 % of all nodes)
    DUAL_h         : two consecutive LEAFNODE_h records (siblings) in one dispatch, coefficients interleaved
    TRANSITION_p   : end of symmetry phase p (see below)
    NEXT / END     : flip the stream buffers / finish

Symmetry phases.  Every variable is a function of ONE site-site distance; a hydrogen (or lone-pair)
swap inside one molecule permutes the variables and leaves the polynomial invariant.  A node whose leaf coefficients
are invariant under a set S of swaps is a linear form in the S-ORBIT SUMS of the variables, and its gradient
contribution is the same number for every member of an orbit.  The stream is run in PHASES, one per subgroup S that
pays for itself: the registers hold the orbit sums of that phase (slot = lowest variable of the orbit), the leaf
handlers work on slots, TRANSITION scatters the register gradients to the orbit members and loads the next phase.
Leaf handlers are specialised on register-slot SETS chosen by an agglomerative merge (`merge_library`); payloads are
dense per handler (zero padding), so a handler needs no indexing at all.

DUAL records (round 2).  90 % (3B) / 95 % (2B) of the leaf-node records sit next to a sibling with the same handler.
One handler evaluates both: four independent value chains instead of two (the dependent-issue stalls of the value FMAs
were the largest stall reason of the round-1 kernel), the second value FMA of a slot reads x_j from the operand-reuse
cache (a DFMA with three fresh 64-bit register operands issues every third cycle, not every second,
profiles/r01_dfma_operand_microbench.md), and header / dispatch / address arithmetic are paid once per two records.

The round-1 experiments that lost (variable-class kernels, STATIC handlers, constant-bank stream, x_k prefetch, ...)
live in tools/attic/gen_interp_r01.py together with their measurements' pointers; this file is the shipped path plus
the knobs that are still being measured.
"""
from __future__ import annotations

import argparse
import collections
import itertools
import json
import os
import struct

import numpy as np

PAD = 255
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "mbpol_openmm_plugin_b200", "data")
GEN = os.path.join(HERE, "..", "mbpol_openmm_plugin_b200", "csrc", "gen")

NV = {"pip2b": 31, "pip3b": 36}
OP_END, OP_ENTER, OP_LEAVE, OP_NEXT, OP_MASK = 0, 1, 4, 7, 8
# ENTER_d = 1..3, LEAVE_d = 4..6, MASK_h = 8 + h, LEAFNODE_h = 8 + nm + h, TRANSITION_p = 8 + 2 nm + p,
# SIBLING_d = 8 + 2 nm + n_phases + d - 1, DUAL_i = 8 + 2 nm + n_phases + 3 + i (i indexes the handlers that have one)
MAX_DEPTH = 3


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


class Node:
    __slots__ = ("var", "const", "leaves", "kids")

    def __init__(self, var):
        self.var, self.const, self.leaves, self.kids = var, None, {}, []


def build_tree(monos, var=None):
    node = Node(var)
    rest = []
    for vs, a in monos:
        if vs:
            rest.append((vs, a))
        else:
            assert node.const is None
            node.const = a
    for v, sub in split_greedy(rest):
        if len(sub) == 1 and not sub[0][0]:
            node.leaves[v] = sub[0][1]
        else:
            node.kids.append(build_tree(sub, v))
    return node


# ---- symmetry -----------------------------------------------------------------------------------------------------------
def _sites(name):
    """site pair of every variable (the same table as MBPOL_{2B,3B}_VARS in csrc/tuples.cu) and the
    site swaps that leave the polynomial invariant"""
    if name == "pip3b":
        p = [("Ha1", "Ha2"), ("Hb1", "Hb2"), ("Hc1", "Hc2"), ("Oa", "Ha1"), ("Oa", "Ha2"), ("Ob", "Hb1"), ("Ob", "Hb2"),
             ("Oc", "Hc1"), ("Oc", "Hc2")]
        p += [(h, t) for h in ("Ha1", "Ha2") for t in ("Hb1", "Hb2", "Hc1", "Hc2")]
        p += [(h, t) for h in ("Hb1", "Hb2") for t in ("Hc1", "Hc2")]
        p += [("Oa", t) for t in ("Hb1", "Hb2", "Hc1", "Hc2")] + [("Ob", t) for t in ("Ha1", "Ha2", "Hc1", "Hc2")]
        p += [("Oc", t) for t in ("Ha1", "Ha2", "Hb1", "Hb2")] + [("Oa", "Ob"), ("Oa", "Oc"), ("Ob", "Oc")]
        swaps = [("Ha1", "Ha2"), ("Hb1", "Hb2"), ("Hc1", "Hc2")]
    else:
        p = [("Ha1", "Ha2"), ("Hb1", "Hb2"), ("Oa", "Ha1"), ("Oa", "Ha2"), ("Ob", "Hb1"), ("Ob", "Hb2")]
        p += [(h, t) for h in ("Ha1", "Ha2") for t in ("Hb1", "Hb2")]
        p += [("Oa", "Hb1"), ("Oa", "Hb2"), ("Ob", "Ha1"), ("Ob", "Ha2"), ("Oa", "Ob")]
        p += [(x, t) for x in ("Xa1", "Xa2") for t in ("Hb1", "Hb2")] + [(x, t) for x in ("Xb1", "Xb2") for t in ("Ha1", "Ha2")]
        p += [("Oa", "Xb1"), ("Oa", "Xb2"), ("Ob", "Xa1"), ("Ob", "Xa2")]
        p += [(x, t) for x in ("Xa1", "Xa2") for t in ("Xb1", "Xb2")]
        swaps = [("Ha1", "Ha2"), ("Hb1", "Hb2"), ("Xa1", "Xa2"), ("Xb1", "Xb2")]
    return p, swaps


def swap_perms(name):
    """variable permutations induced by the site swaps"""
    sites, swaps = _sites(name)
    vid = {frozenset(p): i for i, p in enumerate(sites)}
    perms = []
    for a, b in swaps:
        m = {a: b, b: a}
        perms.append([vid[frozenset((m.get(p, p), m.get(q, q)))] for p, q in sites])
    return perms


def _orbits(nv, perms):
    seen, orbs = set(), []
    for v in range(nv):
        if v in seen:
            continue
        o, todo = {v}, [v]
        while todo:
            u = todo.pop()
            for pm in perms:
                if pm[u] not in o:
                    o.add(pm[u])
                    todo.append(pm[u])
        seen |= o
        orbs.append(sorted(o))
    return orbs


def adapted_basis(name, nv, perms, opts):
    """Symmetry-adapted slot basis: an invertible 0/1 change of basis of the variables, block by block over the orbits
    of the whole swap group:
        orbit {v}                    -> v
        orbit {v1, v2}               -> v1 + v2,  v2
        orbit {w11, w12, w21, w22}   -> sum of all four,  w21 + w22,  w12 + w22,  w22
    (`opts`: per block kind "2" = as written, "1" = keep the members with index 1 instead, "none" = leave the block alone;
    kind = (orbit size, site kinds), see `kind`).  A node's linear form sum_v c_v x_v is rewritten in this basis -- the
    inverse has entries 0, +-1, the new coefficients are sums / differences of the old ones -- so leaves that are symmetric
    under a swap collapse onto the sum rows: the saving of the orbit-sum phases for ALL swaps at once, in one phase and in
    the same registers; a lone leaf on a "first" member costs up to four slots instead of one.
    -> rows (sorted member lists; with the "2" orientation every row has a distinct lowest variable)"""
    sites, _ = _sites(name)

    def hidx(v):
        return tuple(t[-1] if t[0] in "HX" else "0" for t in sites[v])

    def kind(o):
        p, q = sites[o[0]]
        return (len(o), tuple(sorted(t[0] + ("i" if p[1] == q[1] else "x") for t in (p, q))))
    rows = []
    for o in _orbits(nv, perms):
        opt = opts.get(kind(o), "none") if len(o) > 1 else "none"
        if opt == "none":
            rows += [[v] for v in o]
        elif len(o) == 2:
            rows += [o, [max(o, key=hidx) if opt == "2" else min(o, key=hidx)]]
        else:
            assert len(o) == 4
            ids = {v: hidx(v) for v in o}
            a, b = [q for q in range(2) if len({ids[v][q] for v in o}) == 2]
            rows += [o, [v for v in o if ids[v][a] == opt], [v for v in o if ids[v][b] == opt],
                     [v for v in o if ids[v][a] == opt and ids[v][b] == opt]]
    rows = [sorted(r) for r in rows]
    assert len(rows) == nv
    return rows


# measured best orientation per block kind (tools/gen_interp.py --basis; all 3^k combinations were counted)
BASIS_OPTS = {"pip2b": {(2, ("Hx", "Ox")): "2", (2, ("Ox", "Xx")): "2", (4, ("Hx", "Xx")): "2", (4, ("Xx", "Xx")): "2"},
              "pip3b": {(2, ("Hi", "Oi")): "2", (2, ("Hx", "Ox")): "2", (4, ("Hx", "Hx")): "2"}}
# ("1" rows share their lowest variable with the sum row: usable by gen_msym.py, whose slots are numbered, not here)


def _popcount(a):
    a = a.copy()
    c = np.zeros(a.shape, dtype=np.int64)
    while a.any():
        c += (a & np.uint64(1)).astype(np.int64)
        a >>= np.uint64(1)
    return c


def merge_library(req, K):
    """req: {slot bitset: frequency}.  Agglomerative merge to K handler slot sets minimising sum(freq * |handler|)."""
    sets, freq = [int(k) for k in sorted(req)], [int(req[k]) for k in sorted(req)]
    while len(sets) > K:
        s, f = np.array(sets, dtype=np.uint64), np.array(freq, dtype=np.int64)
        pc, pu = _popcount(s), _popcount(s[:, None] | s[None, :])
        d = f[:, None] * (pu - pc[:, None]) + f[None, :] * (pu - pc[None, :])
        d[np.tril_indices(len(sets))] = 1 << 60
        i, j = divmod(int(np.argmin(d)), len(sets))
        sets[i] |= sets[j]
        freq[i] += freq[j]
        del sets[j], freq[j]
    return sets


def serialise(name, tab, K, skeleton_weight=40, dual_max=0, basis=False):
    """-> (records, masks, phases, dual).  records: (opcode, k, payload doubles); masks: handler slot tuples; phases: list
    of orbit lists (phase 0 = identity); dual: handler ids that have a DUAL record type, in opcode order."""
    nv = int(tab["n_vars"])
    coef = tab["coef"]
    monos = [(tuple(int(v) for v in row if v != PAD), int(c)) for row, c in zip(tab["mono_vars"], tab["mono_coef"])]
    perms = swap_perms(name)
    table = {tuple(sorted(vs)): c for vs, c in monos}
    for pm in perms:                                    # the premise: the polynomial is invariant under every swap
        assert all(table.get(tuple(sorted(pm[v] for v in vs))) == c for vs, c in monos)
    subsets = [c for r in range(len(perms) + 1) for c in itertools.combinations(range(len(perms)), r)]
    orbs = {S: _orbits(nv, [perms[i] for i in S]) for S in subsets}
    root = build_tree(monos)
    nodes = []

    def walk(node, depth):
        nodes.append((node, depth))
        for k in node.kids:
            walk(k, depth + 1)
    walk(root, 0)
    # every leaf-bearing node: the subgroups under which its leaf coefficients are invariant, with the slots each needs
    valid = {}
    for node, _ in nodes:
        if node.leaves:
            valid[id(node)] = [(S, {o[0]: node.leaves[o[0]] for o in orbs[S] if o[0] in node.leaves}) for S in subsets
                               if all(len({node.leaves.get(v) for v in o}) == 1 for o in orbs[S])]

    def classify(allowed):
        return {i: min((c for c in v if c[0] in allowed), key=lambda c: (len(c[1]), -len(c[0]))) for i, v in valid.items()}

    def cost(allowed):
        """2 FMAs per slot + the tree skeleton repeated per phase (4 FP64 + two dispatches ~ 40 issue slots per node)"""
        c = classify(allowed)
        tot = 2 * sum(len(need) for _, need in c.values())

        def sk(node):
            ph = {c[id(node)][0]} if node.leaves else set()
            for k in node.kids:
                ph |= sk(k)
            nonlocal tot
            if node.kids:
                tot += skeleton_weight * len(ph)
            return ph
        sk(root)
        return tot
    allowed = set(subsets)
    while True:
        allowed = {S for S, _ in classify(allowed).values()} | {()}
        base = cost(allowed)
        trial = [(cost(allowed - {S}), S) for S in sorted(allowed) if S != ()]
        if not trial or min(trial)[0] >= base:
            break
        allowed.discard(min(trial)[1])
    cls = classify(allowed)
    used = sorted(allowed, key=lambda S: (len(S), S))
    pidx = {S: i for i, S in enumerate(used)}
    phases = [orbs[S] for S in used]
    if basis:       # ONE phase in the symmetry-adapted basis (adapted_basis): need = {slot: coefficient VALUE}
        rows = adapted_basis(name, nv, perms, BASIS_OPTS[name])
        assert len({r[0] for r in rows}) == nv           # a row's register slot = its lowest variable: must be distinct
        B = np.zeros((nv, nv))
        for q, m in enumerate(rows):
            B[q, m] = 1.0
        BinvT = np.rint(np.linalg.inv(B)).T
        assert np.array_equal(BinvT.T @ B, np.eye(nv))
        cls = {}
        for node, _ in nodes:
            if node.leaves:
                c = np.zeros(nv)
                for v, ci in node.leaves.items():
                    c[v] = coef[ci]
                cp = BinvT @ c
                cls[id(node)] = ((), {rows[q][0]: float(cp[q]) for q in range(nv) if cp[q] != 0.0})
        used, pidx, phases = [()], {(): 0}, [rows]
    bits = lambda need: sum(1 << j for j in need)  # noqa: E731
    req = collections.Counter(bits(need) for _, need in cls.values())
    lib = merge_library(req, K)
    fr = collections.Counter()
    pick = {}
    for b, f in req.items():
        h = min((h for h in lib if h & b == b), key=lambda h: (bin(h).count("1"), h))
        pick[b] = h
        fr[h] += f
    lib = [h for h, _ in fr.most_common()]
    hid = {h: i for i, h in enumerate(lib)}
    masks = [tuple(j for j in range(nv) if h >> j & 1) for h in lib]
    nm = len(masks)
    leafph = {}

    def leaf_phases(node):
        s = set()
        if node.leaves:
            s.add(pidx[cls[id(node)][0]])
        for k in node.kids:
            s |= leaf_phases(k)
        leafph[id(node)] = s
        return s
    leaf_phases(root)
    recs = []

    def payload(node):
        need = cls[id(node)][1]
        return [(need[j] if basis else coef[need[j]]) if j in need else 0.0 for j in masks[hid[pick[bits(need)]]]]

    def emit(node, depth, p):
        assert depth <= MAX_DEPTH
        own = bool(node.leaves) and pidx[cls[id(node)][0]] == p
        kids = [k for k in node.kids if p in leafph[id(k)]]
        const = coef[node.const] if (node.const is not None and p == min(leafph[id(node)])) else 0.0
        h = hid[pick[bits(cls[id(node)][1])]] if own else -1

        def order(k):       # evaluation order of siblings is free: childless ones grouped by handler -> equal records adjacent
            childless = not any(p in leafph[id(q)] for q in k.kids)
            return (0, hid[pick[bits(cls[id(k)][1])]]) if childless else (1, 0)
        kids.sort(key=order)
        if depth > 0 and not kids:
            assert own
            recs.append((OP_MASK + nm + h, node.var, [const] + payload(node)))
            return
        if depth > 0:
            recs.append((OP_ENTER + depth - 1, node.var, [const]))
        else:
            assert node.const is None
        if own:
            recs.append((OP_MASK + h, 0, payload(node)))
        for k in kids:
            emit(k, depth + 1, p)
        if depth > 0:
            recs.append((OP_LEAVE + depth - 1, node.var, []))
    for p in range(len(phases)):
        if p in leafph[id(root)]:
            emit(root, 0, p)
        recs.append((OP_MASK + 2 * nm + p, 0, []))
    recs.append((OP_END, 0, []))
    # SIBLING_d = LEAVE_d k_old immediately followed by ENTER_d k_new, c0: one record (k = k_new | k_old << 8)
    base, out, i = OP_MASK + 2 * nm + len(phases), [], 0
    while i < len(recs):
        op, k, pl = recs[i]
        if OP_LEAVE <= op < OP_NEXT and i + 1 < len(recs) and recs[i + 1][0] == op - OP_LEAVE + OP_ENTER:
            out.append((base + op - OP_LEAVE, recs[i + 1][1] | (k << 8), recs[i + 1][2]))
            i += 2
        else:
            out.append(recs[i])
            i += 1
    recs = out
    # DUAL_i = two consecutive LEAFNODE_h records (always siblings: distinct node variables), for handlers of at most
    # dual_max slots: k = k_A | k_B << 8, payload = c0_A, c0_B, c_A1, c_B1, c_A2, c_B2, ... (interleaved)
    dual = [h for h in range(nm) if len(masks[h]) <= dual_max]
    if dual:
        dbase, out, i = base + MAX_DEPTH, [], 0
        while i < len(recs):
            op, k, pl = recs[i]
            h = op - OP_MASK - nm
            if 0 <= h < nm and h in dual and i + 1 < len(recs) and recs[i + 1][0] == op:
                kb, plb = recs[i + 1][1], recs[i + 1][2]
                assert kb != k and len(pl) == len(plb)
                out.append((dbase + dual.index(h), k | (kb << 8), [c for ab in zip(pl, plb) for c in ab]))
                i += 2
            else:
                out.append(recs[i])
                i += 1
        recs = out
    return recs, masks, phases, dual


def rec_words(n_payload):
    return 1 + (n_payload if n_payload % 2 else n_payload + 1)


def pack(recs, chunk_bytes):
    """u64 words, organised in chunks of ``chunk_bytes`` that the kernel stages through shared memory one at a time.
    Inside a chunk, word 0 is padding so that every record header sits at byte offset 8 (mod 16) and every payload is
    16-byte aligned (payloads are padded to an odd number of words).  No record straddles a chunk: the last record of a
    chunk is OP_NEXT.  One all-zero chunk is appended so that the kernel's look-ahead loads never leave the array.
    Look-ahead opcodes: the low word of record r's header holds the opcode of record r + 1 (word 0 of the stream holds
    the first opcode), so a handler knows its successor on entry and its closing brx.idx never waits for a load."""
    cw = chunk_bytes // 8
    chunks = []
    cur = [0]
    hdrs = []                                     # (chunk index, word index, op) of every header in execution order
    for op, k, payload in recs:
        need = rec_words(len(payload))
        if len(cur) + need + 18 > cw:             # room for OP_NEXT + the look-ahead words the handlers preload
            hdrs.append((len(chunks), len(cur), OP_NEXT))
            cur += [OP_NEXT, 0]
            cur += [0] * (cw - len(cur))
            chunks.append(cur)
            cur = [0]
        hdrs.append((len(chunks), len(cur), op))
        cur.append(op | (k << 32))
        p = list(payload)
        if len(p) % 2 == 0:
            p.append(0.0)
        cur.extend(struct.unpack("<Q", struct.pack("<d", float(c)))[0] for c in p)
    cur += [0] * (cw - len(cur))
    chunks.append(cur)
    chunks.append([0] * cw)
    chunks[0][0] = hdrs[0][2]
    for (c, i, _), (_, _, nxt) in zip(hdrs, hdrs[1:] + [(0, 0, OP_END)]):
        chunks[c][i] = (chunks[c][i] & ~0xFFFFFFFF) | nxt
    return np.array([w for c in chunks for w in c], dtype=np.uint64)


def interpret(name, words, x, chunk_bytes, masks, phases, dual=()):
    """numpy/python reference interpreter of the packed stream (used by the CPU tests): the same records, in the same
    order, with the same meaning as the PTX handlers below."""
    cw = chunk_bytes // 8
    nv = NV[name]
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros(nv)                       # the shared-memory accumulators (all of dE/dx at the end)
    xr, gr = x.copy(), np.zeros(nv)        # register slots: orbit sums / orbit gradients of the current phase
    for o in phases[0]:
        xr[o[0]] = sum(x[i] for i in o)
    f = words.view(np.float64)
    nm = len(masks)
    t_base = OP_MASK + 2 * nm
    s_base = t_base + len(phases)
    d_base = s_base + MAX_DEPTH
    pc = 1
    v, a = 0.0, 1.0
    stack = []
    nxt = int(words[0]) & 0xFFFFFFFF
    while True:
        hdr = int(words[pc])
        op, k = hdr & 0xFFFFFFFF, hdr >> 32
        op, nxt = nxt, op
        if op == OP_END:
            return v, g
        if op == OP_NEXT:
            pc = (pc // cw + 1) * cw + 1
        elif OP_ENTER <= op < OP_LEAVE:
            stack.append((v, a, x[k]))
            a = a * x[k]
            v = f[pc + 1]
            pc += 2
        elif OP_LEAVE <= op < OP_NEXT:
            w = v
            v, a, xk = stack.pop()
            v += xk * w
            g[k] += a * w
            pc += 2
        elif op < OP_MASK + nm:
            m = masks[op - OP_MASK]
            for q, j in enumerate(m):
                c = f[pc + 1 + q]
                v += c * xr[j]
                gr[j] += c * a
            pc += rec_words(len(m))
        elif op < t_base:                                   # LEAFNODE_h: ENTER k, const; MASK_h; LEAVE k
            m = masks[op - OP_MASK - nm]
            a2 = a * x[k]
            u = f[pc + 1]
            for q, j in enumerate(m):
                c = f[pc + 2 + q]
                u += c * xr[j]
                gr[j] += c * a2
            v += x[k] * u
            g[k] += a * u
            pc += rec_words(1 + len(m))
        elif op < s_base:                                   # TRANSITION_p
            p = op - t_base
            for o in phases[p]:
                for i in o:
                    g[i] += gr[o[0]]
            if p + 1 < len(phases):
                for o in phases[p + 1]:
                    xr[o[0]] = x[o[0]]
                    for i in o[1:]:
                        xr[o[0]] += x[i]
                    gr[o[0]] = 0.0
            pc += 2
        elif op < d_base:                                   # SIBLING_d: LEAVE k_old + ENTER k_new, c0
            w = v
            v, a, xk = stack.pop()
            v += xk * w
            g[k >> 8] += a * w
            stack.append((v, a, x[k & 0xFF]))
            a = a * x[k & 0xFF]
            v = f[pc + 1]
            pc += 2
        else:                                               # DUAL_i: two sibling leaf nodes, interleaved payload
            m = masks[dual[op - d_base]]
            for s, kk in enumerate((k & 0xFF, k >> 8)):
                a2 = a * x[kk]
                u = f[pc + 1 + s]
                for q, j in enumerate(m):
                    c = f[pc + 3 + 2 * q + s]
                    u += c * xr[j]
                    gr[j] += c * a2
                v += x[kk] * u
                g[kk] += a * u
            pc += rec_words(2 + 2 * len(m))


def kernel_ptx(name, entry, stream, masks, nv, phases, dual, block=256, min_blocks=1, piece=64, fall=(), npre=8,
               fall_copies=2, group=4, maxnreg=0, slim=False, anchors=False):
    """One interpreter kernel

        {entry}(const double* x /*[nv][ld]*/, double* out /*[1+nv][ld]: E, dE/dx*/, u32 n, u32 ld, u32 mode)

    mode bit 0: add into `out` instead of overwriting it; bit 1: rows 1.. receive x_i dE/dx_i instead of dE/dx_i (every
    variable is exp(..) [/ r] of one distance, so its chain rule needs exactly that product: the force kernels then
    never read x again, csrc/tuples.cu).

    One thread per tuple; x / g slots and the adjoints in registers.  The stream is staged through a double-buffered
    shared-memory window: when chunk c starts every thread issues cp.async copies of its 16-byte pieces of chunk c + 1
    straight into the idle buffer; at NEXT a cp.async.wait_all and one bar.sync flip the buffers.  Handlers read
    coefficients with broadcast ld.shared.v2.f64 (every lane the same address).  Every handler pre-loads the next
    record's header and first `npre` payload words -- before its own g_k store, which ptxas could not move them across --
    and ends with its own brx.idx on the opcode it already knew on entry.
    Returns (lines, dynamic shared bytes)."""
    NT = block
    CH = NT * piece                    # chunk bytes = one piece per thread
    assert piece > 16 and piece % 16 == 0 and npre % 2 == 0
    nm = len(masks)
    t_base = OP_MASK + 2 * nm
    s_base = t_base + len(phases)
    d_base = s_base + MAX_DEPTH
    n_ops = d_base + len(dual)
    SG = nv * NT * 8                   # smem: x copy | g accumulators | 2 stream buffers
    SB = 2 * nv * NT * 8
    out = []
    w = out.append
    w(f".visible .entry {entry}(.param .u64 p_x, .param .u64 p_out, .param .u32 p_n, .param .u32 p_ld, .param .u32 p_acc)")
    w(f".maxntid {NT}, 1, 1")
    # A register cap below what the program would take makes ptxas keep the pre-loaded payload words -- warp-uniform
    # values that are live across the indirect branch -- in UNIFORM registers (one R2UR per 32 bits after the load): the
    # FMAs that consume them then read two vector operands instead of three and issue every second cycle, not every third.
    w(f".maxnreg {maxnreg}" if maxnreg else f".minnctapersm {min_blocks}")
    w("{")
    w(f".reg .f64 %x<{nv}>, %g<{nv}>, %c<72>, %vc, %vd, %ac, %w, %gk, %one, %u0, %u1, %u2, %u3, %a0, %a1, %xk0, %n<24>;")
    w(".reg .f64 " + ", ".join(f"%vs{d}, %vt{d}, %as{d}, %xk{d}" for d in range(1, MAX_DEPTH + 1)) + ";")
    w(".reg .u64 %px, %po, %off, %ldb, %gs; .reg .u32 %n, %ld, %t, %tc, %r<4>, %op, %k, %nop, %nk, %sx, %sg, %addr;")
    w(".reg .u32 %sp, %bufcur, %bufoth, %mine, %ka, %kb, %addrb; .reg .f64 %xkb, %gkb;")
    w(".reg .pred %valid, %pacc, %pself, %pscale; .reg .u32 %acc, %accb;")
    w("ld.param.u32 %acc, [p_acc]; and.b32 %accb, %acc, 1; setp.ne.u32 %pacc, %accb, 0;")
    w("and.b32 %accb, %acc, 2; setp.ne.u32 %pscale, %accb, 0;")
    w("ld.param.u64 %px, [p_x]; ld.param.u64 %po, [p_out]; ld.param.u32 %n, [p_n]; ld.param.u32 %ld, [p_ld];")
    w("cvta.to.global.u64 %px, %px; cvta.to.global.u64 %po, %po;")
    w("mov.u32 %r0, %ctaid.x; mov.u32 %r1, %ntid.x; mov.u32 %r2, %tid.x;")
    w("mad.lo.u32 %t, %r0, %r1, %r2;")
    w("setp.lt.u32 %valid, %t, %n;")
    w("sub.u32 %tc, %n, 1; min.u32 %tc, %tc, %t;")
    w("mul.wide.u32 %off, %tc, 8; add.u64 %px, %px, %off;")
    w("mul.wide.u32 %ldb, %ld, 8;")
    w("mov.u32 %r3, pip_smem;")
    w("shl.b32 %sx, %r2, 3; add.u32 %sx, %sx, %r3;")             # &xs[0][tid]
    w(f"add.u32 %sg, %sx, {SG};")                                  # &gs[0][tid]
    w(f"add.u32 %bufcur, %r3, {SB}; add.u32 %bufoth, %bufcur, {CH};")

    def async_fetch(buf):
        """this thread's 16-byte pieces of the chunk at %gs -> shared buffer `buf` (asynchronously)"""
        w(f"add.u32 %addr, {buf}, %mine;")
        for q in range(piece // 16):
            w(f"cp.async.ca.shared.global [%addr+{q * NT * 16}], [%gs+{q * NT * 16}], 16;")
    w("mul.lo.u32 %mine, %r2, 16;")
    # stream chunk 0 -> buffer 0 (waited for), chunk 1 -> buffer 1 (in flight)
    w(f"mov.u64 %gs, {stream}; cvta.to.global.u64 %gs, %gs; cvt.u64.u32 %off, %mine; add.u64 %gs, %gs, %off;")
    async_fetch("%bufcur")
    w("cp.async.wait_all;")
    w(f"add.u64 %gs, %gs, {CH};")
    async_fetch("%bufoth")
    w("mov.f64 %one, 0d3FF0000000000000;")
    for i in range(nv):
        w(f"ld.global.nc.f64 %x{i}, [%px]; add.u64 %px, %px, %ldb;")
    for i in range(nv):
        w(f"st.shared.f64 [%sx+{i * NT * 8}], %x{i}; st.shared.f64 [%sg+{i * NT * 8}], 0d0000000000000000;")
        w(f"mov.f64 %g{i}, 0d0000000000000000;")
    for o in phases[0]:                           # (basis mode: phase 0 is not the identity; a thread reads its own column)
        for i in o[1:]:
            w(f"ld.shared.f64 %w, [%sx+{i * NT * 8}]; add.rn.f64 %x{o[0]}, %x{o[0]}, %w;")
    w("mov.f64 %vc, 0d0000000000000000; mov.f64 %vd, 0d0000000000000000; mov.f64 %ac, %one;")
    w("bar.sync 0;")
    w("add.u32 %sp, %bufcur, 8;")
    NPRE = npre                                   # payload words of the NEXT record pre-loaded by every handler
    w("TBL: .branchtargets " + ", ".join(f"H{i}" for i in range(n_ops)) + ";")

    def preload(byte_off, start=0):
        """payload words start..NPRE of the record at %sp+byte_off -> %n regs (garbage past short records: unused)"""
        for q in range(start, NPRE, 2):
            w(f"ld.shared.v2.f64 {{%n{q}, %n{q + 1}}}, [%sp+{byte_off + 8 + 8 * q}];")

    def head(nwords):
        """current k <- header fetched by the previous handler; fetch the next header"""
        w("mov.u32 %k, %nk; mov.u32 %op, %nop;")
        w(f"ld.shared.v2.u32 {{%nop, %nk}}, [%sp+{8 * nwords}];")

    fall_state = [None]

    def tail(nwords, preloaded=0, anchor=False):
        preload(8 * nwords, preloaded)
        w(f"add.u32 %sp, %sp, {8 * nwords};")
        if anchor and anchors:
            # "Uniform anchor".  %op / %nop form ONE live-range web across all handlers (they meet at every indirect
            # branch).  ptxas either keeps that web in uniform registers -- dispatches become BRXU with the jump-table read
            # hoisted to the top of the handler -- or in vector registers (BRX R: `branch_resolving` stalls, ncu 0.34
            # instead of 0.06 cycles per issue, ~8 % of the kernel).  Which one it picks depends on register pressure and
            # on how many dispatch sites test the opcode with a uniform predicate, like the duplicated ("fall") handlers
            # do.  The rarely executed handlers (ENTER / LEAVE / MASK: 9 % of the records) can add such sites at the
            # price of one BRA.U each; for the 2B kernel this is what keeps BRXU once the slim stack has removed the
            # register pressure that used to force it (and with it 8 R2UR per record): 1.036 instead of 1.075 ms.
            w("setp.ne.u32 %pself, %op, 0x7fffffff; @%pself brx.idx %op, TBL; bra.uni FINISH;")
            return
        if fall_state[0] is not None:
            # a second copy of this handler follows: when the next record has the same opcode no branch is taken
            w(f"setp.ne.u32 %pself, %op, {fall_state[0]}; @%pself brx.idx %op, TBL;")
            return
        w("brx.idx %op, TBL;")              # per-handler dispatch: ptxas hoists the jump-table read

    def coef(i):
        return f"%n{i}" if i < NPRE else f"%c{i - NPRE}"

    w("ld.shared.u32 %op, [%bufcur];")         # word 0 of the stream: the first opcode
    w("ld.shared.v2.u32 {%nop, %nk}, [%sp];")
    preload(0)
    w("brx.idx %op, TBL;")
    # ---- END
    w(f"H{OP_END}:")
    w("bra.uni FINISH;")
    # ---- NEXT: chunk c + 1 has arrived in the idle buffer: flip, start fetching chunk c + 2 into the buffer just left
    w(f"H{OP_NEXT}:")
    w("mov.u32 %op, %nop;")
    w("cp.async.wait_all;")
    w("bar.sync 0;")
    w("mov.u32 %addr, %bufcur; mov.u32 %bufcur, %bufoth; mov.u32 %bufoth, %addr;")
    w(f"add.u64 %gs, %gs, {CH};")
    async_fetch("%bufoth")
    w("add.u32 %sp, %bufcur, 8;")
    w("ld.shared.v2.u32 {%nop, %nk}, [%sp];")
    preload(0)
    w("brx.idx %op, TBL;")
    # ---- ENTER_d  (payload: const in %n0)
    # slim: a saved level keeps ONE value (v = vc + vd) and its adjoint; x_k is re-read from shared memory when the level
    # is left (4 instead of 8 registers per level).
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_ENTER + d - 1}:")
        head(2)
        xk = "%xk0" if slim else f"%xk{d}"
        w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 {xk}, [%addr];")
        if slim:
            w(f"add.rn.f64 %vs{d}, %vc, %vd; mov.f64 %as{d}, %ac;")
        else:
            w(f"mov.f64 %vs{d}, %vc; mov.f64 %vt{d}, %vd; mov.f64 %as{d}, %ac;")
        w("mov.f64 %vc, %n0; mov.f64 %vd, 0d0000000000000000;")
        preload(16)
        w(f"mul.rn.f64 %ac, %ac, {xk};")
        tail(2, NPRE, anchor=True)
    # ---- LEAVE_d
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_LEAVE + d - 1}:")
        head(2)
        w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        if slim:
            w(f"mad.lo.u32 %addrb, %k, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addrb];")
        preload(16)
        w("add.rn.f64 %w, %vc, %vd;")
        w(f"mov.f64 %ac, %as{d};")
        if slim:
            w(f"fma.rn.f64 %vc, %xk0, %w, %vs{d}; mov.f64 %vd, 0d0000000000000000;")
        else:
            w(f"fma.rn.f64 %vc, %xk{d}, %w, %vs{d}; mov.f64 %vd, %vt{d};")
        w("fma.rn.f64 %gk, %ac, %w, %gk; st.shared.f64 [%addr], %gk;")
        tail(2, NPRE, anchor=True)

    # ---- MASK_h and fused LEAFNODE_h
    def block_handler(label, m, fused):
        """loads first, then `group` value FMAs on two chains alternating with the `group` gradient FMAs of the same
        slots; once both FMAs of a 16-byte payload piece are issued its registers are dead and the same piece of the NEXT
        record is loaded into them; the chain reduction + parent update close the record."""
        w(f"{label}:")
        slots = ([None] if fused else []) + list(m)          # None = the node constant (first payload word)
        n = len(slots)
        nwords = rec_words(n)
        head(nwords)
        if fused:
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addr];")
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        for i in range(NPRE, n, 2):                          # the rest of this record's payload
            if i + 1 < n:
                w(f"ld.shared.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [%sp+{8 + 8 * i}];")
            else:
                w(f"ld.shared.f64 {coef(i)}, [%sp+{8 + 8 * i}];")
        assert n - NPRE <= 64
        acc, adj = (["%u0", "%u1"], "%a0") if fused else (["%vc", "%vd"], "%ac")
        started = set() if fused else {"%vc", "%vd"}
        vops, gops = [], []
        const0 = None                                        # the node constant starts chain %u0 as the addend of its
        for i, j in enumerate(slots):                        # first FMA (no register move)
            c = coef(i)
            if j is None:
                const0 = c
                vops.append(None)
                continue
            a = acc[i % 2]
            if a in started:
                vops.append(f"fma.rn.f64 {a}, {c}, %x{j}, {a};")
            elif a == "%u0" and const0 is not None:
                vops.append(f"fma.rn.f64 {a}, {c}, %x{j}, {const0};")
                started.add(a)
                const0 = None
            else:
                vops.append(f"mul.rn.f64 {a}, {c}, %x{j};")
                started.add(a)
            gops.append(f"fma.rn.f64 %g{j}, {c}, {adj}, %g{j};")
        if const0 is not None:                               # (a fused node with a single slot: chain %u0 never started)
            vops[0] = f"mov.f64 %u0, {const0};"
            started.add("%u0")
        if fused:
            w("mul.rn.f64 %a0, %ac, %xk0;")
        done = 0
        gi = 0
        for q in range(0, n, group):
            cnt = min(group, n - q)
            for u in range(cnt):
                if vops[q + u] is not None:
                    w(vops[q + u])
            for u in range(cnt):
                if slots[q + u] is not None:
                    w(gops[gi])
                    gi += 1
            for p2 in range(q, q + cnt + (cnt & 1), 2):
                if p2 + 1 < NPRE:
                    w(f"ld.shared.v2.f64 {{%n{p2}, %n{p2 + 1}}}, [%sp+{8 * nwords + 8 + 8 * p2}];")
                    done = p2 + 2
        preload(8 * nwords, done)
        if fused:
            if "%u1" in started:
                w("add.rn.f64 %u0, %u0, %u1;")
            w("fma.rn.f64 %vc, %xk0, %u0, %vc;")
            w("fma.rn.f64 %gk, %ac, %u0, %gk;")
            w("st.shared.f64 [%addr], %gk;")
        tail(nwords, NPRE, anchor=not fused)

    def dual_handler(label, m):
        """two sibling LEAFNODE records A, B of the same handler.  Payload words: c0_A, c0_B, then (c_Aj, c_Bj) per slot.
        A accumulates on %u0 / %u1 and updates the parent through %vc, B on %u2 / %u3 through %vd; the two value FMAs of
        a slot are adjacent and share x_j (operand reuse), the gradient FMAs of a group run A-then-B so that each
        adjoint stays in the reuse cache."""
        w(f"{label}:")
        n = len(m)
        nw = 2 + 2 * n                                       # payload words
        nwords = rec_words(nw)
        head(nwords)
        w("and.b32 %ka, %k, 255; shr.u32 %kb, %k, 8;")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addr];")
        w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sx; ld.shared.f64 %xkb, [%addrb];")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sg; ld.shared.f64 %gkb, [%addrb];")
        for i in range(NPRE, nw, 2):
            w(f"ld.shared.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [%sp+{8 + 8 * i}];")
        w("mul.rn.f64 %a0, %ac, %xk0;")
        w("mul.rn.f64 %a1, %ac, %xkb;")
        started = set()
        pending = {"%u0": "%n0", "%u2": "%n1"}               # node constants: addends of the first FMA of their chains
        if n < 2:                                            # (chains %u0 / %u2 get no FMA: plain moves, before %n0 / %n1
            w("mov.f64 %u0, %n0; mov.f64 %u2, %n1;")         #  are refilled with the next record's words)
            started |= {"%u0", "%u2"}
            pending = {}
        done = 0
        for q in range(0, n, group):
            js = list(range(q, min(q + group, n)))
            for j in js:                                     # value FMAs: A and B of one slot adjacent (x_j reused)
                for s, accs in ((0, ("%u0", "%u1")), (1, ("%u2", "%u3"))):
                    a, c = accs[(j + 1) % 2], coef(2 + 2 * j + s)
                    if a in started:
                        w(f"fma.rn.f64 {a}, {c}, %x{m[j]}, {a};")
                    elif a in pending:
                        w(f"fma.rn.f64 {a}, {c}, %x{m[j]}, {pending.pop(a)};")
                        started.add(a)
                    else:
                        w(f"mul.rn.f64 {a}, {c}, %x{m[j]};")
                        started.add(a)
            for s, adj in ((0, "%a0"), (1, "%a1")):          # gradient FMAs of the group: all of A, then all of B
                for j in js:
                    w(f"fma.rn.f64 %g{m[j]}, {coef(2 + 2 * j + s)}, {adj}, %g{m[j]};")
            # payload words < 2 + 2 * (js[-1] + 1) are dead now: refill their registers with the next record's words
            dead = 2 + 2 * (js[-1] + 1)
            while done + 2 <= min(dead, NPRE):
                w(f"ld.shared.v2.f64 {{%n{done}, %n{done + 1}}}, [%sp+{8 * nwords + 8 + 8 * done}];")
                done += 2
        preload(8 * nwords, done)
        assert not pending
        if "%u1" in started:
            w("add.rn.f64 %u0, %u0, %u1;")
        if "%u3" in started:
            w("add.rn.f64 %u2, %u2, %u3;")
        w("fma.rn.f64 %vc, %xk0, %u0, %vc;")
        w("fma.rn.f64 %vd, %xkb, %u2, %vd;")
        w("fma.rn.f64 %gk, %ac, %u0, %gk;")
        w("fma.rn.f64 %gkb, %ac, %u2, %gkb;")
        w("st.shared.f64 [%addr], %gk;")
        w("st.shared.f64 [%addrb], %gkb;")
        tail(nwords, NPRE)

    for mi, m in enumerate(masks):
        block_handler(f"H{OP_MASK + mi}", m, False)
    for mi, m in enumerate(masks):
        op = OP_MASK + nm + mi
        copies = fall_copies if op in fall else 1
        # back-to-back copies: every copy but the last falls through into the next one when the next record has the
        # same opcode (runs of equal records are the rule: symmetry-related siblings)
        for c in range(copies):
            fall_state[0] = op if c + 1 < copies else None
            block_handler(f"H{op}" + ("" if c == 0 else "BCDEFGH"[c - 1]), m, True)
    # ---- TRANSITION_p: scatter the register gradients of phase p to every member of their orbit (shared-memory
    # accumulators), then load the orbit sums of phase p + 1 and clear the register gradients
    for p, orbs in enumerate(phases):
        w(f"H{t_base + p}:")
        head(2)
        for o in orbs:
            for i in o:
                w(f"ld.shared.f64 %gk, [%sg+{i * NT * 8}]; add.rn.f64 %gk, %gk, %g{o[0]}; st.shared.f64 [%sg+{i * NT * 8}], %gk;")
        if p + 1 < len(phases):
            for o in phases[p + 1]:
                w(f"ld.shared.f64 %x{o[0]}, [%sx+{o[0] * NT * 8}]; mov.f64 %g{o[0]}, 0d0000000000000000;")
                for i in o[1:]:
                    w(f"ld.shared.f64 %w, [%sx+{i * NT * 8}]; add.rn.f64 %x{o[0]}, %x{o[0]}, %w;")
        tail(2)
    # ---- SIBLING_d: LEAVE_d k_old + ENTER_d k_new, c0 in one record (k = k_new | k_old << 8)
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{s_base + d - 1}:")
        head(2)
        w("shr.u32 %ka, %k, 8; and.b32 %kb, %k, 255;")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sx; ld.shared.f64 %xkb, [%addrb];")
        xold = f"%xk{d}"
        if slim:
            xold = "%xk0"
            w(f"mad.lo.u32 %addrb, %ka, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addrb];")
        w("add.rn.f64 %w, %vc, %vd;")
        w("mov.f64 %vc, %n0; mov.f64 %vd, 0d0000000000000000;")
        preload(16)
        w(f"fma.rn.f64 %vs{d}, {xold}, %w, %vs{d};")          # the parent's value, still saved
        w(f"fma.rn.f64 %gk, %as{d}, %w, %gk; st.shared.f64 [%addr], %gk;")
        if slim:
            w(f"mul.rn.f64 %ac, %as{d}, %xkb;")
        else:
            w(f"mov.f64 %xk{d}, %xkb; mul.rn.f64 %ac, %as{d}, %xkb;")
        tail(2, NPRE)
    # ---- DUAL_i
    for di, h in enumerate(dual):
        dual_handler(f"H{d_base + di}", masks[h])
    # ---- epilogue: the last TRANSITION has scattered every register gradient
    w("FINISH:")
    w("@!%valid bra DONE;")
    w("add.rn.f64 %vc, %vc, %vd;")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off;")
    w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %vc, %vc, %w;")
    w("st.global.f64 [%po], %vc; add.u64 %po, %po, %ldb;")
    for i in range(nv):
        w(f"ld.shared.f64 %gk, [%sg+{i * NT * 8}];")
        w(f"@%pscale ld.shared.f64 %w, [%sx+{i * NT * 8}]; @%pscale mul.rn.f64 %gk, %gk, %w;")
        w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %gk, %gk, %w;")
        w("st.global.f64 [%po], %gk; add.u64 %po, %po, %ldb;")
    w("DONE:")
    w("ret;")
    w("}")
    return out, SB + 2 * CH


def count_fp64(recs, masks, phases, dual):
    nm = len(masks)
    t_base = OP_MASK + 2 * nm
    s_base = t_base + len(phases)
    d_base = s_base + MAX_DEPTH
    n_fp = 0
    hist = collections.Counter()
    for op, k, payload in recs:
        if op >= d_base:
            n_fp += 2 * (2 * len(masks[dual[op - d_base]]) + 4)
            hist["dual"] += 1
        elif op >= s_base:
            n_fp += 4
            hist["sibling"] += 1
        elif op >= t_base:
            p = op - t_base
            n_fp += sum(len(o) for o in phases[p])
            if p + 1 < len(phases):
                n_fp += sum(len(o) - 1 for o in phases[p + 1])
            hist["transition"] += 1
        elif op >= OP_MASK + nm:
            n_fp += 2 * len(masks[op - OP_MASK - nm]) + 4
            hist["leafnode"] += 1
        elif op >= OP_MASK:
            n_fp += 2 * len(masks[op - OP_MASK])
            hist["mask"] += 1
        elif op == OP_NEXT:
            pass
        elif op >= OP_LEAVE:
            n_fp += 3
            hist["leave"] += 1
        elif op >= OP_ENTER:
            n_fp += 1
            hist["enter"] += 1
    return n_fp, hist


def generate(name, out_dir, block=256, min_blocks=1, piece=64, sym=32, skeleton_weight=40, fall=4, npre=8, fall_copies=2,
             dual_max=0, group=4, maxnreg=0, slim=False, anchors=False, basis=False):
    """Write {name}_eval.ptx and {name}_meta.json.  Returns (meta, words, masks, phases, dual)."""
    tab = dict(np.load(os.path.join(DATA, name + ".npz")))
    nv = int(tab["n_vars"])
    assert nv == NV[name]
    chunk = block * piece
    recs, masks, phases, dual = serialise(name, tab, sym, skeleton_weight, dual_max, basis)
    nm = len(masks)
    runs = collections.Counter(a[0] for a, b in zip(recs, recs[1:]) if a[0] == b[0])
    lo, hi = OP_MASK + nm, OP_MASK + 2 * nm
    fall_ops = tuple(sorted(op for op, c in runs.items() if fall and c >= fall and lo <= op < hi))
    words = pack(recs, chunk)
    n_fp, hist = count_fp64(recs, masks, phases, dual)
    symbol = f"{name}_stream0"
    lines = ["// GENERATED by tools/gen_interp.py -- command-stream adjoint-Horner PIP evaluator (see the tool's docstring).",
             ".version 8.7", ".target sm_100a", ".address_size 64", "", ".extern .shared .align 16 .b8 pip_smem[];", "",
             f".global .align 16 .b64 {symbol}[{len(words)}] = {{"]
    for i in range(0, len(words), 8):
        lines.append(", ".join("0x%016X" % int(v) for v in words[i:i + 8]) + ("," if i + 8 < len(words) else ""))
    lines.append("};")
    entry = f"{name}_eval"
    kl, smem = kernel_ptx(name, entry, symbol, masks, nv, phases, dual, block, min_blocks, piece, fall_ops, npre,
                          fall_copies, group, maxnreg, slim, anchors)
    lines += [""] + kl
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"{name}_eval.ptx"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    meta = {"name": name, "block": block, "smem_bytes": smem, "fp64_instr": n_fp, "chunk_bytes": chunk,
            "records": dict(hist), "kind": "interp-sym",
            "kernels": [{"entry": entry, "vars": nv, "masks": nm, "stream_bytes": int(words.nbytes), "fp64_instr": n_fp}],
            "phases": phases, "handlers": [len(m) for m in masks], "dual": list(dual)}
    with open(os.path.join(out_dir, f"{name}_meta.json"), "w") as fh:
        json.dump(meta, fh)
    return meta, words, masks, phases, dual


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GEN)
    ap.add_argument("--only", default="", help="pip2b or pip3b (default: both)")
    for two in ("", "2"):
        d = {"": (256, 64, 32, 8, 0), "2": (384, 32, 12, 4, 0)}[two]
        ap.add_argument(f"--block{two}", type=int, default=d[0], help="threads per CTA" + (" (2B)" if two else " (3B)"))
        ap.add_argument(f"--piece{two}", type=int, default=d[1], help="stream bytes each thread stages per chunk")
        ap.add_argument(f"--sym{two}", type=int, default=d[2], help="size of the leaf-handler library")
        ap.add_argument(f"--npre{two}", type=int, default=d[3], help="payload words of the next record every handler pre-loads")
        ap.add_argument(f"--dual{two}", type=int, default=d[4], help="DUAL records for handlers of at most N slots")
        ap.add_argument(f"--maxnreg{two}", type=int, default=0, help="register cap (0: none); see kernel_ptx")
        ap.add_argument(f"--slim{two}", type=int, default=0, help="1: saved tree levels keep 2 doubles instead of 4")
        ap.add_argument(f"--anchors{two}", type=int, default=0, help="1: uniform anchors in the rare handlers (see tail())")
        ap.add_argument(f"--basis{two}", type=int, default=0, help="1: one phase in the symmetry-adapted basis (adapted_basis)")
    ap.add_argument("--skeleton-weight", type=int, default=40)
    ap.add_argument("--fall", type=int, default=4, help="duplicate leaf handlers that follow themselves at least N times")
    ap.add_argument("--fall-copies", type=int, default=2)
    ap.add_argument("--group", type=int, default=4, help="slots per value / gradient FMA group inside a handler")
    ap.add_argument("--min-blocks", type=int, default=1)
    args = ap.parse_args()
    for name in ("pip2b", "pip3b"):
        if args.only and args.only != name:
            continue
        t = "2" if name == "pip2b" else ""
        meta, *_ = generate(name, args.out, getattr(args, "block" + t), args.min_blocks, getattr(args, "piece" + t),
                            getattr(args, "sym" + t), args.skeleton_weight, args.fall, getattr(args, "npre" + t),
                            args.fall_copies, getattr(args, "dual" + t), args.group, getattr(args, "maxnreg" + t),
                            bool(getattr(args, "slim" + t)), bool(getattr(args, "anchors" + t)),
                            bool(getattr(args, "basis" + t)))
        print({k: v for k, v in meta.items() if k != "phases"})


if __name__ == "__main__":
    main()
