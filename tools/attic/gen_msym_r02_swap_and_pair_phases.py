"""ATTIC (round 2, measured and rejected): gen_msym.py with its two multi-phase modes,
    --mode swap  : nodes invariant under hydrogen swap a as orbit sums, contexts {e, (bc)} x 3 passes      (2.19 ms vs 1.72)
    --mode pairs : a second phase whose 36 registers hold the 21 swap pair sums + 15 singles, 6 contexts   (1.91 ms vs 1.72)
(profiles/r02_microbench_msym_variants.log).  The shipped generator keeps `single` and `basis`.

Code generator, second generation: MOLECULE-SYMMETRISED adjoint-Horner evaluator (command stream + sm_100a PTX kernel).

What tools/gen_interp.py (round 1 / 2) does per tuple and where it stops.  One thread evaluates one tuple; every leaf
coefficient of the Horner tree is a warp-uniform double that has to reach the FMA through a vector register: one
broadcast ld.shared wavefront (256 B of register write-back) per coefficient, for TWO FMAs.  ncu of the shipped 3B kernel
(profiles/r02_ncu_pip3b_eval_final_details.txt): LSU data pipe 88 % busy, FP64 pipe 56 %, 2 warps per scheduler
waiting on two dependent value chains.  The kernel is bound by coefficient DELIVERY, not by arithmetic.

The polynomial is invariant under relabelling the molecules (3B: S3, 6 elements; 2B: the exchange a <-> b).  Hence

    P(x) = sum_{sigma in S} Q(sigma x),     Q = one representative monomial per S-orbit (coefficient / stabiliser order),

and the SAME coefficient serves |S| evaluations ("contexts") of the SAME tuple: context c reads x_{sigma_c(j)} where the
identity context reads x_j -- another register of the same x array -- and adds into g_{sigma_c(j)}.  One coefficient load
now feeds 2 |S| FMAs on |S| independent value chains; headers, dispatches and address arithmetic are paid once per |S|
nodes; the stream shrinks |S|-fold.  The per-node shared-memory traffic (x_k, g_k of the node variable) stays.

Choice of representatives.  Any choice is exact; the number of tree nodes and of leaf slots depends on it.  The tree is
built by the greedy Horner split WITH IMAGE FREEDOM: an orbit may follow any of its images, so whole orbits pile up under
the most popular prefixes (3B: 365 leaf nodes / 5,326 slots for 5,657 orbits instead of 610 / 5,121 with fixed
lexicographic representatives).

Swap phases.  Nodes whose leaf coefficients are constant on the orbits of a hydrogen swap are linear forms in orbit sums
(gen_interp.py, symmetry phases).  With molecule contexts only the contexts that normalise the swap can share the orbit-sum
registers (3B, swap a: e and (bc)); the other cosets are separate PASSES over the same nodes with the registers loaded in
permuted order.  A phase is therefore (swap set S, pass permutation tau, context list), see `Phase`.

Records (all carry the opcode of their successor in the header, like gen_interp.py):
    LEAF_h   k[C], c0, coefs : for every context c:  a' = a_c x_{k_c};  u = c0 + sum_j coef_j R_{pi_c(j)};
                               G_{pi_c(j)} += coef_j a';  v_c += x_{k_c} u;  g_{k_c} += a_c u
    ENTER_d  k[C]            : (d = 2: push v_c, a_c)  a_c *= x_{k_c} (d = 1: a_c = x_{k_c}; root partial sums -> vr);  v_c = 0
    LEAVE_d  k[C]            : w = v_c; (d = 2: pop v_c, a_c; d = 1: v_c = 0, a_c = 1)  v_c += x_{k_c} w;  g_{k_c} += a_c w
    TRANS_p                  : scatter the register gradients of phase p, load the registers of phase p + 1
    NEXT / END
An inner node's own constant and leaves are emitted as a LEAF record with the node's variable in front of its ENTER
(x_k (c0 + L + kids) = x_k (c0 + L) + x_k kids), so there is no separate family of MASK handlers.
"""
from __future__ import annotations

import argparse
import collections
import itertools
import json
import os
import struct

import numpy as np

import gen_interp as gi

PAD = 255
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = gi.DATA
GEN = gi.GEN
NV = gi.NV

OP_END, OP_NEXT, OP_SKEL = 0, 1, 2
# per phase class cl: ENTER_1, ENTER_2, LEAVE_1, LEAVE_2 = OP_SKEL + 4 cl + 0..3; then TRANS_p; then LEAF_h
MAX_ENTER_DEPTH = 2


# ---- symmetry -----------------------------------------------------------------------------------------------------------
def mol_perms(name):
    """variable permutations induced by relabelling the molecules: {label tuple: perm}, identity first"""
    sites, _ = gi._sites(name)
    vid = {frozenset(p): i for i, p in enumerate(sites)}
    mols = "abc" if name == "pip3b" else "ab"
    out = collections.OrderedDict()
    for pm in itertools.permutations(mols):
        mp = dict(zip(mols, pm))

        def f(s):
            return s[0] + mp[s[1]] + s[2:]
        out[pm] = [vid[frozenset((f(p), f(q)))] for p, q in sites]
    return out


def img(perm, m):
    return tuple(sorted(perm[v] for v in m))


def compose(p, q):
    """(p o q)[v] = p[q[v]]"""
    return [p[v] for v in q]


class Node:
    __slots__ = ("var", "const", "leaves", "kids")

    def __init__(self, var):
        self.var, self.const, self.leaves, self.kids = var, None, {}, []


def s_orbits(monos, coef, perms):
    """-> [(set of the distinct images of one S-orbit of monomials, coefficient value / stabiliser order)]"""
    out, seen = [], set()
    for m, c in monos.items():
        if m in seen:
            continue
        orb = collections.Counter(img(p, m) for p in perms)
        seen |= set(orb)
        stab = next(iter(orb.values()))
        assert all(v == stab for v in orb.values())
        out.append((set(orb), float(coef[c]) / stab))
    return out


def build_free(orbs, var, score):
    """Greedy Horner split with image freedom.  orbs: [(set of candidate residual monomials, value)]; the variable that
    can divide most orbits is pulled out, every orbit it can divide follows it with the candidates that contain it."""
    node = Node(var)
    rest = []
    for cands, val in orbs:
        if () in cands:
            assert node.const is None
            node.const = val
        else:
            rest.append((cands, val))
    while rest:
        cnt = collections.Counter()
        for cands, _ in rest:
            for v in {v for c in cands for v in c}:
                cnt[v] += 1
        best = max(cnt.items(), key=lambda kv: (kv[1], score(kv[0]), -kv[0]))[0]
        sub, other = [], []
        for cands, val in rest:
            nc = set()
            for c in cands:
                if best in c:
                    lst = list(c)
                    lst.remove(best)
                    nc.add(tuple(lst))
            (sub if nc else other).append((nc, val) if nc else (cands, val))
        rest = other
        if len(sub) == 1 and () in sub[0][0]:
            node.leaves[best] = sub[0][1]
        else:
            node.kids.append(build_free(sub, best, score))
    return node


def expand(node, prefix=()):
    """the polynomial a tree stands for: {monomial: value}"""
    out = {}
    p = prefix + ((node.var,) if node.var is not None else ())
    if node.const is not None:
        out[tuple(sorted(p))] = node.const
    for v, val in node.leaves.items():
        out[tuple(sorted(p + (v,)))] = val
    for k in node.kids:
        for m, val in expand(k, p).items():
            assert m not in out
            out[m] = val
    return out


class Phase:
    """slot family, pass permutation tau, context permutations rho_c (identity first).
    Register slot `sid[q]` holds sum_{i in orbs[q]} x[tau[i]]; context c reads slot pi_c(j) = the slot whose member set
    is rho_c(members of j) -- the family must be closed under the contexts -- and its node variable is tau[rho_c[k]].
    Member sets may overlap (a variable sits in the pair sums of two swaps): the gradient of a slot goes to all its
    members, the sums over slots add up."""

    def __init__(self, nv, slots, tau, ctx, cls):
        self.tau, self.ctx, self.cls = tau, ctx, cls
        self.sid = [j for j, _ in slots]
        self.orbs = [list(m) for _, m in slots]
        key = {frozenset(m): j for j, m in slots}
        assert len(key) == len(slots) and len(set(self.sid)) == len(slots) and max(self.sid) < nv
        self.pi = [{j: key[frozenset(rho[i] for i in m)] for j, m in slots} for rho in ctx]
        self.kmap = [compose(tau, rho) for rho in ctx]


PAIR_SINGLES = {"pip3b": [21, 23, 25, 27, 29, 31, 22, 24, 26, 28, 30, 32, 33, 34, 35]}     # O-H and O-O between molecules


def pair_slots(name, nv, swaps):
    """the mixed family of `pairs` mode: the 21 hydrogen-swap pair sums x_v + x_swap(v) of all three swaps (closed under
    the molecule relabellings, so ALL six contexts can use them) + the singles that leaf nodes with pairs need most,
    filling the same 36 registers"""
    pairs = sorted({tuple(sorted((v, s[v]))) for s in swaps for v in range(nv) if s[v] != v})
    members = [list(p) for p in pairs] + [[v] for v in PAIR_SINGLES[name]]
    assert len(members) <= nv
    return [(q, m) for q, m in enumerate(members)]


def adapted_basis(name, nv, swaps, keep="2"):
    """Symmetry-adapted slot family of `basis` mode: an invertible 0/1 change of basis of the nv variables, block by block
    over the orbits of the hydrogen-swap group, closed under the molecule relabellings:
        orbit {v}                      -> v
        orbit {v1, v2} (one swap)      -> v1 + v2,  v2      (O-H inside a molecule; O-H between molecules stays v1, v2:
                                                              measured, the change of basis costs slots there)
        orbit {w11, w12, w21, w22}     -> sum of all four,  w21 + w22,  w12 + w22,  w22       (H-H between two molecules)
    A node's linear form sum_v c_v x_v is rewritten in this basis (the inverse has entries 0, +-1, so the new coefficients
    are sums / differences of the old ones): leaves that are symmetric under a swap collapse onto the sum slots -- the
    saving of the orbit-sum phases of gen_interp.py, for all three swaps at once, in the SAME 36 registers and for all six
    contexts; a lone leaf on a "first" member costs up to four slots instead of one.  3B: 4,166 slots instead of 5,326
    (all orientations of the three block types were tried: 4,166 ... 5,452).
    The register of a slot holds the sum of its members; its gradient is added to every member at the end."""
    rows = gi.adapted_basis(name, nv, swaps, {(2, ("Hi", "Oi")): keep, (4, ("Hx", "Hx")): keep})
    return [(q, m) for q, m in enumerate(rows)]


def symmetrise(name, mode="single"):
    """-> (root of the tree of Q, nv, phases, molecule permutations).
    mode single: one phase, 36 single-variable slots, all molecule permutations as contexts.
         swap  : + class 1 = orbit sums of swap a, contexts e and (bc), three passes (measured: slower; kept for the record)
         pairs : + class 1 = pair sums of all swaps + 15 singles, all contexts (see pair_slots)"""
    tab = dict(np.load(os.path.join(DATA, name + ".npz")))
    nv = int(tab["n_vars"])
    coef = tab["coef"]
    monos = {tuple(sorted(int(v) for v in row if v != PAD)): int(c) for row, c in zip(tab["mono_vars"], tab["mono_coef"])}
    mp = mol_perms(name)
    perms = list(mp.values())
    for p in perms:                                     # the premise: invariance under every molecule relabelling
        assert all(monos.get(img(p, m)) == c for m, c in monos.items())
    swaps = gi.swap_perms(name)
    sites, _ = gi._sites(name)
    ha = [any(s.startswith("Ha") for s in p) for p in sites]
    orbs = s_orbits(monos, coef, perms)
    root = build_free(orbs, None, (lambda v: 0 if ha[v] else 1) if mode == "swap" else (lambda v: 0))
    # exactness of the representation: sum over the images of Q == P, term by term
    acc = collections.defaultdict(float)
    for m, val in expand(root).items():
        for p in perms:
            acc[img(p, m)] += val
    assert set(acc) == set(monos)
    assert max(abs(acc[m] - coef[c]) / abs(coef[c]) for m, c in monos.items()) < 1e-15
    ident = list(range(nv))
    if mode.startswith("basis"):
        phases = [Phase(nv, adapted_basis(name, nv, swaps, mode[5:] or "2"), ident, perms, 0)]
        return root, nv, phases, perms
    phases = [Phase(nv, [(v, [v]) for v in range(nv)], ident, perms, 0)]
    if mode == "swap":
        assert name == "pip3b"
        bc, ab, ac = mp[("a", "c", "b")], mp[("b", "a", "c")], mp[("c", "b", "a")]
        slots = [(o[0], o) for o in gi._orbits(nv, [swaps[0]])]
        for tau in (ident, ab, ac):
            phases.append(Phase(nv, slots, tau, [ident, bc], 1))
        cover = sorted(tuple(compose(ph.tau, rho)) for ph in phases[1:] for rho in ph.ctx)
        assert cover == sorted(tuple(p) for p in perms)
    elif mode == "pairs":
        phases.append(Phase(nv, pair_slots(name, nv, swaps), ident, perms, 1))
    else:
        assert mode == "single"
    return root, nv, phases, perms


def classify(root, phases, mode, overhead=3):
    """-> {id(node): [(class, need = {slot: value}, constant or None), ...]} for every node with leaves or a constant.
    single: everything in class 0.  swap: the class that needs fewest slots.  pairs: a node is evaluated in class 0, in
    class 1 or SPLIT between them (its linear form is a sum of two), whatever costs least with `overhead` slots per record."""
    by_cls = {}
    for ph in phases:
        by_cls.setdefault(ph.cls, ph)
    out = {}

    def express(n, ph):
        """greedy cover of the leaves with the slots of ph: pair slots where both members carry the same coefficient,
        single slots otherwise; -> (need, leaves left over)"""
        left, need = dict(n.leaves), {}
        for j, m in zip(ph.sid, ph.orbs):
            if len(m) > 1 and all(v in left for v in m) and len({left[v] for v in m}) == 1:
                need[j] = left[m[0]]
                for v in m:
                    del left[v]
        for j, m in zip(ph.sid, ph.orbs):
            if len(m) == 1 and m[0] in left:
                need[j] = left.pop(m[0])
        return need, left

    if mode.startswith("basis"):
        ph = by_cls[0]
        nv = len(ph.sid)
        B = np.zeros((nv, nv))
        for q, m in zip(ph.sid, ph.orbs):
            B[q, m] = 1.0
        BinvT = np.rint(np.linalg.inv(B)).T                # entries 0, +-1: sum_v c_v x_v = sum_q (B^-T c)_q (B x)_q
        assert np.array_equal(BinvT.T @ B, np.eye(nv))

    def walk(n):
        if n.leaves or n.const is not None:
            full = {v: c for v, c in n.leaves.items()}
            if mode.startswith("basis"):
                c = np.zeros(nv)
                for v, val in n.leaves.items():
                    c[v] = val
                cp = BinvT @ c
                out[id(n)] = [(0, {q: float(cp[q]) for q in range(nv) if cp[q] != 0.0}, n.const)]
            elif mode == "single":
                out[id(n)] = [(0, full, n.const)]
            elif mode == "swap":
                best = (0, full)
                need, left = express(n, by_cls[1])
                if not left and len(need) < len(full):
                    best = (1, need)
                out[id(n)] = [(best[0], best[1], n.const)]
            else:
                need, left = express(n, by_cls[1])
                c0 = len(full) + overhead
                c1 = len(need) + overhead if not left else 1 << 30
                c2 = len(need) + len(left) + 2 * overhead if need and left else 1 << 30
                if c0 <= min(c1, c2):
                    out[id(n)] = [(0, full, n.const)]
                elif c1 <= c2:
                    out[id(n)] = [(1, need, n.const)]
                else:
                    out[id(n)] = [(1, need, n.const), (0, left, None)]
        for k in n.kids:
            walk(k)
    walk(root)
    return out


# ---- serialisation ------------------------------------------------------------------------------------------------------
def op_layout(phases, handlers=()):
    ncls = 1 + max(ph.cls for ph in phases)
    op_trans = OP_SKEL + 4 * ncls
    op_leaf0 = op_trans + len(phases)
    return ncls, op_trans, op_leaf0, op_leaf0 + len(handlers)


def serialise(name, K=(12, 12), mode="single", overhead=3):
    """-> (records, handlers, phases, nv).  records: (op, [node variable per context], payload doubles, [parent's node
    variables: LEAVE_2 only]); handlers: list of (class, slot tuple) in opcode order (LEAF_h = op_leaf0 + h)."""
    root, nv, phases, perms = symmetrise(name, mode)
    cls = classify(root, phases, mode, overhead)
    ncls, op_trans, op_leaf0, _ = op_layout(phases)
    bits = lambda need: sum(1 << j for j in need)  # noqa: E731
    libs, pick = [], []
    for c in range(ncls):
        req = collections.Counter(bits(need) for parts in cls.values() for cc, need, _ in parts if cc == c)
        lib = gi.merge_library(req, K[c]) if req else []
        pk = {}
        for b in req:
            pk[b] = min((h for h in lib if h & b == b), key=lambda h: (bin(h).count("1"), h))
        used = collections.Counter()
        for b, f in req.items():
            used[pk[b]] += f
        libs.append([h for h, _ in used.most_common()])
        pick.append(pk)
    keys = [(c, h) for c in range(ncls) for h in libs[c]]
    handlers = [(c, tuple(j for j in range(nv) if h >> j & 1)) for c, h in keys]
    hid = {k: i for i, k in enumerate(keys)}
    has = {}

    def mark(n):
        s = {c for c, _, _ in cls.get(id(n), [])}
        for k in n.kids:
            s |= mark(k)
        has[id(n)] = s
        return s
    mark(root)
    recs = []
    for p, ph in enumerate(phases):
        def ks(var):
            return [km[var] for km in ph.kmap]

        def own_part(n):
            for part in cls.get(id(n), []):
                if part[0] == ph.cls:
                    return part
            return None

        def emit(n, depth, parent=None):
            part = own_part(n)
            kids = [k for k in n.kids if ph.cls in has[id(k)]]
            if depth == 0:
                assert id(n) not in cls, "constants / leaves at the root: not needed by pip3b"
            else:
                if part:
                    c, need, const = part
                    i = hid[(c, pick[c][bits(need)])]
                    pl = [const if const is not None else 0.0] + [need.get(j, 0.0) for j in handlers[i][1]]
                    recs.append((op_leaf0 + i, ks(n.var), pl, []))
                if not kids:
                    return
                assert depth <= MAX_ENTER_DEPTH
                recs.append((OP_SKEL + 4 * ph.cls + depth - 1, ks(n.var), [], []))

            def order(k):           # childless kids grouped by handler: equal records adjacent
                part = own_part(k)
                if part and not any(ph.cls in has[id(q)] for q in k.kids):
                    return (0, hid[(part[0], pick[part[0]][bits(part[1])])])
                return (1, 0)
            for k in sorted(kids, key=order):
                emit(k, depth + 1, n)
            if depth > 0:       # LEAVE_2 re-creates the parent's adjoints x_{k1}: it carries the parent's variables too
                recs.append((OP_SKEL + 4 * ph.cls + 2 + depth - 1, ks(n.var), [], ks(parent.var) if depth == 2 else []))
        if ph.cls in has[id(root)]:
            emit(root, 0)
        recs.append((op_trans + p, [], [], []))
    recs.append((OP_END, [], [], []))
    return recs, handlers, phases, nv


def node_vars(recs):
    """the variables that are addressed dynamically (x_k loads, g_k updates): only these need shared-memory copies"""
    return sorted({k for _, ks, _, pk in recs for k in ks + pk})


def hdr_words(cmax):
    """header: word 0 = successor opcode; then the node-variable byte offsets of the contexts, two u32 per word; padded
    to an even number of words so that payloads are 16-byte aligned"""
    h = 1 + (cmax + 1) // 2
    return h + (h & 1)


def rec_words(n_payload, cmax):
    return hdr_words(cmax) + n_payload + (n_payload & 1)


def pack(recs, chunk_bytes, cmax, nt, npre):
    """u64 words in chunks of chunk_bytes (staged through shared memory one at a time).  Words 0, 1 of a chunk are
    padding (word 0 of the stream: the first opcode); records are 16-byte aligned and never straddle a chunk, the last
    record of a chunk is NEXT; a trailing all-zero chunk keeps the look-ahead loads inside the array.  A node variable
    is stored as the byte offset of its row in the compact shared-memory copies (row = rank among node_vars)."""
    cw = chunk_bytes // 8
    H = hdr_words(cmax)
    kidx = {k: i for i, k in enumerate(node_vars(recs))}
    chunks, cur, hdrs = [], [0, 0], []

    def header(op, ks, pks=()):
        """header words; LEAVE_2: the parent's node variables follow as H - 1 extra words + 1 word of padding"""
        w = [op] + [0] * (H - 1 + (H if pks else 0))
        for c, k in enumerate(ks):
            w[1 + c // 2] |= (kidx[k] * nt * 8) << (32 * (c % 2))
        for c, k in enumerate(pks):
            w[H + c // 2] |= (kidx[k] * nt * 8) << (32 * (c % 2))
        return w
    for op, ks, payload, pks in recs:
        need = len(header(op, ks, pks)) + len(payload) + (len(payload) & 1)
        if len(cur) + need + 2 * H + npre + 2 > cw:       # room for NEXT + the look-ahead words
            hdrs.append((len(chunks), len(cur)))
            cur += header(OP_NEXT, [])
            cur += [0] * (cw - len(cur))
            chunks.append(cur)
            cur = [0, 0]
        hdrs.append((len(chunks), len(cur)))
        cur += header(op, ks, pks)
        p = list(payload) + ([0.0] if len(payload) & 1 else [])
        cur.extend(struct.unpack("<Q", struct.pack("<d", float(c)))[0] for c in p)
    cur += [0] * (cw - len(cur))
    chunks.append(cur)
    chunks.append([0] * cw)
    ops = [chunks[c][i] & 0xFFFFFFFF for c, i in hdrs]
    chunks[0][0] = ops[0]
    for (c, i), nxt in zip(hdrs, ops[1:] + [OP_END]):     # look-ahead: header word 0 = opcode of the NEXT record
        chunks[c][i] = nxt
    return np.array([w for c in chunks for w in c], dtype=np.uint64)


def interpret(name, words, x, chunk_bytes, handlers, phases, nt, kvars):
    """numpy / python interpreter of the packed stream: the records, in order, with the meaning of the PTX handlers"""
    nv = NV[name]
    cmax = max(len(ph.ctx) for ph in phases)
    H = hdr_words(cmax)
    cw = chunk_bytes // 8
    x = np.asarray(x, dtype=np.float64)
    f = words.view(np.float64)
    g = np.zeros(nv)
    ncls, op_trans, op_leaf0, _ = op_layout(phases, handlers)

    def load_phase(ph):
        return {j: sum(x[ph.tau[i]] for i in o) for j, o in zip(ph.sid, ph.orbs)}, {j: 0.0 for j in ph.sid}
    p = 0
    ph = phases[0]
    R, Gr = load_phase(ph)
    C = len(ph.ctx)
    v, a = [0.0] * C, [1.0] * C
    vr = 0.0
    stack = None
    pc = 2
    op = int(words[0]) & 0xFFFFFFFF
    while True:
        nxt = int(words[pc]) & 0xFFFFFFFF
        ks = [kvars[(int(words[pc + 1 + c // 2]) >> (32 * (c % 2)) & 0xFFFFFFFF) // (nt * 8)] for c in range(C)]
        if op == OP_END:
            return vr + sum(v), g
        if op == OP_NEXT:
            pc = (pc // cw + 1) * cw + 2
        elif op < op_trans:
            cl, kind = divmod(op - OP_SKEL, 4)
            assert cl == ph.cls
            if kind < 2:                                     # ENTER_1, ENTER_2
                if kind == 0:
                    vr += sum(v)
                    a = [x[k] for k in ks]
                else:
                    assert stack is None
                    stack = (list(v), list(a))
                    a = [a[c] * x[ks[c]] for c in range(C)]
                v = [0.0] * C
            else:                                            # LEAVE_1, LEAVE_2
                w = v
                if kind == 2:
                    v, a = [0.0] * C, [1.0] * C
                else:
                    v, a = stack
                    pk = [kvars[(int(words[pc + H + c // 2]) >> (32 * (c % 2)) & 0xFFFFFFFF) // (nt * 8)] for c in range(C)]
                    assert a == [x[k] for k in pk]               # what the kernel re-creates the adjoints from
                    stack = None
                for c in range(C):
                    v[c] += x[ks[c]] * w[c]
                    g[ks[c]] += a[c] * w[c]
            pc += 2 * H if kind == 3 else H
        elif op < op_leaf0:                                  # TRANS_p
            assert op - op_trans == p
            for j, o in zip(ph.sid, ph.orbs):
                for i in o:
                    g[ph.tau[i]] += Gr[j]
            vr += sum(v)
            v = [0.0] * C
            p += 1
            if p < len(phases):
                ph = phases[p]
                R, Gr = load_phase(ph)
                C = len(ph.ctx)
                v, a = [0.0] * C, [1.0] * C
            pc += H
        else:
            cl, slots = handlers[op - op_leaf0]
            assert cl == ph.cls
            for c in range(C):
                a2 = a[c] * x[ks[c]]
                u = f[pc + H]
                for q, j in enumerate(slots):
                    cf = f[pc + H + 1 + q]
                    u += cf * R[ph.pi[c][j]]
                    Gr[ph.pi[c][j]] += cf * a2
                v[c] += x[ks[c]] * u
                g[ks[c]] += a[c] * u
            pc += rec_words(1 + len(slots), cmax)
        op = nxt


# ---- PTX ----------------------------------------------------------------------------------------------------------------
def ctx_array(ph, perms_parity):
    """which of the two shared-memory g_k accumulator arrays a context adds into.  Two contexts of one record name the
    same node variable iff they differ by a relabelling that fixes it -- for every variable that is a transposition
    (3B) / the exchange (2B) -- so contexts of equal PARITY never collide and their read-modify-writes are independent."""
    return [perms_parity[tuple(compose(ph.tau, rho))] for rho in ph.ctx]


def perm_parity(p):
    seen, par = set(), 0
    for i in range(len(p)):
        if i in seen:
            continue
        j, n = i, 0
        while j not in seen:
            seen.add(j)
            j = p[j]
            n += 1
        par ^= (n - 1) & 1
    return par


def kernel_ptx(name, entry, stream, handlers, nv, phases, kvars, block=256, piece=64, npre=8, ahead=2, maxnreg=0, anchors=False):
    """{entry}(const double* x /*[nv][ld]*/, double* out /*[1+nv][ld]: E, dE/dx*/, u32 n, u32 ld, u32 mode)
    mode bit 0: add into out; bit 1: rows 1.. receive x_i dE/dx_i (see gen_interp.kernel_ptx).
    Returns (lines, dynamic shared bytes)."""
    NT = block
    CH = NT * piece
    cmax = max(len(ph.ctx) for ph in phases)
    H = hdr_words(cmax)
    HB = 8 * H
    nk = len(kvars)
    kidx = {k: i for i, k in enumerate(kvars)}
    XS = nk * NT * 8                       # smem: x rows of the node variables | g_k (even) | g_k (odd) | 2 stream buffers
    SGA = [XS, 2 * XS]
    SB = 3 * XS
    ncls, op_trans, op_leaf0, n_ops = op_layout(phases, handlers)
    by_cls = {}
    for ph in phases:
        by_cls.setdefault(ph.cls, ph)
    # molecule relabellings act on the SITES; their parity there (not on the variables) is what decides collisions:
    # sigma and sigma t (t a transposition of molecules) have opposite parity as permutations of the molecules
    mp = mol_perms(name)
    mols = "abc" if name == "pip3b" else "ab"
    parity = {tuple(p): perm_parity([mols.index(ch) for ch in lab]) for lab, p in mp.items()}
    for ph in phases:                       # the premise, checked: equal-parity contexts never share a node variable
        arr = ctx_array(ph, parity)
        for k in range(nv):
            for par in (0, 1):
                img_k = [ph.kmap[c][k] for c in range(len(ph.ctx)) if arr[c] == par]
                assert len(set(img_k)) == len(img_k)
    out = []
    w = out.append
    w(f".visible .entry {entry}(.param .u64 p_x, .param .u64 p_out, .param .u32 p_n, .param .u32 p_ld, .param .u32 p_acc)")
    w(f".maxntid {NT}, 1, 1")
    w(f".maxnreg {maxnreg}" if maxnreg else ".minnctapersm 1")
    w("{")
    w(f".reg .f64 %x<{nv}>, %g<{nv}>, %c<80>, %n<{npre + 2}>, %vr, %w, %gk, %one, %zero;")
    w(f".reg .f64 %v<{cmax}>, %a<{cmax}>, %u<{cmax}>, %b<{cmax}>, %xk<{cmax}>, %gq<{cmax}>, %vs<{cmax}>;")
    w(f".reg .u32 %kq<{2 * ((cmax + 1) // 2) + 2}>, %ad<{cmax}>;")
    w(".reg .u64 %px, %po, %off, %ldb, %gs, %pa; .reg .u32 %n, %ld, %t, %tc, %r<4>, %op, %nop, %spare, %sx, %addr;")
    w(".reg .u32 %sp, %bufcur, %bufoth, %mine; .reg .pred %valid, %pacc, %pscale, %pself; .reg .u32 %acc, %accb;")
    w("ld.param.u32 %acc, [p_acc]; and.b32 %accb, %acc, 1; setp.ne.u32 %pacc, %accb, 0;")
    w("and.b32 %accb, %acc, 2; setp.ne.u32 %pscale, %accb, 0;")
    w("ld.param.u64 %px, [p_x]; ld.param.u64 %po, [p_out]; ld.param.u32 %n, [p_n]; ld.param.u32 %ld, [p_ld];")
    w("cvta.to.global.u64 %px, %px; cvta.to.global.u64 %po, %po;")
    w("mov.u32 %r0, %ctaid.x; mov.u32 %r1, %ntid.x; mov.u32 %r2, %tid.x;")
    w("mad.lo.u32 %t, %r0, %r1, %r2;")
    w("setp.lt.u32 %valid, %t, %n;")
    w("sub.u32 %tc, %n, 1; min.u32 %tc, %tc, %t;")
    w("mul.wide.u32 %off, %tc, 8; add.u64 %px, %px, %off;")        # &x[0][tuple]  (kept: the phases load from here)
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off;")         # &out[0][tuple]
    w("mul.wide.u32 %ldb, %ld, 8;")
    w("mov.u32 %r3, pip_smem;")
    w("shl.b32 %sx, %r2, 3; add.u32 %sx, %sx, %r3;")               # &xs[0][tid]
    w(f"add.u32 %bufcur, %r3, {SB}; add.u32 %bufoth, %bufcur, {CH};")

    def async_fetch(buf):
        w(f"add.u32 %addr, {buf}, %mine;")
        for q in range(piece // 16):
            w(f"cp.async.ca.shared.global [%addr+{q * NT * 16}], [%gs+{q * NT * 16}], 16;")
    w("mul.lo.u32 %mine, %r2, 16;")
    w(f"mov.u64 %gs, {stream}; cvta.to.global.u64 %gs, %gs; cvt.u64.u32 %off, %mine; add.u64 %gs, %gs, %off;")
    async_fetch("%bufcur")
    w("cp.async.wait_all;")
    w(f"add.u64 %gs, %gs, {CH};")
    async_fetch("%bufoth")
    w("mov.f64 %one, 0d3FF0000000000000; mov.f64 %zero, 0d0000000000000000;")

    def xaddr(i):
        w(f"mad.lo.u64 %pa, %ldb, {i}, %px;")
    for k in kvars:                                               # node variables -> shared rows; accumulators = 0
        xaddr(k)
        w("ld.global.nc.f64 %w, [%pa];")
        r = kidx[k] * NT * 8
        w(f"st.shared.f64 [%sx+{r}], %w; st.shared.f64 [%sx+{SGA[0] + r}], %zero; st.shared.f64 [%sx+{SGA[1] + r}], %zero;")
    w("mov.f64 %vr, %zero;")

    def load_phase(ph, first=False):
        def ldx(dst, i):                 # the first phase loads from global memory, later ones from the shared copy
            if first:
                xaddr(ph.tau[i])
                w(f"ld.global.nc.f64 {dst}, [%pa];")
            else:
                w(f"ld.shared.f64 {dst}, [%sx+{kidx[ph.tau[i]] * NT * 8}];")
        for j, o in zip(ph.sid, ph.orbs):
            ldx(f"%x{j}", o[0])
            w(f"mov.f64 %g{j}, %zero;")
            for i in o[1:]:
                ldx("%w", i)
                w(f"add.rn.f64 %x{j}, %x{j}, %w;")
        for c in range(len(ph.ctx)):
            w(f"mov.f64 %v{c}, %zero; mov.f64 %a{c}, %one;")
    load_phase(phases[0], True)
    w("bar.sync 0;")
    w("add.u32 %sp, %bufcur, 16;")
    w("TBL: .branchtargets " + ", ".join(f"H{i}" for i in range(n_ops)) + ";")

    def load_hdr(byte_off, C):
        """header of the record at %sp+byte_off: successor opcode -> %nop, node-variable offsets -> %kq"""
        w(f"ld.shared.v2.u32 {{%nop, %spare}}, [%sp+{byte_off}];")
        q, nw = 1, (C + 1) // 2
        while nw > 0:
            k0 = 2 * (q - 1)
            if q % 2 == 0 and nw >= 2:
                w(f"ld.shared.v4.u32 {{%kq{k0}, %kq{k0 + 1}, %kq{k0 + 2}, %kq{k0 + 3}}}, [%sp+{byte_off + 8 * q}];")
                q, nw = q + 2, nw - 2
            else:
                w(f"ld.shared.v2.u32 {{%kq{k0}, %kq{k0 + 1}}}, [%sp+{byte_off + 8 * q}];")
                q, nw = q + 1, nw - 1

    def preload(byte_off, start=0):
        for q in range(start, npre, 2):
            w(f"ld.shared.v2.f64 {{%n{q}, %n{q + 1}}}, [%sp+{byte_off + HB + 8 * q}];")

    def addrs(C):
        for c in range(C):
            w(f"add.u32 %ad{c}, %sx, %kq{c};")

    def dispatch(nwords):
        w(f"add.u32 %sp, %sp, {8 * nwords};")
        if anchors:     # "uniform anchor" (gen_interp.py, tail()): a dispatch site that tests the opcode with a uniform
            # predicate makes ptxas keep the opcode web in uniform registers (BRXU, jump-table read hoisted)
            w("setp.ne.u32 %pself, %op, 0x7fffffff; @%pself brx.idx %op, TBL; bra.uni FINISH;")
        else:
            w("brx.idx %op, TBL;")

    w("ld.shared.u32 %op, [%bufcur];")
    load_hdr(0, len(phases[0].ctx))
    preload(0)
    w("brx.idx %op, TBL;")
    w(f"H{OP_END}:")
    w("bra.uni FINISH;")
    # ---- NEXT (the offsets of the next record are loaded for the widest class: harmless)
    w(f"H{OP_NEXT}:")
    w("mov.u32 %op, %nop;")
    w("cp.async.wait_all;")
    w("bar.sync 0;")
    w("mov.u32 %addr, %bufcur; mov.u32 %bufcur, %bufoth; mov.u32 %bufoth, %addr;")
    w(f"add.u64 %gs, %gs, {CH};")
    async_fetch("%bufoth")
    w("add.u32 %sp, %bufcur, 16;")
    load_hdr(0, cmax)
    preload(0)
    w("brx.idx %op, TBL;")
    # ---- ENTER_d / LEAVE_d per class: rare (a few dozen per tuple).  The g_k updates of a record go context by context,
    # load - fma - store in program order: correct whatever contexts share a node variable.
    for cl in range(ncls):
        ph = by_cls[cl]
        C = len(ph.ctx)
        arr = ctx_array(ph, parity)
        for d in (1, 2):
            w(f"H{OP_SKEL + 4 * cl + d - 1}:")
            w("mov.u32 %op, %nop;")
            addrs(C)
            load_hdr(HB, C)
            for c in range(C):
                w(f"ld.shared.f64 %xk{c}, [%ad{c}];")
            if d == 1:
                for c in range(C):
                    w(f"add.rn.f64 %vr, %vr, %v{c};")
                for c in range(C):
                    w(f"mov.f64 %a{c}, %xk{c}; mov.f64 %v{c}, %zero;")
            else:
                for c in range(C):
                    w(f"mov.f64 %vs{c}, %v{c}; mul.rn.f64 %a{c}, %a{c}, %xk{c}; mov.f64 %v{c}, %zero;")
            preload(HB)
            dispatch(H)
            w(f"H{OP_SKEL + 4 * cl + 2 + d - 1}:")
            w("mov.u32 %op, %nop;")
            addrs(C)
            nw = H if d == 1 else 2 * H                      # LEAVE_2: own header + the parent's node variables
            for c in range(C):
                w(f"ld.shared.f64 %xk{c}, [%ad{c}];")
            if d == 2:                                       # the parent's adjoints: a_c = x_{k1_c} (depth-1 node)
                for c in range(C):
                    w(f"ld.shared.u32 %addr, [%sp+{HB + 4 * c}]; add.u32 %addr, %addr, %sx; ld.shared.f64 %a{c}, [%addr];")
            load_hdr(8 * nw, C)
            for c in range(C):
                w(f"mov.f64 %u{c}, %v{c};")
                if d == 1:
                    w(f"mul.rn.f64 %v{c}, %xk{c}, %u{c}; mov.f64 %a{c}, %one;")
                else:
                    w(f"fma.rn.f64 %v{c}, %xk{c}, %u{c}, %vs{c};")
                g_at = f"[%ad{c}+{SGA[arr[c]]}]"
                w(f"ld.shared.f64 %gk, {g_at}; fma.rn.f64 %gk, %a{c}, %u{c}, %gk; st.shared.f64 {g_at}, %gk;")
            preload(8 * nw)
            dispatch(nw)
    # ---- TRANS_p: the register gradients of phase p are added to the (even) shared-memory accumulators of their member
    # variables, the registers are loaded for phase p + 1
    for p, ph in enumerate(phases):
        w(f"H{op_trans + p}:")
        w("mov.u32 %op, %nop;")
        nxt = phases[p + 1] if p + 1 < len(phases) else None
        load_hdr(HB, len(nxt.ctx) if nxt else 1)
        for j, o in zip(ph.sid, ph.orbs):
            for i in o:
                r = SGA[0] + kidx[ph.tau[i]] * NT * 8
                w(f"ld.shared.f64 %gk, [%sx+{r}]; add.rn.f64 %gk, %gk, %g{j}; st.shared.f64 [%sx+{r}], %gk;")
        for c in range(len(ph.ctx)):
            w(f"add.rn.f64 %vr, %vr, %v{c};")
        if nxt:
            load_phase(nxt)
        preload(HB)
        dispatch(H)

    # ---- LEAF_h
    def leaf_handler(label, ph, slots):
        C = len(ph.ctx)
        arr = ctx_array(ph, parity)
        n = 1 + len(slots)                                   # payload words: c0, coefs
        npad = n + (n & 1)
        nwords = rec_words(n, cmax)
        w(f"{label}:")
        w("mov.u32 %op, %nop;")
        addrs(C)
        load_hdr(8 * nwords, C)
        for c in range(C):
            w(f"ld.shared.f64 %xk{c}, [%ad{c}];")

        def coef(i):
            return f"%n{i}" if i < npre else f"%c{i - npre}"
        pairs = list(range(npre, npad, 2))                   # payload word pairs beyond the preloaded ones
        issued = 0

        def issue():
            nonlocal issued
            if issued < len(pairs):
                i = pairs[issued]
                w(f"ld.shared.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [%sp+{HB + 8 * i}];")
                issued += 1
        for _ in range(ahead):
            issue()
        for c in range(C):
            w(f"mul.rn.f64 %b{c}, %a{c}, %xk{c};")
        for c in range(C):                                   # early: independent of every store of this record
            w(f"ld.shared.f64 %gq{c}, [%ad{c}+{SGA[arr[c]]}];")
        done = 0
        for i in range(1, n):                                # coefficient word i multiplies slot slots[i - 1]
            j = slots[i - 1]
            if i >= npre and (i - npre) % 2 == 0:
                issue()
            for c in range(C):
                r = ph.pi[c][j]
                addend = f"%u{c}" if i > 1 else coef(0)
                w(f"fma.rn.f64 %u{c}, {coef(i)}, %x{r}, {addend};")
            for c in range(C):
                r = ph.pi[c][j]
                w(f"fma.rn.f64 %g{r}, {coef(i)}, %b{c}, %g{r};")
            while done + 2 <= min(i + 1, npre):              # a consumed pair of preload registers: the NEXT record's words
                w(f"ld.shared.v2.f64 {{%n{done}, %n{done + 1}}}, [%sp+{8 * nwords + HB + 8 * done}];")
                done += 2
        if n == 1:
            for c in range(C):
                w(f"mov.f64 %u{c}, {coef(0)};")
        preload(8 * nwords, done)
        # g_k: contexts of equal parity never share a node variable and the two parities have their own arrays, so the
        # loads above and these stores are independent of each other
        for c in range(C):
            w(f"fma.rn.f64 %v{c}, %xk{c}, %u{c}, %v{c};")
            w(f"fma.rn.f64 %gq{c}, %a{c}, %u{c}, %gq{c};")
        for c in range(C):
            w(f"st.shared.f64 [%ad{c}+{SGA[arr[c]]}], %gq{c};")
        dispatch(nwords)

    for hi, (cl, slots) in enumerate(handlers):
        leaf_handler(f"H{op_leaf0 + hi}", by_cls[cl], slots)
    # ---- epilogue: E and dE/dx = even + odd shared-memory accumulators (the last TRANS added the register gradients)
    w("FINISH:")
    w("@!%valid bra DONE;")
    w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %vr, %vr, %w;")
    w("st.global.f64 [%po], %vr; add.u64 %po, %po, %ldb;")
    assert kvars == list(range(nv)), "every variable is some context's node variable (true for pip3b)"
    for k in range(nv):
        r = kidx[k] * NT * 8
        w(f"ld.shared.f64 %gk, [%sx+{SGA[0] + r}]; ld.shared.f64 %w, [%sx+{SGA[1] + r}]; add.rn.f64 %gk, %gk, %w;")
        w(f"@%pscale ld.shared.f64 %w, [%sx+{r}]; @%pscale mul.rn.f64 %gk, %gk, %w;")
        w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %gk, %gk, %w;")
        w("st.global.f64 [%po], %gk; add.u64 %po, %po, %ldb;")
    w("DONE:")
    w("ret;")
    w("}")
    return out, SB + 2 * CH


def count_fp64(recs, handlers, phases):
    ncls, op_trans, op_leaf0, _ = op_layout(phases, handlers)
    n_fp, hist = 0, collections.Counter()
    p = 0
    for op, ks, payload, _ in recs:
        C = len(phases[min(p, len(phases) - 1)].ctx)
        if op >= op_leaf0:
            n_fp += C * (2 * len(handlers[op - op_leaf0][1]) + 3)
            hist["leaf"] += 1
        elif op >= op_trans:
            ph = phases[op - op_trans]
            n_fp += sum(len(o) for o in ph.orbs) + C
            if op - op_trans + 1 < len(phases):
                n_fp += sum(len(o) - 1 for o in phases[op - op_trans + 1].orbs)
            hist["trans"] += 1
            p += 1
        elif op >= OP_SKEL:
            kind = (op - OP_SKEL) % 4
            n_fp += C if kind < 2 else 2 * C
            hist["enter" if kind < 2 else "leave"] += 1
    return n_fp, hist


def generate(name, out_dir, block=256, piece=64, K=(12, 12), npre=8, mode="single", ahead=2, maxnreg=0, anchors=False,
             overhead=3):
    recs, handlers, phases, nv = serialise(name, K, mode, overhead)
    cmax = max(len(ph.ctx) for ph in phases)
    chunk = block * piece
    kvars = node_vars(recs)
    words = pack(recs, chunk, cmax, block, npre)
    n_fp, hist = count_fp64(recs, handlers, phases)
    symbol = f"{name}_stream0"
    lines = ["// GENERATED by tools/gen_msym.py -- molecule-symmetrised command-stream adjoint-Horner PIP evaluator.",
             ".version 8.7", ".target sm_100a", ".address_size 64", "", ".extern .shared .align 16 .b8 pip_smem[];", "",
             f".global .align 16 .b64 {symbol}[{len(words)}] = {{"]
    for i in range(0, len(words), 8):
        lines.append(", ".join("0x%016X" % int(v) for v in words[i:i + 8]) + ("," if i + 8 < len(words) else ""))
    lines.append("};")
    entry = f"{name}_eval"
    kl, smem = kernel_ptx(name, entry, symbol, handlers, nv, phases, kvars, block, piece, npre, ahead, maxnreg, anchors)
    lines += [""] + kl
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"{name}_eval.ptx"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    meta = {"name": name, "block": block, "smem_bytes": smem, "fp64_instr": n_fp, "chunk_bytes": chunk,
            "records": dict(hist), "kind": "msym", "contexts": [len(ph.ctx) for ph in phases], "node_vars": kvars,
            "kernels": [{"entry": entry, "vars": nv, "masks": len(handlers), "stream_bytes": int(words.nbytes), "fp64_instr": n_fp}],
            "handlers": [[c, len(s)] for c, s in handlers]}
    with open(os.path.join(out_dir, f"{name}_meta.json"), "w") as fh:
        json.dump(meta, fh)
    return meta, words, handlers, phases, kvars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GEN)
    ap.add_argument("--only", default="pip3b")
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--piece", type=int, default=64)
    ap.add_argument("--k0", type=int, default=12, help="leaf handlers of class 0 (no swap)")
    ap.add_argument("--k1", type=int, default=12, help="leaf handlers of class 1 (swap phase)")
    ap.add_argument("--npre", type=int, default=8)
    ap.add_argument("--ahead", type=int, default=2)
    ap.add_argument("--mode", default="single", help="single | swap | pairs (see symmetrise)")
    ap.add_argument("--overhead", type=int, default=3, help="pairs mode: cost of a record in slots (node split decision)")
    ap.add_argument("--maxnreg", type=int, default=0)
    ap.add_argument("--anchors", type=int, default=0)
    args = ap.parse_args()
    meta, *_ = generate(args.only, args.out, args.block, args.piece, (args.k0, args.k1), args.npre, args.mode,
                        args.ahead, args.maxnreg, bool(args.anchors), args.overhead)
    print(meta)


if __name__ == "__main__":
    main()
