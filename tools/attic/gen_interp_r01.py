"""Code generator v2: monomial table -> *cache-resident* adjoint-Horner evaluator (command stream + small PTX kernel).

Measured on B200 (profiles/r01_*): the fully unrolled adjoint-Horner program of tools/gen_horner.py (27 k / 72 k FP64
instructions, 0.4 / 1.4 MB of SASS) is instruction-fetch bound -- 38 % of warp stall cycles are `no_instruction`, FP64
pipe 31 % (3B) / 49 % (2B); a loop body of <= ~1.2 k instructions runs the same instruction mix at 84 %.  So the program
is turned into DATA: the Horner tree is serialised depth-first into a command stream (global memory, read by every warp
at the same address -> one L1/L2-resident sector per request) and a ~1 k-instruction kernel interprets it with all
per-tuple state in registers:

    ENTER_d k, c0 : push (v, a);  a *= x_k;  v = c0                      (x_k: dynamic index -> shared-memory copy of x)
    MASK_m coefs  : for every variable j of the blocks in mask m:  v += c_j x_j ;  g_j += c_j a     (static registers)
    LEAVE_d k     : w = v; pop (v, a);  v += x_k w ;  g_k += a w         (g_k: dynamic -> shared-memory accumulator)

Leaves of a node are stored *densely* per symmetry block of variables (zero coefficients where a monomial is absent):
the leaves cluster by block, so the padding costs only 11 % extra FMAs while the MASK handlers need no indexing at all.
Only 32 (3B) / 19 (2B) distinct (block-mask) patterns occur.

FP64 instructions per tuple: 2 * dense slots + 4 per node = ~79 k (3B), ~30 k (2B).

Shipped configuration (mbpol_openmm_plugin_b200/build.py GEN_ARGS), all measured on B200:

  * symmetry phases (`serialise_sym`): hydrogen / lone-pair swaps permute the variables; a node whose leaf coefficients are
    invariant under a set of swaps is evaluated on the ORBIT SUMS of the variables, in a phase of the stream in which the
    registers hold those sums -> 64 k (3B) / 22 k (2B) FP64 instructions; leaf handlers are specialised on arbitrary
    register-slot sets chosen by an agglomerative merge (`merge_library`) instead of on symmetry blocks;
  * SIBLING_d records (LEAVE_d + ENTER_d in one dispatch), look-ahead opcodes (LOOKAHEAD: the dispatch stays in the
    uniform datapath, BRXU), a second copy of the leaf handlers that follow themselves (`fall`);
  * the stream is staged in 16 KB chunks with cp.async straight into the idle shared-memory buffer (`piece` > 16);
  * the 2B kernel runs 384 threads per CTA (168 registers -> 3 warps per scheduler).

Everything else in this file (class split, PAIR / STATIC handlers, constant-bank stream, x_k prefetch, separate tail,
...) is an experiment that was built, verified and timed, and lost (profiles/r01_microbench_*.log).
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import struct

import numpy as np

from gen_horner import DATA, GEN, PAD, split_greedy

# variable blocks = orbits of the variables under the permutation group, split by molecule role
BLOCKS = {
    "pip3b": [[0, 1, 2], [3, 4, 5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 16], [17, 18, 19, 20], [21, 22, 23, 24],
              [25, 26, 27, 28], [29, 30, 31, 32], [33, 34, 35]],
    "pip2b": [[0, 1], [2, 3, 4, 5], [6, 7, 8, 9], [10, 11, 12, 13], [14], [15, 16, 17, 18], [19, 20, 21, 22], [23, 24],
              [25, 26], [27, 28, 29, 30]],
}
OP_END, OP_ENTER, OP_LEAVE, OP_NEXT, OP_MASK = 0, 1, 4, 7, 8   # ENTER_d = 1..3, LEAVE_d = 4..6, MASK_m = 8 + m,
# LEAFNODE_m = 8 + n_masks + m  (ENTER + MASK_m + LEAVE fused for nodes without children: 83 % of all nodes)
MAX_DEPTH = 3
# Header layout.  LOOKAHEAD: the low word of record r's header holds the opcode of record r + 1 (word 0 of the stream
# holds the first opcode), so a handler knows its successor on entry and the jump-table read + address arithmetic of
# its closing brx.idx no longer wait for the shared-memory load of the next header (ncu: 8 % of the warp stall cycles
# were short-scoreboard stalls of exactly that chain).
LOOKAHEAD = bool(int(os.environ.get("PIP_LOOKAHEAD", "1")))


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


def static_combos(recs, nm, nstatic):
    """the `nstatic` most frequent (mask id, node variable) pairs among the fused leaf-node records"""
    freq = collections.Counter((op - OP_MASK - nm, k) for op, k, _ in recs if OP_MASK + nm <= op < OP_MASK + 2 * nm)
    return [c for c, _ in freq.most_common(nstatic)]


def serialise(name, tab, fuse=True, class_blocks=None, with_consts=True, npair=0, nstatic=0):
    """-> (records, masks).  records: list of (opcode, k, payload doubles); masks: list of block-index tuples.

    ``class_blocks`` restricts the program to the leaves whose variable lies in those blocks (the polynomial is linear in
    its leaf sums, so the classes' partial values and gradients simply add); ``with_consts`` says whether this class
    also carries the monomials that end at an inner node."""
    blocks = BLOCKS[name]
    use = list(range(len(blocks))) if class_blocks is None else list(class_blocks)
    cvars = set(v for b in use for v in blocks[b])
    coef = tab["coef"]
    monos = [(tuple(int(v) for v in row if v != PAD), int(c)) for row, c in zip(tab["mono_vars"], tab["mono_coef"])]
    root = build_tree(monos)
    freq = collections.Counter()
    keep = {}

    def has(node):
        r = keep.get(id(node))
        if r is None:
            r = any(v in cvars for v in node.leaves) or (with_consts and node.const is not None)
            for k in node.kids:
                r = has(k) or r
            keep[id(node)] = r
        return r

    def leaves_of(node):
        return {v: c for v, c in node.leaves.items() if v in cvars}

    def mask_of(node):
        return tuple(b for b in use if any(v in node.leaves for v in blocks[b]))

    def count(node):
        if leaves_of(node):
            freq[mask_of(node)] += 1
        for k in node.kids:
            if has(k):
                count(k)
    has(root)
    count(root)
    masks = [m for m, _ in freq.most_common()]
    mask_id = {m: i for i, m in enumerate(masks)}
    recs = []

    def const_of(node):
        return coef[node.const] if (with_consts and node.const is not None) else 0.0

    def emit(node, depth):
        assert depth <= MAX_DEPTH
        lv = leaves_of(node)
        kids = [k for k in node.kids if has(k)]
        # evaluation order of siblings is free: group the childless ones by mask so that equal records are adjacent
        def order(k):
            childless = not any(has(q) for q in k.kids) and bool(leaves_of(k))
            return (0, mask_id[mask_of(k)]) if childless else (1, 0)
        kids.sort(key=order)
        if depth > 0 and not kids and lv and fuse:
            m = mask_of(node)
            payload = [const_of(node)] + [coef[lv[v]] if v in lv else 0.0 for b in m for v in blocks[b]]
            recs.append((OP_MASK + len(masks) + mask_id[m], node.var, payload))
            return
        if depth > 0:
            recs.append((OP_ENTER + depth - 1, node.var, [const_of(node)]))
        else:
            assert node.const is None
        if lv:
            m = mask_of(node)
            payload = [coef[lv[v]] if v in lv else 0.0 for b in m for v in blocks[b]]
            recs.append((OP_MASK + mask_id[m], 0, payload))
        for k in kids:
            emit(k, depth + 1)
        if depth > 0:
            recs.append((OP_LEAVE + depth - 1, node.var, []))
    emit(root, 0)
    recs.append((OP_END, 0, []))
    if nstatic:
        # STATIC_c = LEAFNODE_m whose node variable k is fixed by the opcode: x_k and g_k are plain registers in the
        # handler, the shared-memory read of x_k and the read-modify-write of g_k disappear
        nm = len(masks)
        combos = static_combos(recs, nm, nstatic)
        cid = {c: i for i, c in enumerate(combos)}
        recs = [(OP_MASK + 3 * nm + cid[(op - OP_MASK - nm, k)], k, pl)
                if (OP_MASK + nm <= op < OP_MASK + 2 * nm and (op - OP_MASK - nm, k) in cid) else (op, k, pl)
                for op, k, pl in recs]
        serialise.combos[name] = combos
    if npair:
        # PAIR_m = two consecutive LEAFNODE_m records evaluated by one handler with interleaved, independent chains
        pairable = pair_masks(name, masks, npair)
        nm, merged, i = len(masks), [], 0
        while i < len(recs):
            op, k, pl = recs[i]
            m = op - OP_MASK - nm
            if m in pairable and i + 1 < len(recs) and recs[i + 1][0] == op:
                merged.append((OP_MASK + 2 * nm + pairable.index(m), k | (recs[i + 1][1] << 8),
                               list(pl) + list(recs[i + 1][2])))
                i += 2
            else:
                merged.append(recs[i])
                i += 1
        recs = merged
    return recs, masks


def pair_masks(name, masks, npair, max_slots=18):
    """mask ids (most frequent first) that get a PAIR handler: small enough that two payloads fit the register budget"""
    blocks = BLOCKS[name]
    ok = [mi for mi, m in enumerate(masks) if sum(len(blocks[b]) for b in m) <= max_slots]
    return ok[:npair]


serialise.combos = {}



# ---- symmetry phases ------------------------------------------------------------------------------------------------
# Each variable is a function of ONE site-site distance; a hydrogen (or lone-pair) swap inside one
# molecule permutes the variables and leaves the polynomial invariant.  A Horner node whose leaf coefficients are
# invariant under a set S of swaps is a linear form in the S-ORBIT SUMS of the variables:  sum_i c_i x_i =
# sum_orbits c_O (sum_{i in O} x_i), and its gradient contribution is the same number for every member of an orbit.
# The program is therefore run in PHASES, one per subgroup S that occurs: at the start of a phase the registers are
# loaded with the orbit sums (slot = lowest variable of the orbit), the leaf handlers work on those slots only, and at
# the end of the phase the register gradients are scattered to all orbit members.  ENTER / LEAVE keep using the
# original variables (shared-memory copies), so the tree skeleton is the same in every phase.
def _sites(name):
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


def serialise_sym(name, tab, K, skeleton_weight=40, sib=False):
    """-> (records, masks, phases).  masks: handler slot tuples; phases: list of orbit lists (phase 0 = identity).
    Opcodes: MASK_h = OP_MASK + h, LEAFNODE_h = OP_MASK + nm + h, TRANSITION_p = OP_MASK + 2 nm + p (scatter the
    register gradients of phase p, load the orbit sums of phase p + 1)."""
    import itertools
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
    bits = lambda need: sum(1 << j for j in need)
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
        return [coef[need[j]] if j in need else 0.0 for j in masks[hid[pick[bits(need)]]]]

    def emit(node, depth, p):
        assert depth <= MAX_DEPTH
        own = bool(node.leaves) and pidx[cls[id(node)][0]] == p
        kids = [k for k in node.kids if p in leafph[id(k)]]
        const = coef[node.const] if (node.const is not None and p == min(leafph[id(node)])) else 0.0
        h = hid[pick[bits(cls[id(node)][1])]] if own else -1

        def order(k):
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
    if sib:
        # SIBLING_d = LEAVE_d k_old immediately followed by ENTER_d k_new, c0: one record, one dispatch (k = k_new | k_old << 8)
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
    return recs, masks, phases


def rec_words(n_payload):
    return 1 + (n_payload if n_payload % 2 else n_payload + 1)


def pack(recs, chunk_bytes):
    """u64 words, organised in chunks of ``chunk_bytes`` that the kernel stages through shared memory one at a time.
    Inside a chunk, word 0 is padding so that every record header sits at byte offset 8 (mod 16) and every payload is
    16-byte aligned (payloads are padded to an odd number of words).  No record straddles a chunk: the last record of a
    chunk is OP_NEXT.  One all-zero chunk is appended so that the kernel's look-ahead load never leaves the array.
    With LOOKAHEAD the opcode stored in a header is the NEXT record's (see above)."""
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
    if LOOKAHEAD:
        chunks[0][0] = hdrs[0][2]
        for (c, i, _), (_, _, nxt) in zip(hdrs, hdrs[1:] + [(0, 0, OP_END)]):
            chunks[c][i] = (chunks[c][i] & ~0xFFFFFFFF) | nxt
    return np.array([w for c in chunks for w in c], dtype=np.uint64)


def interpret(name, words, x, chunk_bytes, masks=None, combos=(), phases=None):
    """numpy/python reference interpreter of the packed stream (used by the CPU tests).  With ``phases`` (symmetry
    mode) the handlers work on register slots xr / gr that hold orbit sums, ENTER / LEAVE on the original variables."""
    cw = chunk_bytes // 8
    nv = sum(len(b) for b in BLOCKS[name])
    blocks = [[j] for j in range(nv)] if phases else BLOCKS[name]
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros(nv)                       # the shared-memory accumulators (all of dE/dx at the end)
    xr, gr = (x.copy(), np.zeros(nv)) if phases else (x, g)
    f = words.view(np.float64)
    masks = interpret.masks[name] if masks is None else masks
    nm = len(masks)
    pc = 1
    v, a = 0.0, 1.0
    stack = []
    nxt = int(words[0]) & 0xFFFFFFFF
    while True:
        hdr = int(words[pc])
        op, k = hdr & 0xFFFFFFFF, hdr >> 32
        if LOOKAHEAD:
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
            q = pc + 1
            n = 0
            for b in m:
                for j in blocks[b]:
                    c = f[q]
                    q += 1
                    n += 1
                    v += c * xr[j]
                    gr[j] += c * a
            pc += 1 + (n if n % 2 else n + 1)
        elif phases and op >= OP_MASK + 2 * nm + len(phases):       # SIBLING_d: LEAVE k_old + ENTER k_new, c0
            w = v
            v, a, xk = stack.pop()
            v += xk * w
            g[k >> 8] += a * w
            stack.append((v, a, x[k & 0xFF]))
            a = a * x[k & 0xFF]
            v = f[pc + 1]
            pc += 2
        elif phases and op >= OP_MASK + 2 * nm:  # transition: scatter the gradients of phase p, load phase p + 1
            p = op - OP_MASK - 2 * nm
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
        else:                                   # fused leaf node(s): ENTER k, const; MASK_m; LEAVE k  (x2 for PAIR_m)
            pair = OP_MASK + 2 * nm <= op < OP_MASK + 3 * nm
            if op >= OP_MASK + 3 * nm:
                m = masks[combos[op - OP_MASK - 3 * nm][0]]
            elif pair:
                m = masks[pair_masks(name, masks, 1 << 30)[op - OP_MASK - 2 * nm]]
            else:
                m = masks[op - OP_MASK - nm]
            q = pc + 1
            n = 0
            for kk in ((k & 0xFF, k >> 8) if pair else (k,)):
                a2 = a * x[kk]
                u = f[q]
                q += 1
                n += 1
                for b in m:
                    for j in blocks[b]:
                        c = f[q]
                        q += 1
                        n += 1
                        u += c * xr[j]
                        gr[j] += c * a2
                v += x[kk] * u
                g[kk] += a * u
            pc += 1 + (n if n % 2 else n + 1)


interpret.masks = {}


def hexd(v):
    return "0d%016X" % struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def kernel_ptx(name, entry, stream, words, masks, nv, class_vars, block=256, min_blocks=1, nacc=2, ahead=4, piece=16,
               opaque=True, cmem=False, repeat=1, npair=0, combos=(), split_tail=False, phases=None, sib=False,
               fall=(), npre=6, fall_copies=2):
    """One interpreter kernel

        {entry}(const double* x /*[nv][ld]*/, double* out /*[1+nv][ld]: E, dE/dx*/, u32 n, u32 ld, u32 accumulate)

    One thread per tuple; x / g of the class variables and the adjoints in registers.  The stream is staged through a
    double-buffered shared-memory window: every thread fetches `piece` bytes of chunk c+1 into registers when chunk c
    starts, stores them when chunk c ends (OP_NEXT) and one bar.sync flips the buffers -- the global-load latency is
    never exposed, the handlers read coefficients with broadcast ld.shared (every lane the same address).
    Returns (lines, dynamic shared bytes)."""
    blocks = [[j] for j in range(nv)] if phases else BLOCKS[name]      # symmetry mode: a mask is a tuple of slots
    NT = block
    CH = NT * piece                    # chunk bytes = one piece per thread
    # EXPERIMENT (cmem=True): the whole stream sits in the constant bank (<= 64 KB) and is read with uniform ld.const;
    # no shared-memory staging, no OP_NEXT.  `repeat` re-runs the stream (timing experiments with truncated streams).
    LD, SP, SPV = ("ld.const", "%cp", "%cp") if cmem else ("ld.shared", "%sp", "%spv")
    cset = set(class_vars)
    pfld = "ld.global.nc.v2.f64 {%pf0, %pf1}" if piece == 16 else "ld.global.nc.f64 %pf0"
    pfst = "{%pf0, %pf1}" if piece == 16 else "%pf0"
    pfty = "v2.f64" if piece == 16 else "f64"
    # smem: x copy | g accumulators [| second g accumulator array (split_tail)] | 2 stream buffers
    SX, SG = 0, nv * NT * 8
    SG2 = 2 * nv * NT * 8
    SB = (3 if split_tail else 2) * nv * NT * 8
    out = []
    w = out.append
    w(f".visible .entry {entry}(.param .u64 p_x, .param .u64 p_out, .param .u32 p_n, .param .u32 p_ld, .param .u32 p_acc)")
    w(f".maxntid {NT}, 1, 1")
    w(f".minnctapersm {min_blocks}")
    w("{")
    w(f".reg .f64 %x<{nv}>, %g<{nv}>, %c<64>, %vc, %vd, %ac, %w, %cst, %gk, %one, %u0, %u1, %a0, %xk0, %pf0, %pf1, %n<16>, %u2, %u3;")
    w(".reg .f64 " + ", ".join(f"%vs{d}, %vt{d}, %as{d}, %xk{d}" for d in range(1, MAX_DEPTH + 1)) + ";")
    w(".reg .u64 %px, %po, %off, %ldb, %gs; .reg .u32 %n, %ld, %t, %tc, %r<4>, %op, %k, %nop, %nk, %sx, %sg, %addr;")
    w(".reg .u32 %sp, %spv, %z, %bufcur, %bufoth, %mine, %ka, %kb, %addrb; .reg .f64 %a1, %xkb, %gkb;")
    w(".reg .pred %valid, %pacc, %pself; .reg .u32 %acc; .reg .f64 %xkn; .reg .u32 %addrn;")
    w("ld.param.u32 %acc, [p_acc]; setp.ne.u32 %pacc, %acc, 0;")
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
    ASYNC = piece > 16            # chunks > 4 KB: cp.async straight into the idle buffer (no staging registers)
    assert piece % 16 == 0 or not ASYNC

    def async_fetch(buf):
        """this thread's 16-byte pieces of the chunk at %gs -> shared buffer `buf` (asynchronously)"""
        w(f"add.u32 %addr, {buf}, %mine;")
        for q in range(piece // 16):
            w(f"cp.async.ca.shared.global [%addr+{q * NT * 16}], [%gs+{q * NT * 16}], 16;")
    w(f"mul.lo.u32 %mine, %r2, {16 if ASYNC else piece};")           # this thread's piece of a chunk
    # stream chunk 0 -> buffer 0, chunk 1 -> registers (ASYNC: -> buffer 1, in flight)
    w(f"mov.u64 %gs, {stream}; cvta.to.global.u64 %gs, %gs; cvt.u64.u32 %off, %mine; add.u64 %gs, %gs, %off;")
    if ASYNC:
        async_fetch("%bufcur")
        w("cp.async.wait_all;")
        w(f"add.u64 %gs, %gs, {CH};")
        async_fetch("%bufoth")
    else:
        w(f"{pfld}, [%gs];")
        w(f"add.u32 %addr, %bufcur, %mine; st.shared.{pfty} [%addr], {pfst};")
        w(f"add.u64 %gs, %gs, {CH}; {pfld}, [%gs];")
    w("mov.f64 %one, 0d3FF0000000000000;")
    if opaque:
        # word 0 of the stream is padding (zero).  Reading one of its halves at a thread-dependent address gives a zero
        # ptxas cannot prove uniform: payload loads through %spv = %sp + 0 then stay in vector registers instead of
        # being moved to uniform registers one R2UR per 32 bits (580 R2UR vs 2135 DFMA in the static SASS otherwise).
        w(f"mov.u64 %off, {stream}; cvta.to.global.u64 %off, %off; and.b32 %r3, %r2, 1; mul.wide.u32 %ldb, %r3, 4;")
        w("add.u64 %off, %off, %ldb; ld.global.nc.u32 %z, [%off]; mul.wide.u32 %ldb, %ld, 8;")
    else:
        w("mov.u32 %z, 0;")
    for i in range(nv):
        w(f"ld.global.nc.f64 %x{i}, [%px]; add.u64 %px, %px, %ldb;")
    for i in range(nv):
        w(f"st.shared.f64 [%sx+{i * NT * 8}], %x{i}; st.shared.f64 [%sg+{i * NT * 8}], 0d0000000000000000;")
        if split_tail:
            w(f"st.shared.f64 [%sg+{i * NT * 8 + SG2 - SG}], 0d0000000000000000;")
        if i in cset:
            w(f"mov.f64 %g{i}, 0d0000000000000000;")
    w("mov.f64 %vc, 0d0000000000000000; mov.f64 %vd, 0d0000000000000000; mov.f64 %ac, %one;")
    w("bar.sync 0;")
    w("add.u32 %sp, %bufcur, 8; add.u32 %spv, %sp, %z;")
    n_pairs_ops = len(pair_masks(name, masks, npair)) if npair else 0
    n_ops = OP_MASK + 3 * len(masks) + len(combos) if combos else OP_MASK + 2 * len(masks) + n_pairs_ops
    if phases:
        assert not combos and not npair
        n_ops = OP_MASK + 2 * len(masks) + len(phases) + (MAX_DEPTH if sib else 0)
    NPRE = int(os.environ.get("PIP_NPRE", npre))  # payload words of the NEXT record pre-loaded by every handler
    targets = [f"H{i}" for i in range(n_ops)]
    if combos:                                   # unused pair slots of the opcode table fall through to END
        for i in range(OP_MASK + 2 * len(masks) + n_pairs_ops, OP_MASK + 3 * len(masks)):
            targets[i] = f"H{OP_END}"
    w("TBL: .branchtargets " + ", ".join(targets) + ";")

    XKPRE = bool(int(os.environ.get("PIP_XKPRE", "0")))   # every handler loads x_k of the NEXT record before dispatching

    def next_xk():
        if XKPRE:
            w(f"and.b32 %addrn, %nk, 255; mad.lo.u32 %addrn, %addrn, {NT * 8}, %sx; ld.shared.f64 %xkn, [%addrn];")

    def preload(byte_off):
        """first NPRE payload words of the record at %sp+byte_off -> %n0..%n5 (garbage past short records: unused)"""
        for q in range(0, NPRE, 2):
            w(f"{LD}.v2.f64 {{%n{q}, %n{q + 1}}}, [{SPV}+{byte_off + 8 + 8 * q}];")

    DISP = "%op" if LOOKAHEAD else "%nop"

    def head(nwords):
        """current k <- header fetched by the previous handler; fetch the next header"""
        w("mov.u32 %k, %nk;")
        if LOOKAHEAD:
            w("mov.u32 %op, %nop;")
        if XKPRE:
            w("mov.f64 %xk0, %xkn;")
        w(f"{LD}.v2.u32 {{%nop, %nk}}, [{SP}+{8 * nwords}];")
        next_xk()

    FALL_COPIES = int(os.environ.get("PIP_FALL_COPIES", fall_copies))
    SELFBR = bool(int(os.environ.get("PIP_SELF_BRANCH", "0")))
    cur_label = [None, None]
    fall_state = [None]

    EARLY = bool(int(os.environ.get("PIP_EARLY_PRELOAD", "1")))
    # ptxas cannot prove that the per-thread st.shared of g_k does not alias the stream window, so a preload that follows
    # the store in program order can never be scheduled above it (and its latency ends up exposed, as R2UR /
    # short-scoreboard stalls, at the end of the short handlers): with EARLY every handler issues its preloads of the
    # next record BEFORE its store.

    def early_preload(nwords, start=0):
        """preload pairs start.. of the next record now; returns the `preloaded` argument for tail()"""
        if not EARLY:
            return start
        for q in range(start, NPRE, 2):
            w(f"{LD}.v2.f64 {{%n{q}, %n{q + 1}}}, [{SPV}+{8 * nwords + 8 + 8 * q}];")
        return NPRE

    def tail(nwords, preloaded=0):
        for q in range(preloaded, NPRE, 2):
            w(f"{LD}.v2.f64 {{%n{q}, %n{q + 1}}}, [{SPV}+{8 * nwords + 8 + 8 * q}];")
        if cmem:
            w(f"add.u64 %cp, %cp, {8 * nwords};")
        else:
            w(f"add.u32 %sp, %sp, {8 * nwords}; add.u32 %spv, %spv, {8 * nwords};")
        if SELFBR and cur_label[0] is not None:
            w(f"setp.eq.u32 %pself, {DISP}, {cur_label[1]}; @%pself bra.uni {cur_label[0]};")
        if fall_state[0] is not None:
            # a second copy of this handler follows: when the next record has the same opcode no branch is taken
            w(f"setp.ne.u32 %pself, {DISP}, {fall_state[0]}; @%pself brx.idx {DISP}, TBL;")
            return
        w(f"brx.idx {DISP}, TBL;")              # per-handler dispatch: ptxas can hoist the jump-table read

    SEPTAIL = bool(int(os.environ.get("PIP_SEPTAIL", "0")))     # experiment: see block_handler
    TRICKLE = int(os.environ.get("PIP_TRICKLE", "4"))     # preload word pair q of the next record right after its last use
    FAKE = bool(int(os.environ.get("PIP_FAKE_COEF", "0")))      # timing experiment only: skip the payload loads
    ALIAS = int(os.environ.get("PIP_FAKE_ALIAS", "0"))          # timing experiment only: fold x/g onto N registers

    def coef(i):
        if FAKE:
            return f"%n{i % NPRE}"
        return f"%n{i}" if i < NPRE else f"%c{i - NPRE}"

    if cmem:
        w(f".reg .u64 %cp; .reg .u32 %rep; .reg .pred %prep; mov.u32 %rep, {repeat};")
        w("RESTART:")
        w(f"mov.u64 %cp, {stream}; add.u64 %cp, %cp, 8;")
    if LOOKAHEAD:
        assert not opaque and not cmem
        w("ld.shared.u32 %op, [%bufcur];")         # word 0 of the stream: the first opcode
    w(f"{LD}.v2.u32 {{%nop, %nk}}, [{SP}];")
    next_xk()
    preload(0)
    w(f"brx.idx {DISP}, TBL;")
    # ---- END
    w(f"H{OP_END}:")
    if cmem:
        w("sub.u32 %rep, %rep, 1; setp.ne.u32 %prep, %rep, 0; @%prep bra.uni RESTART;")
    w("bra.uni FINISH;")
    # ---- NEXT: publish chunk c+1 (held in registers since chunk c began), flip buffers, start fetching chunk c+2
    w(f"H{OP_NEXT}:")
    if LOOKAHEAD:
        w("mov.u32 %op, %nop;")
    if ASYNC:
        w("cp.async.wait_all;")
    else:
        w(f"add.u32 %addr, %bufoth, %mine; st.shared.{pfty} [%addr], {pfst};")
    w("bar.sync 0;")
    w("mov.u32 %addr, %bufcur; mov.u32 %bufcur, %bufoth; mov.u32 %bufoth, %addr;")
    if ASYNC:                                   # everyone has left the old buffer: refill it with chunk c + 2
        w(f"add.u64 %gs, %gs, {CH};")
        async_fetch("%bufoth")
    else:
        w(f"add.u64 %gs, %gs, {CH}; {pfld}, [%gs];")
    w("add.u32 %sp, %bufcur, 8; add.u32 %spv, %sp, %z;")
    w("ld.shared.v2.u32 {%nop, %nk}, [%sp];")
    next_xk()
    preload(0)
    w(f"brx.idx {DISP}, TBL;")
    # ---- ENTER_d  (payload: const in %n0)
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_ENTER + d - 1}:")
        head(2)
        if XKPRE:
            w(f"mov.f64 %xk{d}, %xk0;")
        else:
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 %xk{d}, [%addr];")
        w(f"mov.f64 %vs{d}, %vc; mov.f64 %vt{d}, %vd; mov.f64 %as{d}, %ac;")
        w("mov.f64 %vc, %n0; mov.f64 %vd, 0d0000000000000000;")
        pre = early_preload(2)
        w(f"mul.rn.f64 %ac, %ac, %xk{d};")
        tail(2, pre)
    # ---- LEAVE_d
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_LEAVE + d - 1}:")
        head(2)
        w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        pre = early_preload(2)
        w("add.rn.f64 %w, %vc, %vd;")
        w(f"mov.f64 %ac, %as{d};")
        w(f"fma.rn.f64 %vc, %xk{d}, %w, %vs{d}; mov.f64 %vd, %vt{d};")
        w("fma.rn.f64 %gk, %ac, %w, %gk; st.shared.f64 [%addr], %gk;")
        tail(2, pre)

    # ---- MASK_m and fused LEAFNODE_m
    def block_handler(label, m, fused, static_k=None):
        """Hand-scheduled for in-order issue (assembled with ptxas -O1, which keeps this order): loads first, then the
        value FMAs on `nacc` independent chains, the adjoint product as late as its shared-memory operand allows, and the
        chain reduction + parent update interleaved with the (mutually independent) gradient FMAs."""
        w(f"{label}:")
        cur_label[0], cur_label[1] = (label, int(label[1:])) if (fused and label[1:].isdigit()) else (None, None)
        vars_ = [j for b in m for j in blocks[b]]
        slots = ([None] if fused else []) + vars_          # None = the node constant (first payload word)
        n = len(slots)
        nwords = rec_words(n)
        head(nwords)
        xk, gk = ("%xk0", "%gk") if static_k is None else (f"%x{static_k}", f"%g{static_k}")
        if fused and static_k is None:
            if not XKPRE:
                w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addr];")
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
            if split_tail:
                w(f"ld.shared.f64 %gkb, [%addr+{SG2 - SG}];")
        for i in range(NPRE, n if not FAKE else 0, 2):     # the rest of this record's payload
            if i + 1 < n:
                w(f"{LD}.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [{SPV}+{8 + 8 * i}];")
            else:
                w(f"{LD}.f64 {coef(i)}, [{SPV}+{8 + 8 * i}];")
        assert n - NPRE <= 32
        if fused:
            acc, adj = [f"%u{q}" for q in range(max(1, nacc))], "%a0"
        else:
            acc, adj = (["%vc", "%vd"] + [f"%u{q}" for q in range(2, max(2, nacc))])[:max(1, nacc)], "%ac"
        started = set() if fused else {"%vc", "%vd"}
        vops, gops = [], []
        for i, j in enumerate(slots):
            c = coef(i)
            if j is None:
                vops.append(f"mov.f64 %u0, {c};")
                started.add("%u0")
                continue
            a = acc[i % len(acc)]
            if ALIAS:
                j = j % ALIAS
            if a in started:
                vops.append(f"fma.rn.f64 {a}, {c}, %x{j}, {a};")
            else:
                vops.append(f"mul.rn.f64 {a}, {c}, %x{j};")
                started.add(a)
            gops.append(f"fma.rn.f64 %g{j}, {c}, {adj}, %g{j};")
        # reduction of the value chains (+ parent update for fused nodes)
        rops = []
        if not fused:
            for extra, base in (("%u2", "%vc"), ("%u3", "%vd")):
                if extra in acc and extra in started:
                    rops.append(f"add.rn.f64 {base}, {base}, {extra};")
        elif split_tail and static_k is None and len([a for a in acc if a in started]) == 2:
            # no reduction: each value chain feeds its own parent accumulator and its own g_k accumulator array, so the
            # record ends with independent FMA -> STS pairs instead of a DADD -> DFMA -> STS chain
            rops += [f"fma.rn.f64 %vc, {xk}, %u0, %vc;", f"fma.rn.f64 %vd, {xk}, %u1, %vd;",
                     "fma.rn.f64 %gk, %ac, %u0, %gk;", "fma.rn.f64 %gkb, %ac, %u1, %gkb;",
                     "st.shared.f64 [%addr], %gk;", f"st.shared.f64 [%addr+{SG2 - SG}], %gkb;"]
        elif SEPTAIL and static_k is None and len([a for a in acc if a in started]) >= 2:
            # the last reduction level is not taken: the two partial sums go separately into the parent's two value
            # accumulators and (back to back) into g_k -- one FP64 instruction more, one dependent level less
            live = [a for a in acc if a in started]
            while len(live) > 2:
                nxt = []
                for q in range(0, len(live) - 1, 2):
                    rops.append(f"add.rn.f64 {live[q]}, {live[q]}, {live[q + 1]};")
                    nxt.append(live[q])
                if len(live) % 2:
                    nxt.append(live[-1])
                live = nxt
            rops += [f"fma.rn.f64 %vc, {xk}, {live[0]}, %vc;", f"fma.rn.f64 %vd, {xk}, {live[1]}, %vd;",
                     f"fma.rn.f64 %gk, %ac, {live[0]}, %gk;", f"fma.rn.f64 %gk, %ac, {live[1]}, %gk;",
                     "st.shared.f64 [%addr], %gk;"]
        else:
            live = [a for a in acc if a in started]
            while len(live) > 1:
                nxt = []
                for q in range(0, len(live) - 1, 2):
                    rops.append(f"add.rn.f64 {live[q]}, {live[q]}, {live[q + 1]};")
                    nxt.append(live[q])
                if len(live) % 2:
                    nxt.append(live[-1])
                live = nxt
            rops.append(f"fma.rn.f64 %vc, {xk}, {live[0]}, %vc;")
            rops.append(f"fma.rn.f64 {gk}, %ac, {live[0]}, {gk};")
            if static_k is None:
                rops.append("st.shared.f64 [%addr], %gk;")
        if TRICKLE and fused and static_k is None:
            # slot pairs in payload order; once both FMAs of the two slots of a 16-byte payload piece are issued the
            # piece's registers are dead and the same piece of the NEXT record is loaded into them
            w(f"mul.rn.f64 %a0, %ac, {xk};")
            done = 0
            vi = gi = 0
            for q in range(0, n, TRICKLE):
                cnt = min(TRICKLE, n - q)
                for u in range(cnt):
                    w(vops[q + u])
                for u in range(cnt):
                    if slots[q + u] is not None:
                        w(gops[gi])
                        gi += 1
                for p2 in range(q, q + cnt + (cnt & 1), 2):
                    if p2 + 1 < NPRE:
                        w(f"{LD}.v2.f64 {{%n{p2}, %n{p2 + 1}}}, [{SPV}+{8 * nwords + 8 + 8 * p2}];")
                        done = p2 + 2
            done = early_preload(nwords, done)
            for r in rops:
                w(r)
            tail(nwords, done)
            return
        # emit: value ops, adjoint product after the first few, then gradient ops with the reduction spread between
        npre_v = min(len(vops), NPRE)
        for op in vops[:npre_v]:
            w(op)
        if fused:
            w(f"mul.rn.f64 %a0, %ac, {xk};")
        for op in vops[npre_v:]:
            w(op)
        gap = max(1, len(gops) // (len(rops) + 1)) if rops else len(gops) + 1
        gi = 0
        for r in rops:
            for _ in range(gap):
                if gi < len(gops):
                    w(gops[gi])
                    gi += 1
            w(r)
        for op in gops[gi:]:
            w(op)
        tail(nwords)

    def pair_handler(label, m):
        """two LEAFNODE_m records at once: A accumulates on %u0/%u1 -> %vc, B on %u2/%u3 -> %vd"""
        w(f"{label}:")
        vars_ = [j for b in m for j in blocks[b]]
        n1 = 1 + len(vars_)
        n = 2 * n1
        nwords = rec_words(n)
        head(nwords)
        w("and.b32 %ka, %k, 255; shr.u32 %kb, %k, 8;")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addr];")
        w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sx; ld.shared.f64 %xkb, [%addrb];")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sg; ld.shared.f64 %gkb, [%addrb];")
        for i in range(NPRE, n, 2):
            if i + 1 < n:
                w(f"{LD}.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [{SPV}+{8 + 8 * i}];")
            else:
                w(f"{LD}.f64 {coef(i)}, [{SPV}+{8 + 8 * i}];")
        assert n - NPRE <= 48
        seq = {}
        for tag, base, accs, adj in (("A", 0, ["%u0", "%u1"], "%a0"), ("B", n1, ["%u2", "%u3"], "%a1")):
            vops, gops = [f"mov.f64 {accs[0]}, {coef(base)};"], []
            started = {accs[0]}
            for i, j in enumerate(vars_):
                c = coef(base + 1 + i)
                a = accs[(i + 1) % 2]
                if a in started:
                    vops.append(f"fma.rn.f64 {a}, {c}, %x{j}, {a};")
                else:
                    vops.append(f"mul.rn.f64 {a}, {c}, %x{j};")
                    started.add(a)
                gops.append(f"fma.rn.f64 %g{j}, {c}, {adj}, %g{j};")
            seq[tag] = (vops, gops, len(started) > 1)
        for a_op, b_op in zip(seq["A"][0], seq["B"][0]):
            w(a_op)
            w(b_op)
        w("mul.rn.f64 %a0, %ac, %xk0;")
        w("mul.rn.f64 %a1, %ac, %xkb;")
        if seq["A"][2]:
            w("add.rn.f64 %u0, %u0, %u1;")
            w("add.rn.f64 %u2, %u2, %u3;")
        rops = ["fma.rn.f64 %vc, %xk0, %u0, %vc;", "fma.rn.f64 %vd, %xkb, %u2, %vd;",
                "fma.rn.f64 %gk, %ac, %u0, %gk;", "fma.rn.f64 %gkb, %ac, %u2, %gkb;",
                "st.shared.f64 [%addr], %gk;", "st.shared.f64 [%addrb], %gkb;"]
        # all gradient FMAs of A (adjoint %a0 stays in the operand-reuse cache), then all of B: zipping them would
        # alternate the adjoint operand and put two dependent updates of the same g_j back to back
        gops = seq["A"][1] + seq["B"][1]
        gap = max(1, len(gops) // (len(rops) + 1))
        gi = 0
        for r in rops:
            for _ in range(gap):
                if gi < len(gops):
                    w(gops[gi])
                    gi += 1
            w(r)
        for op in gops[gi:]:
            w(op)
        tail(nwords)

    for mi, m in enumerate(masks):
        block_handler(f"H{OP_MASK + mi}", m, False)
    for mi, m in enumerate(masks):
        op = OP_MASK + len(masks) + mi
        if op in fall:
            # FALL_COPIES back-to-back copies: every copy but the last falls through into the next one when the next
            # record has the same opcode (runs of 2, 3 and 4 equal records are the rule: symmetry-related siblings)
            for c in range(FALL_COPIES):
                fall_state[0] = op if c + 1 < FALL_COPIES else None
                block_handler(f"H{op}" + ("" if c == 0 else "BCDEFGH"[c - 1]), m, True)
        else:
            block_handler(f"H{op}", m, True)
    for ci, (mi, kk) in enumerate(combos):
        assert kk in cset, "static node variable must be a register of this class"
        block_handler(f"H{OP_MASK + 3 * len(masks) + ci}", masks[mi], True, static_k=kk)
    pairable = pair_masks(name, masks, npair) if npair else []
    for pi, mi in enumerate(pairable):
        pair_handler(f"H{OP_MASK + 2 * len(masks) + pi}", masks[mi])
    # ---- TRANSITION_p (symmetry mode): scatter the register gradients of phase p to every member of their orbit
    # (shared-memory accumulators), then load the orbit sums of phase p + 1 and clear the register gradients
    for p, orbs in enumerate(phases or []):
        w(f"H{OP_MASK + 2 * len(masks) + p}:")
        cur_label[0] = cur_label[1] = None
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
    # ---- SIBLING_d (symmetry mode): LEAVE_d k_old + ENTER_d k_new, c0 in one record (k = k_new | k_old << 8)
    for d in range(1, MAX_DEPTH + 1) if (phases and sib) else []:
        w(f"H{OP_MASK + 2 * len(masks) + len(phases) + d - 1}:")
        head(2)
        w("shr.u32 %ka, %k, 8; and.b32 %kb, %k, 255;")
        w(f"mad.lo.u32 %addr, %ka, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        if XKPRE:
            w("mov.f64 %xkb, %xk0;")
        else:
            w(f"mad.lo.u32 %addrb, %kb, {NT * 8}, %sx; ld.shared.f64 %xkb, [%addrb];")
        w("add.rn.f64 %w, %vc, %vd;")
        w("mov.f64 %vc, %n0; mov.f64 %vd, 0d0000000000000000;")
        pre = early_preload(2)
        w(f"fma.rn.f64 %vs{d}, %xk{d}, %w, %vs{d};")          # the parent's value, still saved
        w(f"fma.rn.f64 %gk, %as{d}, %w, %gk; st.shared.f64 [%addr], %gk;")
        w(f"mov.f64 %xk{d}, %xkb; mul.rn.f64 %ac, %as{d}, %xkb;")
        tail(2, pre)
    # ---- epilogue
    w("FINISH:")
    w("@!%valid bra DONE;")
    w("add.rn.f64 %vc, %vc, %vd;")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off;")
    w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %vc, %vc, %w;")
    w("st.global.f64 [%po], %vc; add.u64 %po, %po, %ldb;")
    for i in range(nv):
        w(f"ld.shared.f64 %gk, [%sg+{i * NT * 8}];")
        if split_tail:
            w(f"ld.shared.f64 %gkb, [%sg+{i * NT * 8 + SG2 - SG}]; add.rn.f64 %gk, %gk, %gkb;")
        if phases:
            pass                                  # the last TRANSITION has already scattered the register gradients
        elif i in cset and (not int(os.environ.get("PIP_FAKE_ALIAS", "0")) or i < int(os.environ.get("PIP_FAKE_ALIAS", "0"))):
            w(f"add.rn.f64 %gk, %gk, %g{i};")
        w("@%pacc ld.global.f64 %w, [%po]; @%pacc add.rn.f64 %gk, %gk, %w;")
        w("st.global.f64 [%po], %gk; add.u64 %po, %po, %ldb;")
    w("DONE:")
    w("ret;")
    w("}")
    return out, SB + 2 * CH


CLASSES = {            # block ids per class kernel (None = one kernel with every block)
    "pip3b": {1: None, 2: [[1, 2, 3, 4], [0, 5, 6, 7, 8]], 3: [[2, 3, 4], [5, 6, 7], [0, 1, 8]]},
    "pip2b": {1: None, 2: [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]]},
}


def count_fp64(recs, masks, blocks, combos=(), phases=None):
    n_fp = 0
    hist = collections.Counter()
    for op, k, payload in recs:
        if phases and op >= OP_MASK + 2 * len(masks) + len(phases):
            n_fp += 4
            hist["sibling"] += 1
        elif phases and op >= OP_MASK + 2 * len(masks):
            p = op - OP_MASK - 2 * len(masks)
            n_fp += sum(len(o) for o in phases[p])
            if p + 1 < len(phases):
                n_fp += sum(len(o) - 1 for o in phases[p + 1])
            hist["transition"] += 1
        elif op >= OP_MASK + 3 * len(masks):
            n_fp += 2 * sum(len(blocks[b]) for b in masks[combos[op - OP_MASK - 3 * len(masks)][0]]) + 4
            hist["static"] += 1
        elif op >= OP_MASK + 2 * len(masks):
            pm = [mi for mi, m in enumerate(masks) if sum(len(blocks[b]) for b in m) <= 18]
            n_fp += 2 * (2 * sum(len(blocks[b]) for b in masks[pm[op - OP_MASK - 2 * len(masks)]]) + 4)
            hist["pair"] += 1
        elif op >= OP_MASK + len(masks):
            n_fp += 2 * sum(len(blocks[b]) for b in masks[op - OP_MASK - len(masks)]) + 4
            hist["leafnode"] += 1
        elif op >= OP_MASK:
            n_fp += 2 * sum(len(blocks[b]) for b in masks[op - OP_MASK])
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


def generate(name, out_dir, block=256, min_blocks=1, nacc=2, ahead=4, fuse=True, n_classes=1, piece=16, opaque=False,
             npair=0, nstatic=0, split_tail=False, sym=0, skeleton_weight=40, sib=False,
             fall=0, npre=6, fall_copies=2):
    """Write {name}_eval.ptx (one module, one kernel per variable class) and {name}_meta.json.
    Returns (meta, [(words, masks) per class])."""
    tab = dict(np.load(os.path.join(DATA, name + ".npz")))
    nv = int(tab["n_vars"])
    blocks = BLOCKS[name]
    classes = CLASSES[name][n_classes] or [list(range(len(blocks)))]
    chunk = block * piece
    lines = ["// GENERATED by tools/gen_interp.py -- command-stream adjoint-Horner PIP evaluator (see the tool's docstring).",
             ".version 8.7", ".target sm_100a", ".address_size 64", "", ".extern .shared .align 16 .b8 pip_smem[];", ""]
    streams, kernels = [], []
    n_fp = 0
    hist = collections.Counter()
    smem = 0
    phases, fall_ops = None, ()
    for ci, cb in enumerate(classes):
        if sym:                                  # symmetry phases: `sym` = size of the handler library
            assert len(classes) == 1 and not npair and not nstatic
            recs, masks, phases = serialise_sym(name, tab, sym, skeleton_weight, sib)
            combos = []
            runs = collections.Counter(a[0] for a, b in zip(recs, recs[1:]) if a[0] == b[0])
            lo, hi = OP_MASK + len(masks), OP_MASK + 2 * len(masks)
            fall_ops = tuple(sorted(op for op, c in runs.items() if fall and c >= fall and lo <= op < hi))
            words = pack(recs, chunk)
            f, h = count_fp64(recs, masks, [[j] for j in range(nv)], (), phases)
        else:
            recs, masks = serialise(name, tab, fuse, cb, with_consts=(ci == 0), npair=npair, nstatic=nstatic)
            combos = serialise.combos.get(name, []) if nstatic else []
            words = pack(recs, chunk)
            f, h = count_fp64(recs, masks, blocks, combos)
        n_fp += f
        hist.update(h)
        symbol = f"{name}_stream{ci}"
        lines.append(f".global .align 16 .b64 {symbol}[{len(words)}] = {{")
        for i in range(0, len(words), 8):
            lines.append(", ".join("0x%016X" % int(v) for v in words[i:i + 8]) + ("," if i + 8 < len(words) else ""))
        lines.append("};")
        cvars = sorted(v for b in cb for v in blocks[b])
        entry = f"{name}_eval" if len(classes) == 1 else f"{name}_eval_c{ci}"
        kl, sm = kernel_ptx(name, entry, symbol, words, masks, nv, cvars, block, min_blocks, nacc, ahead, piece, opaque,
                            npair=npair, combos=combos, split_tail=split_tail, phases=phases, sib=sib,
                            fall=fall_ops if sym else (), npre=npre, fall_copies=fall_copies)
        lines += [""] + kl
        smem = max(smem, sm)
        streams.append((words, masks, combos))
        kernels.append({"entry": entry, "vars": len(cvars), "masks": len(masks), "stream_bytes": int(words.nbytes),
                        "fp64_instr": f})
    interpret.masks[name] = streams[0][1]
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"{name}_eval.ptx"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    meta = {"name": name, "block": block, "smem_bytes": smem, "fp64_instr": n_fp, "chunk_bytes": chunk,
            "records": dict(hist), "kernels": kernels, "kind": "interp-sym" if sym else "interp",
            "phases": phases, "handlers": [len(m) for m in streams[0][1]] if sym else None, "sib": bool(sib),
            "lookahead": LOOKAHEAD}
    with open(os.path.join(out_dir, f"{name}_meta.json"), "w") as fh:
        json.dump(meta, fh)
    return meta, streams


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GEN)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--min-blocks", type=int, default=1)
    ap.add_argument("--nacc", type=int, default=2)
    ap.add_argument("--ahead", type=int, default=4, help="16-byte coefficient pieces loaded ahead of their use")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--classes3", type=int, default=1, help="variable classes (kernels) for the 3B polynomial")
    ap.add_argument("--classes2", type=int, default=1)
    ap.add_argument("--piece", type=int, default=16, choices=[8, 16, 32, 64, 128],
                    help="stream bytes each thread stages per chunk (> 16: cp.async into the idle buffer)")
    ap.add_argument("--npair", type=int, default=0, help="pair handlers for the N most frequent masks")
    ap.add_argument("--nstatic2", type=int, default=0, help="2B: leaf-node handlers specialised on the node variable")
    ap.add_argument("--nstatic3", type=int, default=0)
    ap.add_argument("--split-tail", action="store_true", help="second g_k accumulator array: no reduction chain at record end")
    ap.add_argument("--block2", type=int, default=0, help="2B: threads per CTA (0 = --block)")
    ap.add_argument("--piece2", type=int, default=0, help="2B: stream bytes per thread and chunk (0 = --piece)")
    ap.add_argument("--npre2", type=int, default=-1, help="2B: preload depth (-1 = --npre)")
    ap.add_argument("--sym2", type=int, default=0, help="2B: symmetry phases with a library of N leaf handlers (0 = off)")
    ap.add_argument("--sym3", type=int, default=0)
    ap.add_argument("--skeleton-weight", type=int, default=40)
    ap.add_argument("--fall-copies", type=int, default=2, help="copies of a duplicated leaf handler")
    ap.add_argument("--npre", type=int, default=6, help="payload words of the next record every handler pre-loads")
    ap.add_argument("--sib", action="store_true", help="symmetry mode: fuse LEAVE_d + ENTER_d into SIBLING_d records")
    ap.add_argument("--fall", type=int, default=0, help="symmetry mode: duplicate leaf handlers that follow themselves "
                    "at least N times (no taken branch between two equal records)")
    args = ap.parse_args()
    for name in ("pip2b", "pip3b"):
        two = name == "pip2b"
        meta, _ = generate(name, args.out, args.block2 if two and args.block2 else args.block, args.min_blocks, args.nacc,
                           args.ahead, not args.no_fuse, args.classes2 if two else args.classes3,
                           args.piece2 if two and args.piece2 else args.piece, npair=args.npair,
                           nstatic=args.nstatic2 if name == "pip2b" else args.nstatic3, split_tail=args.split_tail,
                           sym=args.sym2 if name == "pip2b" else args.sym3, skeleton_weight=args.skeleton_weight,
                           sib=args.sib, fall=args.fall, npre=args.npre2 if two and args.npre2 >= 0 else args.npre,
                           fall_copies=args.fall_copies)
        print(meta)


if __name__ == "__main__":
    main()
