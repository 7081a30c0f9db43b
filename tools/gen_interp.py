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


def serialise(name, tab, fuse=True):
    """-> (records, masks).  records: list of (opcode, k, payload doubles); masks: list of block-index tuples."""
    blocks = BLOCKS[name]
    coef = tab["coef"]
    monos = [(tuple(int(v) for v in row if v != PAD), int(c)) for row, c in zip(tab["mono_vars"], tab["mono_coef"])]
    root = build_tree(monos)
    freq = collections.Counter()

    def mask_of(node):
        return tuple(b for b, blk in enumerate(blocks) if any(v in node.leaves for v in blk))

    def count(node):
        if node.leaves:
            freq[mask_of(node)] += 1
        for k in node.kids:
            count(k)
    count(root)
    masks = [m for m, _ in freq.most_common()]
    mask_id = {m: i for i, m in enumerate(masks)}
    recs = []

    def emit(node, depth):
        assert depth <= MAX_DEPTH
        if depth > 0 and not node.kids and node.leaves and fuse:
            m = mask_of(node)
            payload = [coef[node.const] if node.const is not None else 0.0]
            payload += [coef[node.leaves[v]] if v in node.leaves else 0.0 for b in m for v in blocks[b]]
            recs.append((OP_MASK + len(masks) + mask_id[m], node.var, payload))
            return
        if depth > 0:
            recs.append((OP_ENTER + depth - 1, node.var, [coef[node.const] if node.const is not None else 0.0]))
        else:
            assert node.const is None
        if node.leaves:
            m = mask_of(node)
            payload = [coef[node.leaves[v]] if v in node.leaves else 0.0 for b in m for v in blocks[b]]
            recs.append((OP_MASK + mask_id[m], 0, payload))
        for k in node.kids:
            emit(k, depth + 1)
        if depth > 0:
            recs.append((OP_LEAVE + depth - 1, node.var, []))
    emit(root, 0)
    recs.append((OP_END, 0, []))
    return recs, masks


def rec_words(n_payload):
    return 1 + (n_payload if n_payload % 2 else n_payload + 1)


def pack(recs, chunk_bytes):
    """u64 words, organised in chunks of ``chunk_bytes`` that the kernel stages through shared memory one at a time.
    Inside a chunk, word 0 is padding so that every record header sits at byte offset 8 (mod 16) and every payload is
    16-byte aligned (payloads are padded to an odd number of words).  No record straddles a chunk: the last record of a
    chunk is OP_NEXT.  One all-zero chunk is appended so that the kernel's look-ahead load never leaves the array."""
    cw = chunk_bytes // 8
    chunks = []
    cur = [0]
    for op, k, payload in recs:
        need = rec_words(len(payload))
        if len(cur) + need + 8 > cw:              # room for OP_NEXT + the 6-word look-ahead the handlers preload
            cur += [OP_NEXT, 0]
            cur += [0] * (cw - len(cur))
            chunks.append(cur)
            cur = [0]
        cur.append(op | (k << 32))
        p = list(payload)
        if len(p) % 2 == 0:
            p.append(0.0)
        cur.extend(struct.unpack("<Q", struct.pack("<d", float(c)))[0] for c in p)
    cur += [0] * (cw - len(cur))
    chunks.append(cur)
    chunks.append([0] * cw)
    return np.array([w for c in chunks for w in c], dtype=np.uint64)


def interpret(name, words, x, chunk_bytes):
    """numpy/python reference interpreter of the packed stream (used by the CPU tests)."""
    cw = chunk_bytes // 8
    blocks = BLOCKS[name]
    recs_masks = None
    nv = sum(len(b) for b in blocks)
    x = np.asarray(x, dtype=np.float64)
    g = np.zeros(nv)
    f = words.view(np.float64)
    masks = interpret.masks[name]
    pc = 1
    v, a = 0.0, 1.0
    stack = []
    while True:
        hdr = int(words[pc])
        op, k = hdr & 0xFFFFFFFF, hdr >> 32
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
        elif op < OP_MASK + len(masks):
            m = masks[op - OP_MASK]
            q = pc + 1
            n = 0
            for b in m:
                for j in blocks[b]:
                    c = f[q]
                    q += 1
                    n += 1
                    v += c * x[j]
                    g[j] += c * a
            pc += 1 + (n if n % 2 else n + 1)
        else:                                   # fused leaf node: ENTER k, const; MASK_m; LEAVE k
            m = masks[op - OP_MASK - len(masks)]
            a2 = a * x[k]
            u = f[pc + 1]
            q = pc + 2
            n = 1
            for b in m:
                for j in blocks[b]:
                    c = f[q]
                    q += 1
                    n += 1
                    u += c * x[j]
                    g[j] += c * a2
            v += x[k] * u
            g[k] += a * u
            pc += 1 + (n if n % 2 else n + 1)


interpret.masks = {}


def hexd(v):
    return "0d%016X" % struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def to_ptx(name, words, masks, nv, block=256, min_blocks=1, nacc=2, ahead=4):
    """PTX module: the command stream as initialised .global data plus the interpreter kernel

        {name}_eval(const double* x /*[nv][ld]*/, double* out /*[1+nv][ld]: E, dE/dx*/, u32 n, u32 ld)

    One thread per tuple, x / g / adjoints in registers.  The stream is staged through a double-buffered shared-memory
    window: every thread fetches 16 bytes of chunk c+1 into registers when chunk c starts, stores them when chunk c
    ends (OP_NEXT) and one bar.sync flips the buffers -- the global-load latency is never exposed, the handlers read
    coefficients with broadcast ld.shared (every lane the same address)."""
    blocks = BLOCKS[name]
    NT = block
    CH = NT * 16                       # chunk bytes = one 16-byte piece per thread
    SX, SG, SB = 0, nv * NT * 8, 2 * nv * NT * 8       # smem: x copy | g accumulators | 2 stream buffers
    out = []
    w = out.append
    w("// GENERATED by tools/gen_interp.py -- command-stream adjoint-Horner PIP evaluator (see the tool's docstring).")
    w(".version 8.7")
    w(".target sm_100a")
    w(".address_size 64")
    w("")
    w(f".global .align 16 .b64 {name}_stream[{len(words)}] = {{")
    for i in range(0, len(words), 8):
        w(", ".join("0x%016X" % int(v) for v in words[i:i + 8]) + ("," if i + 8 < len(words) else ""))
    w("};")
    w(".extern .shared .align 16 .b8 pip_smem[];")
    w("")
    w(f".visible .entry {name}_eval(.param .u64 p_x, .param .u64 p_out, .param .u32 p_n, .param .u32 p_ld)")
    w(f".maxntid {NT}, 1, 1")
    w(f".minnctapersm {min_blocks}")
    w("{")
    w(f".reg .f64 %x<{nv}>, %g<{nv}>, %c<32>, %vc, %vd, %ac, %w, %cst, %gk, %one, %u0, %u1, %a0, %xk0, %pf0, %pf1, %n<6>, %u2, %u3;")
    w(".reg .f64 " + ", ".join(f"%vs{d}, %vt{d}, %as{d}, %xk{d}" for d in range(1, MAX_DEPTH + 1)) + ";")
    w(".reg .u64 %px, %po, %off, %ldb, %gs; .reg .u32 %n, %ld, %t, %tc, %r<4>, %op, %k, %nop, %nk, %sx, %sg, %addr;")
    w(".reg .u32 %sp, %bufcur, %bufoth, %mine;")
    w(".reg .pred %valid;")
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
    w("shl.b32 %mine, %r2, 4;")                                    # this thread's 16-byte piece of a chunk
    # stream chunk 0 -> buffer 0, chunk 1 -> registers
    w(f"mov.u64 %gs, {name}_stream; cvta.to.global.u64 %gs, %gs; cvt.u64.u32 %off, %mine; add.u64 %gs, %gs, %off;")
    w("ld.global.nc.v2.f64 {%pf0, %pf1}, [%gs];")
    w("add.u32 %addr, %bufcur, %mine; st.shared.v2.f64 [%addr], {%pf0, %pf1};")
    w(f"add.u64 %gs, %gs, {CH}; ld.global.nc.v2.f64 {{%pf0, %pf1}}, [%gs];")
    w("mov.f64 %one, 0d3FF0000000000000;")
    for i in range(nv):
        w(f"ld.global.nc.f64 %x{i}, [%px]; add.u64 %px, %px, %ldb;")
    for i in range(nv):
        w(f"st.shared.f64 [%sx+{i * NT * 8}], %x{i}; st.shared.f64 [%sg+{i * NT * 8}], 0d0000000000000000;")
        w(f"mov.f64 %g{i}, 0d0000000000000000;")
    w("mov.f64 %vc, 0d0000000000000000; mov.f64 %vd, 0d0000000000000000; mov.f64 %ac, %one;")
    w("bar.sync 0;")
    w("add.u32 %sp, %bufcur, 8;")
    n_ops = OP_MASK + 2 * len(masks)
    NPRE = 6                                   # payload words of the NEXT record pre-loaded by every handler
    w("TBL: .branchtargets " + ", ".join(f"H{i}" for i in range(n_ops)) + ";")

    def preload(byte_off):
        """first NPRE payload words of the record at %sp+byte_off -> %n0..%n5 (garbage past short records: unused)"""
        for q in range(0, NPRE, 2):
            w(f"ld.shared.v2.f64 {{%n{q}, %n{q + 1}}}, [%sp+{byte_off + 8 + 8 * q}];")

    def head(nwords):
        """current k <- header fetched by the previous handler; fetch the next header"""
        w("mov.u32 %k, %nk;")
        w(f"ld.shared.v2.u32 {{%nop, %nk}}, [%sp+{8 * nwords}];")

    def tail(nwords):
        preload(8 * nwords)
        w(f"add.u32 %sp, %sp, {8 * nwords};")
        w("brx.idx %nop, TBL;")                 # per-handler dispatch: ptxas can hoist the jump-table read

    def coef(i):
        return f"%n{i}" if i < NPRE else f"%c{i - NPRE}"

    w("ld.shared.v2.u32 {%nop, %nk}, [%sp];")
    preload(0)
    w("brx.idx %nop, TBL;")
    # ---- END
    w(f"H{OP_END}:")
    w("bra.uni FINISH;")
    # ---- NEXT: publish chunk c+1 (held in registers since chunk c began), flip buffers, start fetching chunk c+2
    w(f"H{OP_NEXT}:")
    w("add.u32 %addr, %bufoth, %mine; st.shared.v2.f64 [%addr], {%pf0, %pf1};")
    w("bar.sync 0;")
    w("mov.u32 %addr, %bufcur; mov.u32 %bufcur, %bufoth; mov.u32 %bufoth, %addr;")
    w(f"add.u64 %gs, %gs, {CH}; ld.global.nc.v2.f64 {{%pf0, %pf1}}, [%gs];")
    w("add.u32 %sp, %bufcur, 8;")
    w("ld.shared.v2.u32 {%nop, %nk}, [%sp];")
    preload(0)
    w("brx.idx %nop, TBL;")
    # ---- ENTER_d  (payload: const in %n0)
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_ENTER + d - 1}:")
        head(2)
        w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 %xk{d}, [%addr];")
        w(f"mov.f64 %vs{d}, %vc; mov.f64 %vt{d}, %vd; mov.f64 %as{d}, %ac;")
        w("mov.f64 %vc, %n0; mov.f64 %vd, 0d0000000000000000;")
        w(f"mul.rn.f64 %ac, %ac, %xk{d};")
        tail(2)
    # ---- LEAVE_d
    for d in range(1, MAX_DEPTH + 1):
        w(f"H{OP_LEAVE + d - 1}:")
        head(2)
        w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        w("add.rn.f64 %w, %vc, %vd;")
        w(f"mov.f64 %ac, %as{d};")
        w(f"fma.rn.f64 %vc, %xk{d}, %w, %vs{d}; mov.f64 %vd, %vt{d};")
        w("fma.rn.f64 %gk, %ac, %w, %gk; st.shared.f64 [%addr], %gk;")
        tail(2)

    # ---- MASK_m and fused LEAFNODE_m
    def block_handler(label, m, fused):
        """Hand-scheduled for in-order issue (assembled with ptxas -O1, which keeps this order): loads first, then the
        value FMAs on `nacc` independent chains, the adjoint product as late as its shared-memory operand allows, and the
        chain reduction + parent update interleaved with the (mutually independent) gradient FMAs."""
        w(f"{label}:")
        vars_ = [j for b in m for j in blocks[b]]
        slots = ([None] if fused else []) + vars_          # None = the node constant (first payload word)
        n = len(slots)
        nwords = rec_words(n)
        head(nwords)
        if fused:
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sx; ld.shared.f64 %xk0, [%addr];")
            w(f"mad.lo.u32 %addr, %k, {NT * 8}, %sg; ld.shared.f64 %gk, [%addr];")
        for i in range(NPRE, n, 2):                        # the rest of this record's payload
            if i + 1 < n:
                w(f"ld.shared.v2.f64 {{{coef(i)}, {coef(i + 1)}}}, [%sp+{8 + 8 * i}];")
            else:
                w(f"ld.shared.f64 {coef(i)}, [%sp+{8 + 8 * i}];")
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
            rops.append(f"fma.rn.f64 %vc, %xk0, {live[0]}, %vc;")
            rops.append(f"fma.rn.f64 %gk, %ac, {live[0]}, %gk;")
            rops.append("st.shared.f64 [%addr], %gk;")
        # emit: value ops, adjoint product after the first few, then gradient ops with the reduction spread between
        npre_v = min(len(vops), NPRE)
        for op in vops[:npre_v]:
            w(op)
        if fused:
            w("mul.rn.f64 %a0, %ac, %xk0;")
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

    for mi, m in enumerate(masks):
        block_handler(f"H{OP_MASK + mi}", m, False)
    for mi, m in enumerate(masks):
        block_handler(f"H{OP_MASK + len(masks) + mi}", m, True)
    # ---- epilogue
    w("FINISH:")
    w("@!%valid bra DONE;")
    w("add.rn.f64 %vc, %vc, %vd;")
    w("mul.wide.u32 %off, %t, 8; add.u64 %po, %po, %off;")
    w("st.global.f64 [%po], %vc; add.u64 %po, %po, %ldb;")
    for i in range(nv):
        w(f"ld.shared.f64 %gk, [%sg+{i * NT * 8}]; add.rn.f64 %g{i}, %g{i}, %gk;")
        w(f"st.global.f64 [%po], %g{i}; add.u64 %po, %po, %ldb;")
    w("DONE:")
    w("ret;")
    w("}")
    return out, SB + 2 * CH


def generate(name, out_dir, block=256, min_blocks=1, nacc=2, ahead=4, fuse=True):
    tab = dict(np.load(os.path.join(DATA, name + ".npz")))
    nv = int(tab["n_vars"])
    recs, masks = serialise(name, tab, fuse)
    chunk = block * 16
    words = pack(recs, chunk)
    interpret.masks[name] = masks
    blocks = BLOCKS[name]
    n_fp = 0
    hist = collections.Counter()
    for op, k, payload in recs:
        if op >= OP_MASK + len(masks):
            n_fp += 2 * sum(len(blocks[b]) for b in masks[op - OP_MASK - len(masks)]) + 4
            hist["leafnode"] += 1
        elif op >= OP_MASK:
            n_fp += 2 * sum(len(blocks[b]) for b in masks[op - OP_MASK])
            hist["mask"] += 1
        elif op >= OP_LEAVE:
            n_fp += 3
            hist["leave"] += 1
        elif op >= OP_ENTER:
            n_fp += 1
            hist["enter"] += 1
    ptx, smem = to_ptx(name, words, masks, nv, block, min_blocks, nacc, ahead)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"{name}_eval.ptx"), "w") as f:
        f.write("\n".join(ptx) + "\n")
    meta = {"name": name, "block": block, "smem_bytes": smem, "fp64_instr": n_fp, "chunk_bytes": chunk,
            "stream_bytes": int(words.nbytes), "records": dict(hist), "masks": len(masks), "kind": "interp"}
    with open(os.path.join(out_dir, f"{name}_meta.json"), "w") as f:
        json.dump(meta, f)
    return meta, words


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=GEN)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--min-blocks", type=int, default=1)
    ap.add_argument("--nacc", type=int, default=2)
    ap.add_argument("--ahead", type=int, default=4, help="16-byte coefficient pieces loaded ahead of their use")
    ap.add_argument("--no-fuse", action="store_true")
    args = ap.parse_args()
    for name in ("pip2b", "pip3b"):
        meta, _ = generate(name, args.out, args.block, args.min_blocks, args.nacc, args.ahead, not args.no_fuse)
        print(meta)


if __name__ == "__main__":
    main()
