"""Micro-benchmark of the generated PIP evaluator cubins (development tool, GPU only).

    tools/mkvariant.sh TAG "ENV=.." [gen_interp.py arguments]        # -> build/mb/TAG/pip{2b,3b}.cubin
    python tools/pip_microbench.py build/mb/*/pip3b.cubin [--waves 8]   # tuples = 148 SMs x CTA size x waves

For every cubin: load it with the driver API, run {name}_eval on synthetic variables, check a few tuples against the
reference polynomial (oracle/_ref, if present) and report time, tuples/s, issued FP64 instructions/s and the
algorithmic TFLOP/s (30,601 flop per dimer, 111,818 per trimer).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys

import numpy as np
import torch
from cuda.bindings import driver as cu

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

ALG_FLOP = {"pip2b": 30601, "pip3b": 111818}
N_VARS = {"pip2b": 31, "pip3b": 36}


def ck(res):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def fp64_instr(cubin):
    log = os.path.join(os.path.dirname(cubin), os.path.basename(cubin)[:5] + "_eval.ptx")
    if not os.path.exists(log):
        return None
    txt = open(log).read()
    return len(re.findall(r"^(?:fma|mul|add)\.rn\.f64", txt, flags=re.M))


def ptx_block(cubin, default):
    ptx = os.path.join(os.path.dirname(cubin), os.path.basename(cubin)[:5] + "_eval.ptx")
    if os.path.exists(ptx):
        m = re.search(r"^\.maxntid (\d+)", open(ptx).read(), flags=re.M)
        if m:
            return int(m.group(1))
    return default


def load_meta(cubin):
    import json as _json
    p = os.path.join(os.path.dirname(cubin), os.path.basename(cubin)[:5] + "_meta.json")
    return _json.load(open(p)) if os.path.exists(p) else {}


def run(cubin, waves, block, reps, ref):
    name = os.path.basename(cubin)[:5]
    block = ptx_block(cubin, block)
    meta = load_meta(cubin)
    smem = int(meta.get("smem_bytes", 0))
    nv = N_VARS[name]
    data = np.fromfile(cubin, dtype=np.uint8)
    mod = ck(cu.cuModuleLoadData(data.ctypes.data))
    entries = [k["entry"] for k in meta.get("kernels", [])] or [f"{name}_eval"]
    fns = []
    for ent in entries:
        fn = ck(cu.cuModuleGetFunction(mod, ent.encode()))
        if smem:
            ck(cu.cuFuncSetAttribute(fn, cu.CUfunction_attribute.CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem))
        fns.append(fn)
    block = int(meta.get("block", block))
    n = 148 * block * waves                 # whole waves of one CTA per SM, whatever the CTA size
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(nv, n, dtype=torch.float64, device="cuda", generator=g) * 1.2 + 0.1
    out = torch.zeros(nv + 1, n, dtype=torch.float64, device="cuda")
    argps = []
    keep = []
    for i in range(len(fns)):
        args = (np.array([x.data_ptr()], dtype=np.uint64), np.array([out.data_ptr()], dtype=np.uint64),
                np.array([n], dtype=np.uint32), np.array([n], dtype=np.uint32), np.array([1 if i else 0], dtype=np.uint32))
        keep.append(args)
        argps.append(np.array([a.ctypes.data for a in args], dtype=np.uint64))
    stream = torch.cuda.current_stream().cuda_stream
    grid = (n + block - 1) // block

    def launch():
        for fn, argp in zip(fns, argps):
            ck(cu.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, smem, stream, argp.ctypes.data, 0))

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res = {"cubin": cubin, "n": n, "block": block, "ms": round(ms, 4), "tuples_per_s": n / ms * 1e3,
           "alg_tflops": n * ALG_FLOP[name] / ms / 1e9}
    ni = meta.get("fp64_instr") or fp64_instr(cubin)
    if ni:
        res["fp64_instr"] = ni
        # B200: 148 SM x 64 lanes -> 148*2 warp-instructions per clock
        res["fp64_pipe_util_at_1965MHz"] = n / 32 * ni / (ms * 1e-3) / (148 * 2 * 1.965e9)
    if ref is not None:
        xs = x[:, :6].T.contiguous().cpu().numpy()
        o = out[:, :6].T.contiguous().cpu().numpy()
        worst_e = worst_g = 0.0
        for i in range(6):
            e, gr = (ref.poly2b if name == "pip2b" else ref.poly3b)(xs[i])
            worst_e = max(worst_e, abs(o[i, 0] - e) / abs(e))
            worst_g = max(worst_g, np.max(np.abs(o[i, 1:] - gr)) / np.max(np.abs(gr)))
        res["rel_err_E"] = worst_e
        res["rel_err_g"] = worst_g
    cu.cuModuleUnload(mod)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cubins", nargs="+")
    ap.add_argument("--waves", type=int, default=8)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    ref = None
    try:
        from oracle.refbind import Ref
        ref = Ref()
    except Exception as e:  # noqa: BLE001
        print("no reference oracle:", e, file=sys.stderr)
    for c in args.cubins:
        try:
            print(json.dumps(run(c, args.waves, args.block, args.reps, ref)), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"cubin": c, "error": str(e)}), flush=True)


if __name__ == "__main__":
    main()
