#!/usr/bin/env python
"""Benchmark of the MB-pol 2B+3B many-body path (BASELINE.json metric: water-steps/s; % of the FP64 roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload synthetic_262144] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A step = one full force evaluation of the box: tuple-list build (every step, like the Reference platform), 2B and 3B
PIP energy + gradient over every listed dimer / trimer, switching functions, deterministic force reduction [+ position
broadcast and force/energy all-reduce for N > 1].  Workload: the synthetic liquid-water box of BASELINE.json
(configs[4], 262,144 waters = 8^3 tiles of the reference's 512-water cell) at every N -> strong scaling.

  value      water-steps/s with positions already resident in HBM (device pointers through the C ABI) and forces /
             energies left in HBM; K steps without instrumentation.  A second pass of the same K steps with per-stage
             CUDA events gives stage_ms and the roofline kernel's launch time.
  e2e        the same through the plugin-style API with HOST buffers: pinned host positions -> device, forces and
             energies -> host, every step
  roofline   the dominant kernel (pip3b_eval): algorithmic flops (111,818 per trimer) / its CUDA-event
             time inside the timed steps, against the FP64 FMA-pipe peak measured on this GPU by a register-resident
             DFMA loop (MEASURED_PEAKS.json carries no FP64 entry); nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2 TF
  cpu_baseline : the reference's own CPU code (oracle/_ref), ONE thread (= the Reference platform, README.md:35), on a
             bounded random sample of the same workload's tuples, extrapolated by the tuple counts; host CPU model and
             core count are stated.
  --impl reference : the same code fanned out over ALL host cores (tuple chunks on threads); a "step" is the bounded
             sample, `ms_per_step` its measured wall time, `value` the water-steps/s the sample rate extrapolates to.
  parity_vs_single (N > 1): before the timed region rank 0 also evaluates an UNPARTITIONED context on the same
             positions; the run exits non-zero if the N-GPU energies / forces differ by more than 1e-9 / 1e-6.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP2, FLOP3 = 30601, 111818
# dram__bytes_read.sum + dram__bytes_write.sum of one pip3b_eval launch inside this benchmark (7,103,084 trimers):
# profiles/r02_ncu_pip3b_eval_msym_details.txt (re-captured for the round-2 kernel: 2.054507 GB read + 2.059910 GB written)
NCU_DRAM_BYTES_PER_TRIMER = (2.054507e9 + 2.059910e9) / 7103084
NCU_DRAM_SOURCE = ("ncu --set full of the shipped kernel inside bench.py (profiles/r02_ncu_pip3b_eval_msym_details.txt): "
                   "579 B per trimer (algorithmic 584 B), scaled by the trimers of this launch")
NOMINAL_FP64_TF = 148 * 64 * 2 * 1.965e9 / 1e12


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def make_forces(n_mol, stride=3):
    from mbpol_openmm_plugin_b200 import MBPolThreeBodyForce, MBPolTwoBodyForce
    from mbpol_openmm_plugin_b200.systems import particle_indices
    idx = particle_indices(n_mol, stride).tolist()
    out = []
    for cls, cut in ((MBPolTwoBodyForce, 0.65), (MBPolThreeBodyForce, 0.45)):       # python/mbpol.xml:31-36
        f = cls()
        f.setCutoff(cut)
        f.setNonbondedMethod(cls.CutoffPeriodic)
        f._particles = idx                      # bulk equivalent of n_mol addParticle() calls
        out.append(f)
    return out


def host_cpu():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count() or 1, len(os.sched_getaffinity(0))


def cpu_reference_rate(pos_mol, box, n_dimers, n_trimers, seed=0, threads=1, lists=None):
    """Reference code (oracle/_ref if built, else the C port) on a random sample of the workload's own tuples, on
    `threads` host threads (1 = the Reference platform's own behaviour).
    Returns (water-steps/s extrapolated by tuple counts, description, kind, info dict)."""
    from oracle.refbind import best, eval_parallel
    o = best()
    posO = pos_mol[:, 0].copy()
    t0 = time.perf_counter()
    if lists is None:
        pairs, trips = o.pairs(posO, box, True, 0.65), o.triplets(posO, box, True, 0.45)
        t_lists = time.perf_counter() - t0
    else:
        pairs, trips, t_lists = lists
    rng = np.random.default_rng(seed)
    sp = pairs[rng.choice(len(pairs), min(n_dimers, len(pairs)), replace=False)]
    st = trips[rng.choice(len(trips), min(n_trimers, len(trips)), replace=False)]
    n = len(pos_mol)
    idx = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    pos = pos_mol.reshape(-1, 3)
    chunk = max(500, -(-max(len(sp), len(st)) // (4 * threads)))
    t0 = time.perf_counter()
    if threads == 1:
        o.eval_2b(pos, idx, sp, box, True)
    else:
        eval_parallel(o, 2, pos, idx, sp, box, True, threads, chunk)
    w2 = time.perf_counter() - t0
    t0 = time.perf_counter()
    if threads == 1:
        o.eval_3b(pos, idx, st, box, True)
    else:
        eval_parallel(o, 3, pos, idx, st, box, True, threads, chunk)
    w3 = time.perf_counter() - t0
    t2, t3 = w2 / len(sp), w3 / len(st)
    step_s = t_lists + len(pairs) * t2 + len(trips) * t3
    model, ncpu, avail = host_cpu()
    desc = (f"{len(sp)} dimers ({t2 * 1e6:.2f} us each) + {len(st)} trimers ({t3 * 1e6:.2f} us each) drawn at random from the "
            f"{len(pairs)} / {len(trips)} tuples of the workload on {threads} thread(s), full list build {t_lists:.2f} s "
            f"(one thread); extrapolated by tuple counts; {o.kind} code; host: {model}, {ncpu} cores ({avail} usable)")
    info = {"pairs": int(len(pairs)), "trimers": int(len(trips)), "sample_wall_s": w2 + w3, "lists_s": t_lists,
            "full_step_s": step_s, "host_cpu": model, "host_cores": ncpu}
    return n / step_s, desc, o.kind, info


def run_reference_arm(args, workload, pos_mol, box):
    """--impl reference: the reference's CPU implementation on all host cores, rank 0 only.  One step = the full list
    build (once, outside the steps, like a Verlet list kept for the sample) + a bounded random tuple sample; the
    line's ms_per_step is the MEASURED wall time of a sample step, its value the water-steps/s that rate extrapolates
    to for the whole box (tuple counts of the workload)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.refbind import best
    threads = len(os.sched_getaffinity(0))
    o = best()
    posO = pos_mol[:, 0].copy()
    t0 = time.perf_counter()
    lists = (o.pairs(posO, box, True, 0.65), o.triplets(posO, box, True, 0.45))
    lists = lists + (time.perf_counter() - t0,)
    rows = []
    desc = kind = info = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r, desc, kind, info = cpu_reference_rate(pos_mol, box, args.ref_dimers, args.ref_trimers, seed=i, threads=threads,
                                                 lists=lists)
        if i >= args.warmup:
            rows.append((r, time.perf_counter() - t0))
    value = float(np.mean([r for r, _ in rows]))
    n = len(pos_mol)
    line = {
        "impl": "reference", "metric": "2B+3B PIP force evals/s (water-steps/s)", "value": value, "unit": "water-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([w for _, w in rows])) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "waters": n, "dimers": info["pairs"], "trimers": info["trimers"],
                   "cutoff_2b_nm": 0.65, "cutoff_3b_nm": 0.45, "periodic": True, "lists": "rebuilt every step",
                   "parallelism": f"host threads x{threads} over tuple chunks",
                   "l2": "n/a (CPU)",
                   "step": f"bounded sample: {args.ref_dimers} dimers + {args.ref_trimers} trimers per step; value = "
                           "waters / (list build + tuples x measured time per tuple)"},
        "full_step_ms_extrapolated": n / value * 1e3,
        "cpu_baseline": {"value": value, "unit": "water-steps/s", "cores": threads, "kind": kind, "sample": desc,
                         "host_cpu": info["host_cpu"], "host_cores": info["host_cores"]},
        "e2e": {"value": value, "unit": "water-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_JSON_FD = None


def _reserve_stdout():
    """Rank 0's stdout must carry exactly ONE JSON line.  Libraries also write there (NCCL prints its version banner on
    stdout at first communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON
    line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthetic_262144")
    ap.add_argument("--ref-dimers", type=int, default=20000)
    ap.add_argument("--ref-trimers", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nve-steps", type=int, default=10, help="steps of the on-device NVE sweep line (0: skip; N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from mbpol_openmm_plugin_b200.systems import WORKLOADS, synthetic_box
    pos_mol, box = synthetic_box(WORKLOADS[args.workload])
    n = len(pos_mol)

    if args.impl == "reference":
        run_reference_arm(args, args.workload, pos_mol, box)
        return

    import torch
    import torch.distributed as dist
    from mbpol_openmm_plugin_b200 import MBPolContext, measure_fp64_peak
    from mbpol_openmm_plugin_b200.distributed import NcclForce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak_tf, _ = measure_fp64_peak()

    ctx = MBPolContext(3 * n, make_forces(n), box_nm=box.tolist(), device=local)
    ctx.engine.set_timing(True)
    # N > 1: broadcast -> owner-partitioned evaluation -> one grouped reduce, all inside the C library (csrc/dist.cu)
    nf = NcclForce(ctx) if world > 1 else None
    pos_dev = torch.as_tensor(pos_mol.reshape(-1, 3)).to(dev)
    f_dev = torch.zeros(3 * n, 3, dtype=torch.float64, device=dev)
    pos_host = torch.as_tensor(pos_mol.reshape(-1, 3)).pin_memory()
    f_host = torch.zeros(3 * n, 3, dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    keys = ("ms_pip_3b", "ms_pip_2b", "ms_lists", "ms_2b", "ms_3b", "ms_reduce", "ms_bcast", "ms_comm_reduce")
    acc = {k: 0.0 for k in keys}
    acc.update({"launches": 0, "n": 0, "t3": 0, "p2": 0, "repartitions": 0})

    def step_device():
        """one force evaluation with positions resident in HBM; energies come back to the host (16 bytes)"""
        if world == 1:
            f_dev.zero_()
            energy = ctx.engine.eval_device(pos_dev, f_dev, 3, stream=stream)
        else:
            energy, _ = nf.step(pos_dev if rank == 0 else None)
        st = ctx.engine.stats()
        for k in keys:
            acc[k] += st[k]
        acc["launches"] += st["n_launches"] + (2 if world > 1 else 1)     # + broadcast and grouped reduce / + memset
        acc["n"] += 1
        acc["t3"], acc["p2"], acc["repartitions"] = st["n_triplets_owned"], st["n_pairs_owned"], st["n_repartitions"]
        return energy

    def step_plain():
        """the same evaluation with NOTHING returned to the host inside the timed region (energies stay in HBM like the
        forces, as on the OpenMM CUDA path and in mbpol_nve_run) and without the per-stage events: the headline loop"""
        if world == 1:
            f_dev.zero_()
            ctx.engine.eval_device(pos_dev, f_dev, 3, want_energy=False, stream=stream)
        else:
            nf.step(pos_dev if rank == 0 else None, want_energy=False)

    def step_e2e():
        """the same through the public API with HOST buffers: pinned positions in, forces + energies out, every step"""
        if world == 1:
            ctx.setPositions(pos_host.numpy())
            ctx.setForceOutput(f_host.numpy())
            ctx.calcForcesAndEnergy()
        else:
            # pinned host -> device -> broadcast -> slices -> reduce in four parts, each copied back as soon as it is done
            nf.step_host(pos_host if rank == 0 else None, f_host if rank == 0 else None)

    # ---- parity of the N-GPU path against ONE unpartitioned context on the same positions (rank 0), before timing
    parity = None
    totals = None
    if world > 1:
        energy = step_device()
        if rank == 0:
            one = MBPolContext(3 * n, make_forces(n), box_nm=box.tolist(), device=local)
            f1 = torch.zeros_like(f_dev)
            e1 = one.engine.eval_device(pos_dev, f1, 3, stream=stream)
            st1 = one.engine.stats()
            totals = (int(st1["n_pairs"]), int(st1["n_triplets"]))
            parity = {"rel_dE2": abs(energy[0] - e1[0]) / abs(e1[0]), "rel_dE3": abs(energy[1] - e1[1]) / abs(e1[1]),
                      "max_abs_dF": float((nf.force - f1).abs().max().item()),
                      "bars": {"rel_dE": 1e-9, "max_abs_dF_kJ_mol_nm": 1e-6},
                      "against": "one unpartitioned context, same positions, same GPU as rank 0"}
            parity["ok"] = bool(parity["rel_dE2"] <= 1e-9 and parity["rel_dE3"] <= 1e-9 and parity["max_abs_dF"] <= 1e-6)
            one.engine.close()
        # ... and the host-buffer path (mbpol_eval_distributed_host: sliced reduce + overlapped copy-back) against it
        step_e2e()
        if rank == 0:
            parity["max_abs_dF_host_path"] = float((f_host.to(dev) - f1).abs().max().item())
            parity["ok"] = bool(parity["ok"] and parity["max_abs_dF_host_path"] <= 1e-6)
            del f1
        ok = torch.tensor([1 if (parity is None or parity["ok"]) else 0], device=dev)
        dist.broadcast(ok, src=0)
        if not int(ok.item()):
            if rank == 0:
                _emit({"error": "N-GPU result differs from the single-GPU result", "parity_vs_single": parity,
                       "n_gpus": world})
            dist.destroy_process_group()
            sys.exit(3)

    for _ in range(args.warmup):
        step_device()
    for k in acc:
        acc[k] = 0 if isinstance(acc[k], int) else 0.0
    rep0 = ctx.engine.stats()["n_repartitions"]
    # ---- instrumented pass: K steps with per-stage CUDA events and counters read back every step -> stage_ms, the roofline
    # kernel's launch time, launch counts
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_instr = timed(step_device, args.steps)
    stats = ctx.engine.stats()
    # ---- headline pass: the same K steps, device-resident results, no instrumentation (the stage events and the per-step
    # read-back of the energies cost a host synchronisation per step: 0.1 % of a 54 ms step, 3 % of an 8 ms step on 8 GPUs)
    ctx.engine.set_timing(False)
    for _ in range(2):
        step_plain()
    rep0 = ctx.engine.stats()["n_repartitions"]
    ms = timed(step_plain, args.steps)
    rep_in_window = ctx.engine.stats()["n_repartitions"] - rep0
    ctx.engine.set_timing(True)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- the rebalancing step (full list build on every rank) happens every MBPOL_REPARTITION_EVERY = 50 evaluations:
    # time single steps until one contains it, so that its cost is on record next to the steady-state number
    repart = None
    if world > 1:
        every = int(os.environ.get("MBPOL_REPARTITION_EVERY", "50"))
        slow = None
        for _ in range(every + 2 if every > 0 else 0):
            before = ctx.engine.stats()["n_repartitions"]
            t = timed(step_device, 1)
            if ctx.engine.stats()["n_repartitions"] != before:
                slow = t
                break
        repart = {"every_evaluations": every, "steps_with_rebalance_in_timed_window": int(rep_in_window),
                  "ms_of_a_step_with_rebalance": slow,
                  "amortised_ms_per_step": (ms / args.steps) + ((slow - ms_instr / args.steps) / every if slow and every else 0.0)}

    # ---- NVE sweep (BASELINE config 5), single GPU: the on-device velocity-Verlet loop of the C ABI (mbpol_nve_run), lists
    # rebuilt every step like the reference, then with a 0.05 nm Verlet skin; 298 K velocities, dt 0.2 fs.  Positions,
    # velocities, forces and energies never leave HBM inside the loop.
    nve_line = None
    if world == 1 and args.nve_steps > 0:
        from mbpol_openmm_plugin_b200 import nve
        ctx.engine.set_timing(False)
        m = torch.as_tensor(nve.masses(n).reshape(-1)).to(dev)
        nve_line = {"steps": args.nve_steps, "dt_fs": 0.2, "temperature_K": 298.0}
        for tag, skin in (("rebuild_every_step", 0.0), ("skin_0.05nm", 0.05)):
            ctx.engine.set_skin(skin)
            p = pos_dev.clone()
            v = torch.as_tensor(nve.maxwell_boltzmann(n, 298.0)).to(dev)
            nve.run_device(ctx, p, v, m, 2)                                   # warm-up (and the first skin build)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            rb0 = ctx.engine.stats()["n_row_builds"]
            e0.record()
            log = nve.run_device(ctx, p, v, m, args.nve_steps)
            e1.record()
            torch.cuda.synchronize()
            evals = args.nve_steps + 1                     # the run evaluates the forces of the start once more
            st = ctx.engine.stats()
            nve_line[tag] = {"ms_per_step": e0.elapsed_time(e1) / evals,
                             "water_steps_per_s": n * evals / (e0.elapsed_time(e1) * 1e-3),
                             "row_builds": int(st["n_row_builds"] - rb0), "force_evaluations": evals,
                             "energy_drift_kJ_mol": float(log[-1].sum() - log[0].sum())}
        ctx.engine.set_skin(0.0)
        ctx.engine.set_timing(True)

    if rank == 0:
        n_pairs, n_trips = totals if totals else (int(stats["n_pairs"]), int(stats["n_triplets"]))
        value = n * args.steps / (ms * 1e-3)
        e2e = n * args.steps / (ms_e2e * 1e-3)
        k = max(acc["n"], 1)
        ms_pip3 = acc["ms_pip_3b"] / k
        achieved = acc["t3"] * FLOP3 / (ms_pip3 * 1e-3) / 1e12 if ms_pip3 > 0 else 0.0
        ms_pip2 = acc["ms_pip_2b"] / k
        ach2 = acc["p2"] * FLOP2 / (ms_pip2 * 1e-3) / 1e12 if ms_pip2 > 0 else 0.0
        both = (acc["p2"] * FLOP2 + acc["t3"] * FLOP3) / ((ms_pip2 + ms_pip3) * 1e-3) / 1e12 if ms_pip2 + ms_pip3 > 0 else 0.0
        from mbpol_openmm_plugin_b200 import _lib
        ni3 = _lib.load().mbpol_pip_instruction_count(3)
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        step_ms = ms / args.steps
        stage = {"lists": acc["ms_lists"] / k, "2b": acc["ms_2b"] / k, "3b": acc["ms_3b"] / k,
                 "force_scatter": acc["ms_reduce"] / k, "pip2b_eval": ms_pip2, "pip3b_eval": ms_pip3,
                 "tuple_stage_2b_without_pip": acc["ms_2b"] / k - ms_pip2, "tuple_stage_3b_without_pip": acc["ms_3b"] / k - ms_pip3}
        if world > 1:
            stage.update({"broadcast": acc["ms_bcast"] / k, "reduce": acc["ms_comm_reduce"] / k})
        # the stages were timed in the instrumented pass: "rest" is what that pass spends outside them (collectives at
        # N > 1, launch gaps, the per-step host read-back); "step_instrumented" its whole step
        stage["step_instrumented"] = ms_instr / args.steps
        stage["rest"] = ms_instr / args.steps - sum(stage[x] for x in ("lists", "2b", "3b"))
        model, ncpu, _ = host_cpu()
        line = {
            "metric": "2B+3B PIP force evals/s (water-steps/s)", "value": value, "unit": "water-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "waters": n, "dimers": int(n_pairs), "trimers": int(n_trips),
                       "cutoff_2b_nm": 0.65, "cutoff_3b_nm": 0.45, "periodic": True, "lists": "rebuilt every step",
                       "parallelism": f"owner-molecule partition x{world}" +
                                      (", NCCL broadcast + grouped reduce inside libmbpol_b200.so" if world > 1 else ""),
                       "l2": "per-step working set (variables, gradients, force blocks: several GB) exceeds the 126 MB L2"},
            "alg_tflops": (n_pairs * FLOP2 + n_trips * FLOP3) * args.steps / (ms * 1e-3) / 1e12,
            "frac_of_nominal_fp64_peak": (n_pairs * FLOP2 + n_trips * FLOP3) * args.steps / (ms * 1e-3) / 1e12
                                         / (NOMINAL_FP64_TF * world),
            "e2e": {"value": e2e, "unit": "water-steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(pos_host.numel() * 8),
                    "d2h_bytes_per_step": int(f_host.numel() * 8 + 16)},
            "gpu_launches": int(acc["launches"]),
            "stage_ms": stage,
            "roofline": {"bound": "fp64", "kernel": "pip3b_eval", "achieved": achieved, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                         # dram__bytes_read+write of one `ncu --set full` capture of this kernel scaled to this launch
                         "traffic": acc["t3"] * NCU_DRAM_BYTES_PER_TRIMER, "traffic_unit": "bytes per launch",
                         "traffic_source": NCU_DRAM_SOURCE,
                         "peak_source": "measured on this GPU: register-resident DFMA loop (mbpol_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "nominal_peak": NOMINAL_FP64_TF, "frac_of_nominal": achieved / NOMINAL_FP64_TF,
                         "fp64_instr_per_trimer": ni3,
                         "fp64_pipe_issue_util": acc["t3"] / 32 * ni3 / (ms_pip3 * 1e-3) / (148 * 2 * sm_hz) if ms_pip3 > 0 else None,
                         "share_of_step": ms_pip3 / (ms_instr / args.steps), "pip2b_eval_tflops": ach2,
                         "timed_region": "the instrumented pass of the same K steps (stage_ms.step_instrumented); the "
                                         "headline ms_per_step comes from a pass without per-stage events",
                         "pip2b_frac": ach2 / peak_tf if peak_tf else None,
                         "pip2b_plus_pip3b_tflops": both, "pip2b_plus_pip3b_frac": both / peak_tf if peak_tf else None,
                         "algorithmic_flops_per_launch": acc["t3"] * FLOP3},
            "clocks": clocks,
            "host": {"cpu": model, "cores": ncpu},
        }
        if parity is not None:
            line["parity_vs_single"] = parity
        if repart is not None:
            line["repartition"] = repart
        if nve_line is not None:
            line["nve"] = nve_line
        if world == 1 and not args.no_cpu_baseline:
            v, desc, kind, info = cpu_reference_rate(pos_mol, box, args.ref_dimers, args.ref_trimers)
            line["cpu_baseline"] = {"value": v, "unit": "water-steps/s", "cores": 1, "kind": kind, "sample": desc,
                                    "host_cpu": info["host_cpu"], "host_cores": info["host_cores"]}
        _emit(line)
    if world > 1:
        dist.barrier()
        nf.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
