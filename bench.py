#!/usr/bin/env python
"""bench.py -- NNGP log-likelihood evaluations per second (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3]

A step = one full log-likelihood evaluation (stages 2-3 fused: covariance build, factorisation,
reduction) over all n locations for one parameter vector, neighbours already built (cfg5 semantics,
SURVEY 8d1).  Workload at any N: BASELINE.json configs[2] -- synthetic 2-D, n = 1e6, m = 15,
Matern nu = 3/2, fp64 -- the configuration the metric is quoted on; with N > 1 the ordering is
sharded in contiguous blocks over the ranks (strong scaling: n is fixed) and the 3 partial statistics
are summed by an NCCL allreduce.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM (L2 flushed
between steps); `e2e` goes through the public API (pyNNGP.NNGP.loglik_terms: host parameters in,
host statistics out, every step); `roofline` is the fused kernel against the measured FP64 FMA
issue peak (and, secondary, the measured HBM peak); `cpu_baseline` is the C oracle on the host cores.
`--impl reference` times the reference's CPU path (the oracle port: the reference has no likelihood
implementation and no compiled sources) on a bounded sample with all host threads.
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

from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

KIDS = {"exponential": 0, "matern32": 1, "matern52": 2}
METRIC = "nngp_loglik_evals_per_sec"


def fp64_instr_per_location(m, D, kernel):
    """Algorithmic FP64-pipe thread-instructions per location (SURVEY 8d4 / BASELINE.md 4)."""
    p = m
    P = p * (p + 1) // 2
    k = {"exponential": 20, "matern32": 21, "matern52": 23}[kernel]
    cov = P * (2 * D + 8 + k)
    chol = (p - 1) * p * (p + 1) // 6 + p * (p - 1) // 2 + 16 * p
    solves = 2 * (p * (p - 1) // 2 + p)
    return cov + chol + solves + 2 * p + 41


def hbm_bytes_per_location(m, D):
    return 4 * m + 8 * D + 8


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, cmax = float(f[0]), float(f[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
            mx.append(cmax)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active") and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:]] if self.rows else []
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_sample_rate(s, y, m, kid, threads, sample, min_seconds):
    """Times the CPU oracle (oracle/nngp_oracle.c) on `sample` contiguous rows from the middle of the
    ordering; returns (seconds per full-n evaluation, description)."""
    from oracle import nngp_oracle as orc  # the checker, used here only as the timed CPU baseline

    n = len(s)
    lo = n // 2
    hi = min(n, lo + sample)
    nbr = orc.c_knn_ordered(s, m, lo=lo, hi=hi, threads=threads)  # untimed setup of the sample's rows
    reps, spent = 0, 0.0
    prm = (PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    while spent < min_seconds:
        t0 = time.perf_counter()
        orc.c_loglik(s, y, nbr, kid, *prm, lo=lo, hi=hi, threads=threads)
        spent += time.perf_counter() - t0
        reps += 1
    per_loc = spent / (reps * (hi - lo))
    return per_loc * n, f"{hi - lo} contiguous rows [{lo},{hi}) x {reps} reps, scaled to n={n}"


def reference_stage1_timing(cfg, sizes=None):
    """The reference's stage 1 as written (nngp.py:49-62: a scikit-learn KD-tree rebuilt for every i) timed
    on this host at small n -- the only part of the path the reference implements, single-threaded by
    construction -- with the quadratic extrapolation to the workload's n (x3.4-4.0 per
    doubling).  A few seconds of CPU work."""
    from oracle import nngp_oracle as orc  # the checker's restatement of the reference's own calls

    if sizes is None:  # NNGP_BENCH_STAGE1_SIZES=200,400 shortens it (the CPU test of this arm's JSON contract)
        sizes = tuple(int(k) for k in os.environ.get("NNGP_BENCH_STAGE1_SIZES", "1000,2000,4000").split(","))
    s, _ = synthetic(max(sizes), cfg["D"], cfg["seed"])
    orc.sk_reference_stage1(s[:200], cfg["m"])  # untimed: imports scikit-learn and warms its first call
    secs = []
    for k in sizes:
        t0 = time.perf_counter()
        orc.sk_reference_stage1(s[:k], cfg["m"])
        secs.append(time.perf_counter() - t0)
    per_pair = secs[-1] / (sizes[-1] ** 2)
    return {"what": "pyNNGP/nngp.py:49-62 as written (KDTree(s[0:i]) per i, scikit-learn), 1 thread",
            "n": list(sizes), "seconds": secs, "extrapolated_seconds_at_workload_n": per_pair * float(cfg["n"]) ** 2,
            "extrapolation": "quadratic from the largest n timed"}


def run_reference(args, cfg):
    """--impl reference: rank 0 only; other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    s, y = synthetic(cfg["n"], cfg["D"], cfg["seed"])
    threads = os.cpu_count() or 1
    kid = KIDS[cfg["kernel"]]
    per_step_budget = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    if args.cpu_seconds is not None:
        per_step_budget = args.cpu_seconds
    for _ in range(args.warmup):
        oracle_sample_rate(s[:200000], y[:200000], cfg["m"], kid, threads, 2048, 0.2)
    times, sample = [], ""
    for _ in range(args.steps):
        sec_per_eval, sample = oracle_sample_rate(s, y, cfg["m"], kid, threads, args.cpu_sample, per_step_budget)
        times.append(sec_per_eval)
    sec = float(np.mean(times))
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample,
                         "stage1_reference": reference_stage1_timing(cfg)},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (pyNNGP/nngp.py:73-96) has no likelihood implementation and no compiled sources; "
                "this arm times the C oracle port of its stubbed path on all host threads",
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, gpus):
    return {"workload": f"synthetic {cfg['D']}-D uniform, n={cfg['n']}, m={cfg['m']}, {cfg['kernel']}, "
                        f"sigma2={PARAMS['sigma2']} phi={PARAMS['phi']} tau2={PARAMS['tau2']}, neighbours prebuilt",
            "n": cfg["n"], "m": cfg["m"], "D": cfg["D"], "kernel": cfg["kernel"], "seed": cfg["seed"],
            "sharding": f"contiguous ordering blocks x{gpus}", "l2": "flushed between timed steps (256 MiB write)"}


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from pynngp_b200 import NNGP, Matern, Exponential, _lib
    from pynngp_b200.dist import DevicePtrView

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's own messages (its version banner under NCCL_DEBUG=VERSION/INFO) go to stderr: stdout carries
        # the one JSON line and nothing else
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    s, y = synthetic(cfg["n"], cfg["D"], cfg["seed"])
    kid = KIDS[cfg["kernel"]]
    spec = {"exponential": Exponential(**PARAMS), "matern32": Matern(1.5, **PARAMS), "matern52": Matern(2.5, **PARAMS)}[cfg["kernel"]]

    # ---- setup (untimed, reported): upload + stage 1 through the public class -------------------
    t0 = time.perf_counter()
    model = NNGP(s, y, 0.0, "S=T", cfg["m"], spec, dtype=args.dtype, device=local)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    eng = model._engine
    knn_s = model._timings["knn_s"]

    prm_host = np.array([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]])
    d_prm = torch.tensor(prm_host, dtype=torch.float64, device="cuda")
    d_out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()  # a real stream: the C ABI treats NULL as 'the handle's own stream'
    torch.cuda.set_stream(stream)

    fused_exchange = world > 1 and model._peer_ok and not args.nccl_allreduce

    def step_device():
        if fused_exchange:  # kernel + sum over ranks through NVLink peer memory in one launch
            eng.loglik_device_allreduce(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
        else:
            eng.loglik_device(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
            if world > 1:
                dist.all_reduce(d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step_device()
    barrier()
    ref_stats = d_out.cpu().numpy().copy()

    # ---- device-timed: K steps, each bracketed by CUDA events on the launching stream -------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    tw0 = time.time()
    for a, b in evs:
        flush.zero_()
        a.record(stream)
        step_device()
        b.record(stream)
    barrier()
    tw1 = time.time()
    launches = eng.launch_count() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = torch.tensor([float(step_ms.sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None

    # warm-L2 variant (no flush), reported in config for context
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        step_device()
    b.record(stream)
    barrier()
    warm_ms = a.elapsed_time(b) / args.steps

    # kernel-only duration for the roofline (no allreduce, flushed L2, per-launch events)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for ka, kb in kev:
        flush.zero_()
        ka.record(stream)
        eng.loglik_device(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
        kb.record(stream)
    barrier()
    kern_ms = float(np.mean([ka.elapsed_time(kb) for ka, kb in kev]))

    # ---- cfg5: a sweep of K parameter vectors in one launch (pair distances shared by the vectors) ----
    from pynngp_b200.synthetic import sweep_params

    Ks = 64
    d_prmK = torch.tensor(sweep_params(Ks), dtype=torch.float64, device="cuda")
    d_outK = torch.zeros((Ks, 3), dtype=torch.float64, device="cuda")

    def sweep_device():
        if fused_exchange:
            eng.loglik_device_allreduce(kid, d_prmK.data_ptr(), Ks, d_outK.data_ptr(), stream.cuda_stream)
        else:
            eng.loglik_device(kid, d_prmK.data_ptr(), Ks, d_outK.data_ptr(), stream.cuda_stream)
            if world > 1:
                dist.all_reduce(d_outK)

    sweep_device()
    barrier()
    sw = []
    for _ in range(3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        sweep_device()
        b.record(stream)
        barrier()
        sw.append(a.elapsed_time(b))
    sweep_t = torch.tensor([float(np.mean(sw))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sweep_t, op=dist.ReduceOp.MAX)
    sweep_ms = float(sweep_t.item())

    # ---- end to end through the public API (host params in, host stats out, every step) ----------
    for _ in range(3):
        model.loglik_terms()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_terms = model.loglik_terms(PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    # ---- the same with the response re-uploaded every step (a sampler that updates the field between
    #      evaluations: nngp_set_y, n doubles host -> device, then the evaluation) -------------------------
    y_alt = [np.ascontiguousarray(y), np.ascontiguousarray(y[::-1])]
    for k in range(2):
        eng.set_y(y_alt[k]); model.loglik_terms()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        eng.set_y(y_alt[k & 1])
        model.loglik_terms(PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    if world > 1:
        dist.barrier()
    e2e_y_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_y_t, op=dist.ReduceOp.MAX)
    e2e_y_s = float(e2e_y_t.item())
    eng.set_y(y_alt[0])

    # ---- cold path through the public API: host arrays -> upload -> stage 1 -> one evaluation ----
    cold = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        mdl = NNGP(s, y, 0.0, "S=T", cfg["m"], spec, dtype=args.dtype, device=local)
        cold_terms = mdl.loglik_terms()
        cold.append(time.perf_counter() - t0)
        cold_knn = mdl._timings["knn_s"]
        del mdl
    cold_t = torch.tensor([min(cold)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cold_t, op=dist.ReduceOp.MAX)
    cold_s = float(cold_t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused kernel --------------------------------------------------------------
    nloc = model._shard[1] - model._shard[0]
    ipl = fp64_instr_per_location(cfg["m"], cfg["D"], cfg["kernel"])
    peak_instr = eng.measure_fma_peak(args.dtype, 4096)
    achieved_tflops = nloc * ipl * 2 / (kern_ms * 1e-3) / 1e12
    peak_tflops = peak_instr * 2 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = nloc * hbm_bytes_per_location(cfg["m"], cfg["D"]) / (kern_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path) and args.config == "cfg3" and world == 1 and args.dtype == "float64":
        traffic = json.load(open(tr_path)).get("fused_dram_bytes_per_launch")  # ncu capture of this very workload
    roofline = {
        "bound": "fp64" if args.dtype == "float64" else "fp32",
        "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
        "traffic": traffic,
        "kernel": "nngp_fused::fused_loglik_kernel", "kernel_ms": kern_ms,
        "algorithmic_instr_per_location": ipl, "locations_per_launch": nloc,
        "peak_source": "measured here: register-resident FMA chains, nngp_measure_fma_peak (no vector-pipe figure in MEASURED_PEAKS.json)",
        "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "bytes_per_location": hbm_bytes_per_location(cfg["m"], cfg["D"]),
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- CPU baseline (the oracle port) on this host ---------------------------------------------
    threads = os.cpu_count() or 1
    sec_cpu, sample = oracle_sample_rate(s, y, cfg["m"], kid, threads, args.cpu_sample, 10.0)
    cpu_baseline = {"value": 1.0 / sec_cpu, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample}
    try:
        cpu_baseline["stage1_reference"] = reference_stage1_timing(cfg)
        cpu_baseline["stage1_reference"]["ours_seconds_at_workload_n"] = knn_s
    except Exception as exc:  # scikit-learn missing on the host: the likelihood baseline above still stands
        cpu_baseline["stage1_reference"] = {"unavailable": repr(exc)}

    # parity spot check on the very numbers being timed (cheap: the sample's rows)
    ms_per_step = total_ms / args.steps
    cfgd = workload_config(cfg, world)
    cfgd["sweep_cfg5"] = {"K": Ks, "ms_per_launch": sweep_ms, "ms_per_eval": sweep_ms / Ks, "evals_per_s": 1e3 * Ks / sweep_ms,
                          "what": "K parameter vectors (synthetic.sweep_params) in ONE launch: distances built once per location"}
    cfgd["exchange"] = ("none (1 GPU)" if world == 1 else
                        "fused into the kernel: P2P stores over NVLink peer memory (CUDA IPC), no NCCL call" if fused_exchange
                        else "NCCL all_reduce of 3 doubles after the kernel")
    cfgd.update({"setup_s": setup_s, "knn_build_s": knn_s, "knn_algo": "grid" if eng.knn_used_grid() else "brute",
                 "cold_e2e": {"ms": cold_s * 1e3, "knn_ms": cold_knn * 1e3, "h2d_bytes": int(cfg["n"]) * 32, "d2h_bytes": 24,
                              "what": "pyNNGP.NNGP(t, y, eps, 'S=T', m, cov) from host arrays (upload + stage 1) + one "
                                      "loglik_terms(); best of 2", "stats": list(cold_terms)}, "warm_l2_ms_per_step": warm_ms, "stats": ref_stats[0].tolist(),
                 "e2e_stats": list(e2e_terms),
                 "e2e_y_upload": {"value": args.steps / e2e_y_s, "unit": "evals/s", "h2d_bytes_per_step": 8 * int(cfg["n"]) + 32,
                                  "d2h_bytes_per_step": 24, "what": "nngp_set_y (the whole response, pageable host memory) + "
                                  "loglik_terms() every step"}})
    line = {
        "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic", "config": cfgd,
        "clocks": clocks,
        "e2e": {"value": args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 24,
                "api": "pyNNGP.NNGP.loglik_terms(sigma2, phi, tau2)"},
        "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--cpu-seconds", type=float, default=None, help="--impl reference: seconds of CPU work per step (default: 120 s over all steps, at most 20 s each)")
    ap.add_argument("--nccl-allreduce", action="store_true", help="multi-GPU: sum the statistics with NCCL instead of the fused peer-memory exchange")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
