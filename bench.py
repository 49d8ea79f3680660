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

`--gpus N` inside torchrun (WORLD_SIZE = N) is one process per GPU; `--gpus N` in a plain process runs the N GPUs
through ONE multi-device handle (nngp_create_multi), timed by events around each device's kernel.
Alongside cfg3 the line carries `detail.cfg4` (n = 1e7, m = 30, 3-D: the north_star's scaling target) and
`parity` (the timed statistics against the oracle over all n rows).
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
    """Times the CPU oracle (oracle/nngp_oracle.c) on `sample` rows strided over the upper three quarters of the
    ordering (the reference arm has no GPU-built table, and the oracle's own exact search costs O(i) per row, so
    the whole of n = 1e6 is out of reach on the host).  The sampled rows keep their true neighbour sets, spread
    over all preceding rows, so the gathers miss the caches as they do in a full pass: they are appended as rows
    n .. n+sample of a copy of the inputs and evaluated there.  Returns (seconds per full-n evaluation, description)."""
    from oracle import nngp_oracle as orc  # the checker, used here only as the timed CPU baseline

    n = len(s)
    sample = int(min(sample, max(1, n - n // 4)))
    rows = np.unique(np.linspace(n // 4, n - 1, sample).astype(np.int64))
    tab = orc.c_knn_rows(s, m, rows, threads=threads)  # untimed setup: the exact ordered search of the sampled rows
    s2 = np.concatenate([s, s[rows]])
    y2 = np.concatenate([y, y[rows]])
    reps, spent = 0, 0.0
    prm = (PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    while spent < min_seconds:
        t0 = time.perf_counter()
        orc.c_loglik_rows(s2, y2, tab, n, kid, *prm, threads=threads)
        spent += time.perf_counter() - t0
        reps += 1
    per_loc = spent / (reps * len(rows))
    return per_loc * n, (f"{len(rows)} rows strided over [{n // 4},{n}) with their true neighbour sets x {reps} reps, scaled to n={n} "
                         f"(extrapolated; {threads} threads)")


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
        oracle_sample_rate(s[:50000], y[:50000], cfg["m"], kid, threads, 512, 0.2)
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
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample, "extrapolated": True,
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


def _stat_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64)[:2], np.asarray(b, dtype=np.float64)[:2]
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from pynngp_b200 import NNGP, Matern, Exponential, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    single_process = world == 1 and args.gpus > 1  # ONE process, a multi-device handle (nngp_create_multi)
    if world != args.gpus and not single_process:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's own messages (its version banner under NCCL_DEBUG=VERSION/INFO) go to stderr: stdout carries
        # the one JSON line and nothing else
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)  # the banner is written to file descriptor 1 by the library itself: park it on stderr
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)  # the communicator is created here, outside every timed or reported region
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    ngpu = args.gpus
    devices = list(range(ngpu)) if single_process else local

    def barrier():
        if world > 1:
            dist.barrier()
        for d in (range(ngpu) if single_process else (local,)):
            torch.cuda.synchronize(d)

    def max_over_ranks(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in (range(ngpu) if single_process else (local,))]

    def flush_l2():
        for f in flushes:
            f.zero_()
        if single_process:  # the engine's own streams are not torch's: order the flush before the evaluation
            for d in range(ngpu):
                torch.cuda.synchronize(d)

    def measure(c, steps, want_extras):
        """Builds the model of configuration c through the public class and times `steps` evaluations."""
        s, y = synthetic(c["n"], c["D"], c["seed"])
        kid = KIDS[c["kernel"]]
        spec = {"exponential": Exponential(**PARAMS), "matern32": Matern(1.5, **PARAMS), "matern52": Matern(2.5, **PARAMS)}[c["kernel"]]
        barrier()
        t0 = time.perf_counter()
        model = NNGP(s, y, 0.0, "S=T", c["m"], spec, dtype=args.dtype, devices=devices)
        barrier()
        out = {"s": s, "y": y, "kid": kid, "spec": spec, "model": model, "setup_s": max_over_ranks(time.perf_counter() - t0),
               "knn_s": max_over_ranks(model._timings["knn_s"])}
        eng = model._engine
        prm_host = np.array([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]])
        fused_exchange = world > 1 and model._peer_ok and not args.nccl_allreduce
        out["fused_exchange"] = fused_exchange
        if single_process:
            eng.set_timing(True)

            def step():
                return eng.loglik(kid, prm_host)[0]

            for _ in range(max(args.warmup, 3)):
                flush_l2()
                ref = step()
            times = []
            launches0 = eng.launch_count()
            tw0 = time.time()
            for _ in range(steps):
                flush_l2()
                step()
                times.append(eng.last_eval_ms())  # events around each device's kernel, max over the devices
            tw1 = time.time()
            out.update(stats=np.asarray(ref), step_ms=np.array(times), launches=eng.launch_count() - launches0, wall=(tw0, tw1),
                       kern_ms=None, warm_ms=None)
            return out
        d_prm = torch.tensor(prm_host, dtype=torch.float64, device="cuda")
        d_out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
        stream = torch.cuda.Stream()  # a real stream: the C ABI treats NULL as 'the handle's own stream'
        torch.cuda.set_stream(stream)

        def step_device(nccl=False):
            if fused_exchange and not nccl:  # kernel + sum over ranks through NVLink peer memory in one launch
                eng.loglik_device_allreduce(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
            else:
                eng.loglik_device(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
                if world > 1:
                    dist.all_reduce(d_out)

        for _ in range(max(args.warmup, 3)):
            flush_l2()
            step_device()
        barrier()
        out["stats"] = d_out.cpu().numpy()[0].copy()
        if world > 1:  # the same numbers through the other exchange (summation order differs)
            step_device(nccl=fused_exchange)
            barrier()
            out["stats_other_exchange"] = d_out.cpu().numpy()[0].copy()
        launches0 = eng.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        tw0 = time.time()
        for a, b in evs:
            flush_l2()
            a.record(stream)
            step_device()
            b.record(stream)
        barrier()
        tw1 = time.time()
        out.update(launches=eng.launch_count() - launches0, wall=(tw0, tw1),
                   step_ms=np.array([a.elapsed_time(b) for a, b in evs]))
        # kernel-only duration for the roofline (no exchange, flushed L2, per-launch events)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for ka, kb in kev:
            flush_l2()
            ka.record(stream)
            eng.loglik_device(kid, d_prm.data_ptr(), 1, d_out.data_ptr(), stream.cuda_stream)
            kb.record(stream)
        barrier()
        out["kern_ms"] = float(np.median([ka.elapsed_time(kb) for ka, kb in kev]))
        if want_extras:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                step_device()
            b.record(stream)
            barrier()
            out["warm_ms"] = a.elapsed_time(b) / steps
            out["stream"], out["d_bufs"] = stream, (d_prm, d_out)
        return out

    # the library's kernels are loaded lazily at their first launch: a 2000-row model pays that once, untimed, so that
    # setup_s / knn_build_s below are the workload's own
    ws_, wy_ = synthetic(2000, cfg["D"], 99)
    for dev_ in (range(ngpu) if single_process else (local,)):
        we_ = _lib.Engine(dev_, args.dtype)
        we_.set_data(ws_, wy_)
        we_.build_neighbors_grid(cfg["m"])
        we_.loglik(KIDS[cfg["kernel"]], np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]))
        we_.close()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    M = measure(cfg, args.steps, True)
    clocks = sampler.stop(*M["wall"]) if rank == 0 else None
    model, eng, s, y, kid = M["model"], M["model"]._engine, M["s"], M["y"], M["kid"]
    # value: the median step, the slowest rank's; the mean is kept beside it
    ms_per_step = max_over_ranks(float(np.median(M["step_ms"])))
    ms_per_step_mean = max_over_ranks(float(np.mean(M["step_ms"])))
    kern_ms = M["kern_ms"] if M["kern_ms"] is not None else ms_per_step

    # ---- cfg5: a sweep of K parameter vectors in one launch (pair distances shared by the vectors) ----
    from pynngp_b200.synthetic import sweep_params

    Ks = 64
    sweep = None
    if not single_process:
        stream = M["stream"]
        d_prmK = torch.tensor(sweep_params(Ks), dtype=torch.float64, device="cuda")
        d_outK = torch.zeros((Ks, 3), dtype=torch.float64, device="cuda")

        def sweep_device():
            if M["fused_exchange"]:
                eng.loglik_device_allreduce(kid, d_prmK.data_ptr(), Ks, d_outK.data_ptr(), stream.cuda_stream)
            else:
                eng.loglik_device(kid, d_prmK.data_ptr(), Ks, d_outK.data_ptr(), stream.cuda_stream)
                if world > 1:
                    dist.all_reduce(d_outK)

        sweep_device()
        barrier()
        sw = []
        for _ in range(3):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sweep_device()
            b.record(stream)
            barrier()
            sw.append(a.elapsed_time(b))
        sweep_ms = max_over_ranks(float(np.median(sw)))
        sweep = {"K": Ks, "ms_per_launch": sweep_ms, "ms_per_eval": sweep_ms / Ks, "evals_per_s": 1e3 * Ks / sweep_ms,
                 "what": "K parameter vectors (synthetic.sweep_params) in ONE launch: distances built once per location"}

    # ---- end to end through the public API (host parameters in, host statistics out, every step) ----------
    for _ in range(5):
        model.loglik_terms()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_terms = model.loglik_terms(PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    if world > 1:
        dist.barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

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
    e2e_y_s = max_over_ranks(time.perf_counter() - t0)
    eng.set_y(y_alt[0])

    # ---- cold path through the public API: host arrays -> upload -> stage 1 -> one evaluation ----
    cold, cold_knn = [], []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        mdl = NNGP(s, y, 0.0, "S=T", cfg["m"], M["spec"], dtype=args.dtype, devices=devices)
        cold_terms = mdl.loglik_terms()
        cold.append(time.perf_counter() - t0)
        cold_knn.append(mdl._timings["knn_s"])
        mdl.close()
        del mdl
    cold_s = max_over_ranks(min(cold))
    cold_knn_s = max_over_ranks(min(cold_knn))

    # ---- BASELINE.json configs[1] (n = 1e5, m = 15, exponential, 1 GPU): neighbour search + likelihood ----------------
    cfg2 = None
    if ngpu == 1 and args.config == "cfg3" and not args.no_cfg4 and args.dtype == "float64":
        c2 = CONFIGS["cfg2"]
        s2, y2 = synthetic(c2["n"], c2["D"], c2["seed"])
        spec2 = Exponential(**PARAMS)
        NNGP(s2[:2000], y2[:2000], 0.0, "S=T", c2["m"], spec2, dtype=args.dtype, devices=devices).close()  # kernels loaded
        both, phases = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            m2 = NNGP(s2, y2, 0.0, "S=T", c2["m"], spec2, dtype=args.dtype, devices=devices)
            t1 = time.perf_counter()
            terms2 = m2.loglik_terms()
            both.append(time.perf_counter() - t0)
            phases.append({"ctor_ms": (t1 - t0) * 1e3, "upload_ms": m2._timings["upload_s"] * 1e3, "knn_ms": m2._timings["knn_s"] * 1e3,
                           "first_eval_ms": (both[-1] - (t1 - t0)) * 1e3})
            if len(both) < 5:
                m2.close()
        e2 = m2._engine
        knn2, ev2 = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            e2.build_neighbors_grid(c2["m"])
            knn2.append(time.perf_counter() - t0)
        for _ in range(20):
            t0 = time.perf_counter()
            m2.loglik_terms()
            ev2.append(time.perf_counter() - t0)
        from oracle import nngp_oracle as orc  # the checker

        want2 = orc.c_loglik(s2, y2, m2._table, KIDS[c2["kernel"]], PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], threads=os.cpu_count() or 1)
        cfg2 = {"workload": workload_config(c2, 1)["workload"].replace(", neighbours prebuilt", ""),
                "search_plus_likelihood_ms": min(both) * 1e3, "phases_of_that_run": phases[int(np.argmin(both))], "search_ms": float(np.median(knn2)) * 1e3,
                "likelihood_ms": float(np.median(ev2)) * 1e3, "stats": list(terms2), "rel_err_vs_oracle": _stat_rel(terms2, want2),
                "what": "pyNNGP.NNGP(t, y, eps, 'S=T', m, cov) from host arrays + one loglik_terms() (best of 5, wall clock, allocations included); the search "
                        "alone through nngp_build_neighbors_grid and one evaluation alone (medians, wall clock through the API)"}
        m2.close()

    # ---- BASELINE.json configs[3] (n = 1e7, m = 30, 3-D): the north_star's scaling target, at every N ------
    cfg4 = None
    if args.config == "cfg3" and not args.no_cfg4 and args.dtype == "float64":
        c4 = CONFIGS["cfg4"]
        M4 = measure(c4, 10, False)
        ms4 = max_over_ranks(float(np.median(M4["step_ms"])))
        k4 = max_over_ranks(M4["kern_ms"]) if M4["kern_ms"] is not None else ms4
        cfg4 = {"workload": workload_config(c4, ngpu)["workload"], "ms_per_eval": ms4, "evals_per_s": 1e3 / ms4, "kernel_ms": k4,
                "knn_build_s": M4["knn_s"], "setup_s": M4["setup_s"], "steps": 10, "stats": M4["stats"].tolist(),
                "exchange_rel_diff": _stat_rel(M4["stats_other_exchange"], M4["stats"]) if "stats_other_exchange" in M4 else None}
        if rank == 0:
            e4 = M4["model"]._engine
            nloc4 = (M4["model"]._shard[1] - M4["model"]._shard[0]) if not single_process else -(-c4["n"] // ngpu)
            ipl4 = fp64_instr_per_location(c4["m"], c4["D"], c4["kernel"])
            peak4 = e4.measure_fma_peak(args.dtype, 4096)
            cfg4["roofline_frac"] = nloc4 * ipl4 / (k4 * 1e-3) / peak4
            # parity of the numbers just timed: a 20 000-row slab of rank 0's shard against the oracle
            from oracle import nngp_oracle as orc  # the checker

            lo4, hi4 = 100000, 120000
            rows4 = e4.get_neighbor_rows(lo4, hi4)
            want4 = orc.c_loglik_rows(M4["s"], M4["y"], rows4, lo4, KIDS[c4["kernel"]], PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"],
                                      threads=os.cpu_count() or 1)
            one4 = _lib.Engine(local, args.dtype)
            one4.set_data(M4["s"][:hi4], M4["y"][:hi4])
            pad = np.full((hi4, c4["m"]), -1, dtype=np.int32)
            pad[lo4:hi4] = rows4
            one4.set_neighbors(pad)
            one4.set_shard(lo4, hi4)
            got4 = one4.loglik(KIDS[c4["kernel"]], np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]))[0]
            one4.close()
            cfg4["parity"] = {"slab_rows": [lo4, hi4], "rel_err_vs_oracle": _stat_rel(got4, want4),
                              "what": "the timed table's rows [lo, hi) evaluated by the fused kernel against oracle/nngp_oracle.c"}
        M4["model"].close()
        del M4

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the fused kernel --------------------------------------------------------------
    nloc = (model._shard[1] - model._shard[0]) if not single_process else -(-cfg["n"] // ngpu)
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
    traffic, traffic_source, units_seen = None, None, None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path) and args.config == "cfg3" and ngpu == 1 and args.dtype == "float64":
        tr = json.load(open(tr_path))
        traffic = tr.get("fused_dram_bytes_per_launch")
        # the same capture's view of the units the kernel loads (constants, like `traffic`): the shared-memory data
        # stage runs closer to its peak than the FP64 pipe the roofline is quoted against
        units_seen = {k: tr.get(k) for k in ("smem_data_stage_pct_of_peak", "smem_wavefronts_per_launch", "fp64_pipe_pct_active",
                                             "issue_slots_pct_active")}
        traffic_source = ("constant read from profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                          f"capture of this workload's kernel ({tr.get('source', 'see profiles/README.md')}); not measured in this run")
    roofline = {
        "bound": "fp64" if args.dtype == "float64" else "fp32",
        "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
        "traffic": traffic, "traffic_source": traffic_source, "ncu_units": units_seen,
        "kernel": "nngp_fused::fused_loglik_kernel", "kernel_ms": kern_ms,
        "algorithmic_instr_per_location": ipl, "locations_per_launch": nloc,
        "peak_source": "measured here: register-resident FMA chains, nngp_measure_fma_peak (no vector-pipe figure in MEASURED_PEAKS.json)",
        "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "bytes_per_location": hbm_bytes_per_location(cfg["m"], cfg["D"]),
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- parity of the numbers being timed + CPU baseline: the oracle over ALL n rows on this host -------------
    from oracle import nngp_oracle as orc  # the checker; also the timed CPU baseline below

    threads = os.cpu_count() or 1
    table = model._table  # one rank alone may read it (a rank of a sharded run searches the missing rows itself)
    prm = (PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"])
    reps, spent, want = 0, 0.0, None
    budget = args.cpu_seconds if args.cpu_seconds is not None else 10.0
    while spent < budget or reps == 0:
        t0 = time.perf_counter()
        want = orc.c_loglik(s, y, table, kid, *prm, threads=threads)
        spent += time.perf_counter() - t0
        reps += 1
    parity = {"rows": [0, int(cfg["n"])], "rel_err_vs_oracle": _stat_rel(M["stats"], want), "n_bad": [float(M["stats"][2]), float(want[2])],
              "e2e_rel_err_vs_oracle": _stat_rel(e2e_terms, want),
              "tolerance": 1e-10 if args.dtype == "float64" else 1e-4,
              "exchange_rel_diff": _stat_rel(M["stats_other_exchange"], M["stats"]) if "stats_other_exchange" in M else None,
              "what": "the statistics of the timed evaluations (all n rows, summed over the ranks) against oracle/nngp_oracle.c on the "
                      "same table; exchange_rel_diff = fused peer-memory exchange vs NCCL all_reduce of the same launch"}
    limits = None
    try:
        from threadpoolctl import threadpool_info

        limits = [{k: d.get(k) for k in ("user_api", "internal_api", "num_threads")} for d in threadpool_info()]
    except Exception:
        pass
    cpu_baseline = {"value": reps / spent, "unit": "evals/s", "cores": threads, "kind": "port", "extrapolated": False,
                    "sample": f"all {cfg['n']} rows x {reps} full evaluations ({spent:.1f} s) of oracle/nngp_oracle.c, {threads} threads "
                              "(ThreadPoolExecutor over row chunks; no BLAS inside)",
                    "os_cpu_count": os.cpu_count(), "threadpoolctl": limits}
    try:
        cpu_baseline["stage1_reference"] = reference_stage1_timing(cfg)
        cpu_baseline["stage1_reference"]["ours_seconds_at_workload_n"] = M["knn_s"]
    except Exception as exc:  # scikit-learn missing on the host: the likelihood baseline above still stands
        cpu_baseline["stage1_reference"] = {"unavailable": repr(exc)}

    exchange = ("none (1 GPU)" if ngpu == 1 else
                "ONE process, multi-device handle: per-device launcher threads, sum fused into the kernels' tails over peer memory" if single_process else
                "fused into the kernel: P2P stores over NVLink peer memory (CUDA IPC), no NCCL call" if M["fused_exchange"]
                else "NCCL all_reduce of 3 doubles after the kernel")
    detail = {"exchange": exchange, "setup_s": M["setup_s"], "knn_build_s": M["knn_s"], "knn_algo": "grid" if eng.knn_used_grid() else "brute",
              "ms_per_step_mean": ms_per_step_mean, "ms_per_step_min": float(np.min(M["step_ms"])), "warm_l2_ms_per_step": M["warm_ms"],
              "stats": M["stats"].tolist(), "e2e_stats": list(e2e_terms), "sweep_cfg5": sweep,
              "cold_e2e": {"ms": cold_s * 1e3, "knn_ms": cold_knn_s * 1e3, "h2d_bytes": int(cfg["n"]) * 8 * (cfg["D"] + 1), "d2h_bytes": 48,
                           "what": "pyNNGP.NNGP(t, y, eps, 'S=T', m, cov) from host arrays (upload + stage 1) + one "
                                   "loglik_terms(); best of 2", "stats": list(cold_terms)},
              "e2e_y_upload": {"value": args.steps / e2e_y_s, "unit": "evals/s", "h2d_bytes_per_step": 8 * int(cfg["n"]) + 32,
                               "d2h_bytes_per_step": 48, "what": "nngp_set_y (the whole response, pageable host memory) + "
                               "loglik_terms() every step"},
              "cfg2": cfg2, "cfg4": cfg4}
    line = {
        "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "evals/s", "n_gpus": ngpu, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic",
        "config": workload_config(cfg, ngpu), "value_is": "median step, slowest rank",
        "clocks": clocks,
        "e2e": {"value": args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 48,
                "api": "pyNNGP.NNGP.loglik_terms(sigma2, phi, tau2)",
                "transport": "parameters: 32 bytes by value in the kernel arguments; statistics: 3 stamped 16-byte lines written by the "
                             "kernel's last block into mapped pinned host memory and polled by the caller"},
        "gpu_launches": int(M["launches"]),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "detail": detail,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--no-cfg4", action="store_true", help="skip the n = 1e7 block (BASELINE.json configs[3]) that rides along with cfg3")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
