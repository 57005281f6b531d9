#!/usr/bin/env python
"""bench.py — uEI acquisition throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W                     # this framework (CUDA path through the C-ABI)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                              # the CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic candidates (config S): posterior mean / variance of
m=3 attributes at the candidates against n=2000 training points (fused K* + blocked triangular solve), the fused MC uEI kernel
(n_w=64 x L=32), top-20 anchor selection and — for N>1 ranks — `bopl_topk_allgather` (pack kernel, one ncclAllGather of the 20 best
records per rank, merge kernel; no host synchronisation).  value = evals/s = candidates * n_w * L * H / step time, whole job.

Scaling (`--scaling`, default strong): BASELINE.json's multi-GPU configuration is ONE batch of 2^20 candidates sharded over the
GPUs (configs[4]), so by default the 2^20 rows are split with `sharding.shard_bounds` and per-GPU work shrinks as N grows
("scaling": "strong"); `--scaling weak` gives every rank its own 2^20 rows (round 1's lines).

The posterior kernel timed is the LIBRARY DEFAULT (int8x7: the solve's n^2/2 update as exact 55-bit digit products on the int8 tensor
pipe; `--posterior fp64|int8` times the others).  On 1 GPU the line also carries, outside the timed region (`extras`):
all three posterior kernels' launch times and roofline fractions, the peaks measured in this run (cuBLAS DGEMM, int8 tcgen05 issue
loop), the reference's sampling depth H=10, the quadratic / exponential utilities, the value+gradient path at batch size, and the CPU
baselines (vectorised port on all host cores + the reference's loop form)."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "uEI acquisition evals/sec (candidates x MC x theta samples)"
UNIT = "evals/s"
K_ANCHOR = 20
C_K = 13          # flops per K* entry besides the distance: Matern52 polynomial + exp counted as 1
DIGITS = {"int8": 6, "int8x7": 7}


def workload(args):
    return dict(name="S", n=args.n, N=args.candidates, n_w=args.n_w, L=args.L, H=1)


def workload_string(w, utility):
    """Identical in both arms (the driver compares the strings): what is computed, not how it is sampled or sharded."""
    return "S: d=10 m=3 n=%d N=%d n_w=%d L=%d H=1 Matern52-ARD noise=1e-6 %s utility top-%d" % (
        w["n"], w["N"], w["n_w"], w["L"], utility, K_ANCHOR)


def algorithmic_flops_per_candidate(m, n, d, H=1):
    """m*H*[n^2 (triangular solve) + 2n (mean) + 2n (square-sum) + n*(3d + c_K)]."""
    return m * H * (n * n + 2 * n + 2 * n + n * (3 * d + C_K))


def int8_ops_per_launch(S, m, N, n):
    """S (S + 1) / 2 exact digit products per fp64 multiply-add of the n^2/2 update, 2 ops each."""
    return 2.0 * (S * (S + 1) // 2) * m * float(N) * n * n / 2.0


def set_blas_threads():
    """All host cores for the CPU arm's BLAS / LAPACK, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
        return cores, "threadpoolctl"
    except Exception as e:                  # pragma: no cover
        return int(os.environ.get("OMP_NUM_THREADS", cores)), "env (%r)" % (e,)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 100 ms while the timed region runs (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max, self.err = index, False, [], set(), None, None
        self.power, self.power_limit = [], None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
            except Exception:
                self.power_limit = None
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:                      # pragma: no cover
            self.err = repr(e)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w": float(np.median(self.power)) if self.power else None, "power_w_max": float(np.max(self.power)) if self.power else None,
                "power_limit_w": self.power_limit, **({"error": self.err} if self.err else {})}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_run(p, n_sample, steps, warmup):
    """The NumPy/SciPy restatement of the reference's path (oracle/, `kind: port`) on the host cores: evals/s."""
    from oracle import acquisition as oacq
    from tests.helpers import oracle_model
    om = oracle_model(p)
    X = p["Xc"][:n_sample]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        marg = oacq.uei_marginal_vec(om, p["utility"], p["theta"], p["Z"], X, 1, chunk=4096)
        acq = oacq.aggregate(marg, p["use_full_support"], p["prob"])
        oacq.top_k(-acq[:, 0], K_ANCHOR)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return len(X) * p["Z"].shape[0] * len(p["theta"]) / t, t


def cpu_loop_form(p, n_sample):
    """the reference's own algorithmic form — uEI._marginal_acq's Python loops over theta, W and candidates with one
    utility callable invocation per evaluation (uEI.py:64-84) — single process, on a small subsample."""
    from oracle import acquisition as oacq, utilities
    from tests.helpers import oracle_model
    om = oracle_model(p)
    func, _ = utilities.get(p["utility"], p["m"])
    X = p["Xc"][:n_sample]
    t0 = time.perf_counter()
    marg = oacq.uei_marginal_loop(om, func, p["theta"], p["Z"], X, 1, incumbent="stale")
    oacq.aggregate(marg, p["use_full_support"], p["prob"])
    t = time.perf_counter() - t0
    return {"value": len(X) * p["Z"].shape[0] * len(p["theta"]) / t, "unit": UNIT, "cores": 1, "seconds": t,
            "sample": "%d candidates, uEI._marginal_acq loop semantics (uEI.py:64-84): one Python utility call per evaluation" % len(X)}


def run_reference(args):
    from bopl_b200.synthetic import make_problem
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, how = set_blas_threads()
    w = workload(args)
    n_sample = args.cpu_sample
    p = make_problem("S", n=w["n"], N=max(n_sample, args.loop_sample), n_w=w["n_w"], L=w["L"], utility=args.utility)
    value, t = cpu_oracle_run(p, n_sample, max(1, args.steps), args.warmup)
    sample = "%d of the %d candidates per step (n=%d, m=3, n_w=%d, L=%d), vectorised NumPy + LAPACK dtrtrs" % (
        n_sample, w["N"], w["n"], w["n_w"], w["L"])
    cpu = {"value": value, "unit": UNIT, "cores": cores, "blas_threads_set_by": how, "kind": "port", "sample": sample}
    if args.loop_sample > 0:
        cpu["loop_form"] = cpu_loop_form(p, args.loop_sample)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": scaling_mode(args), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(w, args.utility), "sample_per_step": n_sample},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def scaling_mode(args):
    return "weak" if args.scaling == "weak" else "strong"


# ------------------------------------------------------------------------------------------------ GPU arm
def _events(torch, n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def measure_peaks(torch, eng, dev, index=0):
    """The two roofline denominators, measured in this run on this device (each about a second)."""
    import ctypes
    out = {}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(a, b)
    best = 1e30
    for _ in range(10):
        e = _events(torch, 2)
        e[0].record()
        torch.matmul(a, b)
        e[1].record()
        torch.cuda.synchronize()
        best = min(best, e[0].elapsed_time(e[1]))
    out["fp64_dgemm_tflops"] = 2.0 * n ** 3 / (best * 1e-3) * 1e-12
    out["fp64_dgemm_how"] = "torch.matmul fp64 8192^3 (cuBLAS), best of 10, CUDA events"
    del a, b
    torch.cuda.empty_cache()
    # int8 tensor pipe: back-to-back 150 ms launches of the issue loop for ~3 s.  "burst" = the fastest launch (a kernel timed alone, cool
    # chip), "sustained" = the mean of the last second (a kernel inside a long step: the board sits at its 1 kW power cap and the SM
    # clock settles well below 1965 MHz, exactly as MEASURED_PEAKS.json reports for the cuBLAS bf16 GEMM)
    tops, ms = ctypes.c_double(), ctypes.c_double()
    rates, t_end = [], time.perf_counter() + 3.0
    sampler = ClockSampler(index)
    sampler.start()
    while time.perf_counter() < t_end:
        eng._check(eng.lib.bopl_probe_int8_peak(eng.h, 150.0, ctypes.byref(tops), ctypes.byref(ms)))
        rates.append((time.perf_counter(), tops.value))
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    out["int8_probe_clocks"] = sampler.summary()
    last = [r for t, r in rates if t >= rates[-1][0] - 1.0]
    out["int8_tops_burst"] = max(r for _, r in rates)
    out["int8_tops_sustained"] = float(np.mean(last))
    out["int8_how"] = ("bopl_probe_int8_peak: back-to-back tcgen05.mma kind::i8 M=128 N=256 K=32 on resident operands, all SMs, %d launches of "
                       "~150 ms over 3 s; burst = fastest launch, sustained = mean of the last second (power cap)" % len(rates))
    return out


def run_gpu(args):
    import torch
    from bopl_b200.synthetic import make_problem
    from bopl_b200.models import MultiOutputGP, state_from_hyperparameters
    from bopl_b200.acquisition import uEI
    from bopl_b200.anchor_points import BoxSpace
    from bopl_b200.utility import Utility, UtilityDistribution
    from bopl_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: anything a library prints to fd 1 meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GPU arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    w = workload(args)
    mode = scaling_mode(args)
    N_total = w["N"] * (world if mode == "weak" else 1)
    start, stop = sharding.shard_bounds(N_total, world, rank)
    N = stop - start                                     # this rank's rows of the global candidate matrix
    p = make_problem("S", n=w["n"], N=16, n_w=w["n_w"], L=w["L"], utility=args.utility)
    d, m = p["d"], p["m"]
    state = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_winv=False,
                                       want_linv=False)
    model = MultiOutputGP.from_state(state, device=local_rank, posterior_kernel=args.posterior)
    eng = model.engine
    kernel_name = eng.posterior_kernel
    assert args.posterior is None or kernel_name == args.posterior

    def make_acq(mdl, pp):
        th = pp["theta"]
        ud = UtilityDistribution(support=th, prob_dist=np.ones(len(th)) / len(th), use_full_support=False,
                                 prior_sample_generator=lambda k, seed=None: th[:k])
        a = uEI(mdl, BoxSpace([(0.0, 1.0)] * d), utility=Utility(func=None, parameter_distribution=ud, kind=pp["utility"]))
        a.W_samples = pp["Z"]
        a.utility_parameter_samples = th
        return a

    theta = p["theta"]
    acq = make_acq(model, p)

    # rows [start, stop) of the global candidate matrix, pinned on the host.  Every rank draws its rows from its own stream so that
    # no rank has to materialise the whole matrix; the global row index is start + local index.
    Xh = torch.empty((N, d), dtype=torch.float64).pin_memory()
    Xh.numpy()[:] = np.random.RandomState(3000 + rank).uniform(size=(N, d))
    acq_h = torch.empty(N, dtype=torch.float64).pin_memory()
    index_offset = start
    gatherer = sharding.TopkGatherer(eng) if world > 1 else None

    # device-resident inputs for `value`
    Xd = Xh.to(dev)
    Zd, thd = eng.to_dev(p["Z"]), eng.to_dev(theta)
    wd = eng.to_dev(np.full(len(theta), 1.0 / len(theta)))
    best = eng.incumbent_dev(args.utility, thd)[:1].contiguous()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    scale = 1.0 / w["n_w"]
    post_ev, coll_ev = [], []

    def step(timed):
        flush.fill_(0.0)                                     # 256 MB write: evicts L2 (126 MB) between steps
        if timed:
            e = _events(torch, 4)
            e[0].record()
        mean, var = eng.posterior_dev(Xd, 0)                 # kernels (1)+(2): fused K* + TRSM + mean + variance
        if timed:
            e[1].record()
            post_ev.append((e[0], e[1]))
        a = eng.uei_mc_dev(mean, var, Zd, args.utility, thd, wd, best, scale)  # kernel (3)
        val, idx = eng.topk_dev(a, min(K_ANCHOR, N), index_offset)
        if gatherer is not None:
            if timed:
                e[2].record()
            val, idx, _ = gatherer.allgather_dev(val, idx, Xd, index_offset, K_ANCHOR)      # pack + ncclAllGather + merge, stream-ordered
            if timed:
                e[3].record()
                coll_ev.append((e[2], e[3]))
        return val, idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    t0, t1 = _events(torch, 2)
    t0.record()
    for _ in range(args.steps):
        val, idx = step(True)
    t1.record()
    barrier()
    launches = eng.launch_count - launches0                 # this library's kernels only (the L2-flush fill is torch's)
    ms_local = t0.elapsed_time(t1) / args.steps
    post_local = float(np.mean([a.elapsed_time(b) for a, b in post_ev]))
    coll_local = float(np.mean([a.elapsed_time(b) for a, b in coll_ev])) if coll_ev else 0.0
    stats = torch.tensor([ms_local, post_local, coll_local], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allst = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        per_rank = torch.stack(allst).cpu().numpy()
        stats = torch.from_numpy(per_rank.max(axis=0))
    ms_step, post_ms, coll_ms = [float(x) for x in stats.tolist()]

    # end to end through the public API with HOST buffers: pinned candidates in, scores + anchors out (+ the exchange for N > 1)
    e2e_steps = max(1, args.e2e_steps if args.e2e_steps > 0 else args.steps)
    acq.top_candidates(Xh, min(K_ANCHOR, N), index_offset)          # warm (allocates the staging buffers)
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(0.0)
        v, i, _ = acq.top_candidates(Xh, min(K_ANCHOR, N), index_offset, want_acq=True, out=acq_h)
        if gatherer is not None:
            gatherer.allgather_host(v, i, Xh.numpy()[i - index_offset], K_ANCHOR)
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - te) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    evals_per_step = float(N_total) * w["n_w"] * w["L"] * w["H"]
    out = None
    if rank == 0:
        extras = {}
        peaks_committed = {"fp64_dgemm_tflops": float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))["fp64_dgemm_tflops"]),
                           "dmma_issue_tflops": float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))["microbench"]["dmma_acc8_tpb256_tflops"]),
                           "int8_tops": float(json.load(open(os.path.join(ROOT, "profiles", "r01_int8_peaks.json")))["int8_tops_N256"]),
                           "source": "profiles/r01_fp64_peaks.json, profiles/r01_int8_peaks.json (round 1, another lease)"}
        peaks_now = None
        if world == 1 and not args.no_extras:
            peaks_now = measure_peaks(torch, eng, dev, local_rank)
        peak_tf = peaks_now["fp64_dgemm_tflops"] if peaks_now else peaks_committed["fp64_dgemm_tflops"]
        # the posterior kernel is timed INSIDE a long step (back-to-back 180 ms launches at the power cap): the sustained figure is its
        # denominator (B200_PROFILING.md); the burst figure is reported beside it (roofline.frac_of_burst_peak)
        peak_i8 = peaks_now["int8_tops_sustained"] if peaks_now else peaks_committed["int8_tops"]
        peak_i8_burst = peaks_now["int8_tops_burst"] if peaks_now else peaks_committed["int8_tops"]
        peak_src = ("measured in this run (roofline.peak_measured_this_run): int8 SUSTAINED issue rate under the power cap; DGEMM best of 10"
                    if peaks_now else peaks_committed["source"])
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "trsm_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = {"fp64": "dram_bytes_per_launch", "int8": "dram_bytes_per_launch_int8", "int8x7": "dram_bytes_per_launch_int8x7"}[kernel_name]
            if tj.get(key) is not None and N == w["N"]:
                traffic, traffic_src = tj[key], "profiles/trsm_traffic.json (ncu --set full capture of this kernel at N=2^20, not this run)"

        def roofline_of(kind, kernel_ms, n_cand):
            flops = algorithmic_flops_per_candidate(m, w["n"], d) * float(n_cand)
            fp64_equiv = flops / (kernel_ms * 1e-3) * 1e-12
            if kind == "fp64":
                return {"bound": "tensor", "kernel": "posterior_trsm_t_kernel (fused K*, FP64 DMMA blocked triangular solve, mean, variance)",
                        "achieved": fp64_equiv, "peak": peak_tf, "unit": "TFLOP/s", "frac": fp64_equiv / peak_tf, "kernel_ms": kernel_ms,
                        "flops_per_launch": flops}
            S = DIGITS[kind]
            ops = int8_ops_per_launch(S, m, n_cand, w["n"])
            ach = ops / (kernel_ms * 1e-3) * 1e-12
            r = {"bound": "tensor",
                 "kernel": "posterior_ozaki_kernel<%d> (fused K*, blocked triangular solve with the n^2/2 update as %d exact int8 digit products "
                           "per fp64 multiply-add on tcgen05.mma kind::i8 / TMEM, fp64 diagonal solve on DMMA, mean, variance)" % (S, S * (S + 1) // 2),
                 "achieved": ach, "peak": peak_i8, "unit": "TFLOP/s", "frac": ach / peak_i8, "frac_of_burst_peak": ach / peak_i8_burst,
                 "unit_note": "int8 tensor tera-ops/s (2 per multiply-add); algorithmic ops = 2 * %d * m * N * n^2/2" % (S * (S + 1) // 2),
                 "kernel_ms": kernel_ms, "int8_ops_per_launch": ops, "flops_per_launch": flops,
                 "fp64_equivalent": {"achieved": fp64_equiv, "peak": peak_tf, "unit": "TFLOP/s", "frac": fp64_equiv / peak_tf,
                                     "note": "the same algorithmic fp64 flops against the cuBLAS DGEMM peak the fp64 kernel is bound by"}}
            # The FP64 pipe and the int8 tensor pipe time-share on this part (profiles/r01_fp64_vs_int8.json), so the launch is bounded by the
            # SUM of its int8 time at the int8 peak and its fp64 time at the fp64 peak: K*, mean, square-sum and the 64-row diagonal solves
            # stay in fp64 (m N [64 n + 4 n + n (3d + c_K)] flops), the rest of the n^2 triangular solve runs as digit products.
            fp64_part = m * float(n_cand) * (64.0 * w["n"] + 4.0 * w["n"] + w["n"] * (3 * d + C_K))
            dmma = peaks_committed["dmma_issue_tflops"]
            t_i8, t_f64 = ops / (peak_i8 * 1e12) * 1e3, fp64_part / (dmma * 1e12) * 1e3
            r["time_shared_bound"] = {"int8_ms": t_i8, "fp64_ms": t_f64, "bound_ms": t_i8 + t_f64, "frac": (t_i8 + t_f64) / kernel_ms,
                                      "fp64_flops_per_launch": fp64_part, "fp64_peak": dmma,
                                      "note": "launch time against (int8 ops at the int8 peak) + (fp64 flops at the DMMA issue peak): the two pipes time-share"}
            return r

        roof = roofline_of(kernel_name, post_ms, N)
        roof["traffic"], roof["traffic_source"] = traffic, traffic_src
        roof["peak_source"] = peak_src
        roof["peak_committed"] = peaks_committed
        if peaks_now:
            roof["peak_measured_this_run"] = peaks_now

        if world == 1 and not args.no_extras:
            extras = run_extras(args, torch, dev, local_rank, p, state, model, Xd, flush, roofline_of, make_acq)

        out = {
            "metric": METRIC, "value": evals_per_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": mode, "vs_baseline": None,
            "dtype": "f64",
            "dtype_note": "every result is f64; with the int8x7 (default) / int8 posterior kernels the triangular solve's n^2/2 update is carried "
                          "out as exact int8 digit products with s32 accumulation on 55- / 47-bit fixed-point splits of its f64 operands (DESIGN.md 3.1b)",
            "data": "synthetic",
            "config": {"workload": workload_string(w, args.utility),
                       "posterior_kernel": kernel_name, "posterior_kernel_is_library_default": args.posterior is None,
                       "candidates_total": N_total, "candidates_per_rank": N,
                       "sharding": "candidates row-sharded with sharding.shard_bounds; bopl_topk_allgather (one ncclAllGather of top-%d records per rank, "
                                   "device merge, no host sync)" % K_ANCHOR if world > 1 else "single GPU",
                       "l2": "256 MB flush write between steps (L2 is 126 MB)"},
            "clocks": sampler.summary(),
            "e2e": {"value": evals_per_step / float(e2e_t.item()), "unit": UNIT,
                    "h2d_bytes_per_step": int(N * d * 8 + (p["Z"].size + theta.size + 2 * len(theta)) * 8 + (K_ANCHOR * (2 + d) * 8 if world > 1 else 0)),
                    "d2h_bytes_per_step": int(N * 8 + K_ANCHOR * 16 + (K_ANCHOR * (2 + d) * 8 if world > 1 else 0)), "steps": e2e_steps,
                    "api": "uEI.top_candidates -> bopl_uei_host (pinned host candidates -> scores + top-20 on the host)" +
                           (" -> TopkGatherer.allgather_host -> bopl_topk_allgather_host" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if world > 1:
            out["per_rank"] = {"ms_per_step": per_rank[:, 0].tolist(), "posterior_kernel_ms": per_rank[:, 1].tolist(),
                               "collective_ms": per_rank[:, 2].tolist(),
                               "collective": "pack kernel + ncclAllGather (%d bytes per rank) + merge kernel, CUDA events on the compute stream" % (K_ANCHOR * (2 + d) * 8)}
        if extras:
            out["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores, how = set_blas_threads()
            pc = make_problem("S", n=w["n"], N=max(args.cpu_sample, args.loop_sample), n_w=w["n_w"], L=w["L"], utility=args.utility)
            v, t = cpu_oracle_run(pc, args.cpu_sample, 1, 0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "blas_threads_set_by": how, "kind": "port",
                                   "sample": "%d of the %d candidates, one pass of the NumPy/SciPy (LAPACK dtrtrs) restatement, %.1f s"
                                             % (args.cpu_sample, w["N"], t)}
            if args.loop_sample > 0:
                out["cpu_baseline"]["loop_form"] = cpu_loop_form(pc, args.loop_sample)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if gatherer is not None:
        gatherer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_extras(args, torch, dev, local_rank, p, state, model, Xd, flush, roofline_of, make_acq):
    """Sub-records of the 1-GPU line, all OUTSIDE the timed region (each a few launches, CUDA events, L2 flush between launches)."""
    from bopl_b200.synthetic import make_problem
    from bopl_b200.models import MultiOutputGP
    w = workload(args)
    N, d, m = Xd.shape[0], p["d"], p["m"]
    ex = {}

    def time_calls(fn, reps=2, warm=1):
        for _ in range(warm):
            fn()
        ms = []
        for _ in range(reps):
            flush.fill_(0.0)
            e = _events(torch, 2)
            e[0].record()
            fn()
            e[1].record()
            torch.cuda.synchronize()
            ms.append(e[0].elapsed_time(e[1]))
        return float(np.mean(ms))

    # (1) the three posterior kernels on the same inputs (+ the default kernel launched one CTA per tile instead of as row-paired clusters):
    # launch time, roofline fraction, SM clock and board power while they run, variance difference from the fp64 kernel
    models = {model.engine.posterior_kernel: model}
    for kind in ("int8x7", "fp64", "int8"):
        if kind not in models:
            models[kind] = MultiOutputGP.from_state(state, device=local_rank, posterior_kernel=kind)
    saved = os.environ.get("BOPL_OZ_CLUSTER")
    os.environ["BOPL_OZ_CLUSTER"] = "0"                   # read by bopl_create
    models["int8x7_single_cta"] = MultiOutputGP.from_state(state, device=local_rank, posterior_kernel="int8x7")
    if saved is None:
        del os.environ["BOPL_OZ_CLUSTER"]
    else:
        os.environ["BOPL_OZ_CLUSTER"] = saved
    kernels, var64 = {}, None
    for kind in ("fp64", "int8x7", "int8", "int8x7_single_cta"):
        eng = models[kind].engine
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms = time_calls(lambda: eng.posterior_dev(Xd, 0), reps=4)
        sampler.stop_flag = True
        sampler.join(timeout=2.0)
        _, var = eng.posterior_dev(Xd, 0)
        base = kind.split("_")[0]
        rec = roofline_of(base, ms, N)
        rec = {k: rec[k] for k in ("kernel", "kernel_ms", "achieved", "peak", "unit", "frac") + (("frac_of_burst_peak", "time_shared_bound") if base != "fp64" else ())}
        rec["evals_per_s_if_it_were_the_step"] = float(N) * w["n_w"] * w["L"] / (ms * 1e-3)
        rec["clocks"] = sampler.summary()
        if kind == "int8x7_single_cta":
            rec["launch"] = "BOPL_OZ_CLUSTER=0: one CTA per candidate tile, 148 solved panels (bit-identical results)"
        if kind == "fp64":
            var64 = var
        else:
            rec["max_rel_variance_difference_from_fp64_kernel"] = float(((var - var64).abs() / var64.abs()).max().item())
        kernels[kind] = rec
        del var
    ex["posterior_kernels"] = kernels
    del var64
    for kind in list(models):
        if models[kind] is not model:
            del models[kind]
    torch.cuda.empty_cache()

    eng = model.engine
    # (2) quadratic / exponential utilities (experiments/test_dtlz2a_quad.py:51-57, test_vlmop3_exp.py:41-52) on the headline kernel
    util = {}
    mean, var = eng.posterior_dev(Xd, 0)
    for kind in ("linear", "quadratic", "exponential"):
        pu = make_problem("S", n=w["n"], N=16, n_w=w["n_w"], L=w["L"], utility=kind)
        Zd, thd = eng.to_dev(pu["Z"]), eng.to_dev(pu["theta"])
        wd = eng.to_dev(np.full(len(pu["theta"]), 1.0 / len(pu["theta"])))
        best = eng.incumbent_dev(kind, thd)[:1].contiguous()
        mc_ms = time_calls(lambda: eng.uei_mc_dev(mean, var, Zd, kind, thd, wd, best, 1.0 / w["n_w"]), reps=3)

        def full():
            mu, vv = eng.posterior_dev(Xd, 0)
            a = eng.uei_mc_dev(mu, vv, Zd, kind, thd, wd, best, 1.0 / w["n_w"])
            eng.topk_dev(a, K_ANCHOR, 0)
        st_ms = time_calls(full)
        util[kind] = {"mc_kernel_ms": mc_ms, "ms_per_step": st_ms, "value": float(N) * w["n_w"] * w["L"] / (st_ms * 1e-3), "unit": UNIT}
    ex["utilities"] = util
    del mean, var

    # (3) the reference's sampling depth: H = 10 hyper-parameter samples (GPyOpt/models/gpmodel.py:33, uEI.py:32), fitted on the device
    H = 10
    rs = np.random.RandomState(77)
    ell = p["lengthscale"][:, :1, :] * rs.uniform(0.85, 1.15, size=(m, H, d))
    var_h = np.repeat(p["variance"][:, :1], H, axis=1) * rs.uniform(0.9, 1.1, size=(m, H))
    gmH = MultiOutputGP(m, kernel=list(p["kinds"]), n_samples=H, device=local_rank, gradients=False)
    gmH.set_hyperparameter_samples(var_h, ell, np.full((m, H), 1e-6), kinds=list(p["kinds"]))
    t0 = time.perf_counter()
    gmH.updateModel(p["X"], list(p["Y"]))
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    eH = gmH.engine
    Zd, thd = eH.to_dev(p["Z"]), eH.to_dev(p["theta"])
    wd = eH.to_dev(np.full(len(p["theta"]), 1.0 / len(p["theta"])))
    bestH = eH.incumbent_dev(p["utility"], thd).contiguous()

    def fullH():
        a = eH.uei_dev(Xd, Zd, p["utility"], thd, wd, bestH, H)
        eH.topk_dev(a, K_ANCHOR, 0)
    msH = time_calls(fullH, reps=2, warm=1)
    ex["H10"] = {"H": H, "ms_per_step": msH, "value": float(N) * w["n_w"] * w["L"] * H / (msH * 1e-3), "unit": UNIT,
                 "posterior_kernel": eH.posterior_kernel, "device_fit_of_30_gps_s": fit_s,
                 "note": "bopl_uei over 10 hyper-samples x 3 attributes (30 factors of n=2000 resident), per-hyper-sample incumbents"}
    del gmH, eH, bestH
    torch.cuda.empty_cache()

    # (4) value + gradient path at batch size (GPy/core/gp.py:464-490: beta = -2 K* Winv, 2 m n^2 flops per candidate)
    from bopl_b200.models import state_from_hyperparameters
    stg = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_winv=True, want_linv=False)
    gmg = MultiOutputGP.from_state(stg, device=local_rank)
    eg = gmg.engine
    Zd, thd = eg.to_dev(p["Z"]), eg.to_dev(p["theta"])
    wd = eg.to_dev(np.full(len(p["theta"]), 1.0 / len(p["theta"])))
    bestg = eg.incumbent_dev(p["utility"], thd)[:1].contiguous()
    peak_tf = roofline_of("fp64", 1.0, 1)["peak"]
    grad = {}
    for Ng in (4096, 65536):
        Xg = Xd[:Ng].contiguous()
        g_ms = time_calls(lambda: eg.posterior_grad_dev(Xg, 0), reps=2)
        f_ms = time_calls(lambda: eg.uei_grad_dev(Xg, Zd, p["utility"], thd, wd, bestg, 1), reps=2)
        flops = 2.0 * m * w["n"] * w["n"] * Ng
        grad["N=%d" % Ng] = {"posterior_grad_ms": g_ms, "uei_value_and_gradient_ms": f_ms, "beta_gemm_flops": flops,
                             "posterior_grad_tflops": flops / (g_ms * 1e-3) * 1e-12, "frac_of_dgemm_peak": flops / (g_ms * 1e-3) * 1e-12 / peak_tf,
                             "candidates_per_s": Ng / (f_ms * 1e-3)}
    grad["note"] = ("posterior_grad = kstar_kernel + beta_gemm_kernel (DMMA, 2 m n^2 flops per candidate) + grad_contract_kernel for d mean and d var; "
                    "the fraction counts only the beta GEMM flops against the DGEMM peak")
    ex["gradient_path"] = grad
    return ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bopl_b200", choices=["bopl_b200", "reference"])
    ap.add_argument("--candidates", type=int, default=2 ** 20, help="rows of the global candidate matrix (per rank with --scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--n_w", type=int, default=64)
    ap.add_argument("--L", type=int, default=32)
    ap.add_argument("--utility", default="linear", choices=["linear", "quadratic", "exponential"])
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--loop-sample", type=int, default=2048, help="candidates of the reference's loop-form CPU figure (0: skip)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: as many as --steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed sub-records (other kernels, peaks, H=10, utilities, gradients)")
    ap.add_argument("--posterior", default=None, choices=["int8x7", "int8", "fp64"],
                    help="posterior kernel; default: the library default (int8x7 = 7-digit exact split on the int8 tensor pipe)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
