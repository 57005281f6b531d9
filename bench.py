#!/usr/bin/env python
"""bench.py — uEI acquisition throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W                     # this framework (CUDA path through the C-ABI)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                              # the CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic candidates per GPU (config S):
posterior mean/variance of m=3 attributes at N=2^20 candidates against n=2000 training points (fused K* +
blocked FP64-tensor-core triangular solve), the fused MC uEI kernel (n_w=64 x L=32), top-20 anchor selection, and —
for N>1 ranks — the all-gather of the 20 best records per rank.  Candidates shard across ranks (weak scaling: every
rank scores its own 2^20-candidate block of a (gpus x 2^20)-row batch).  value = evals/s = gpus*N*n_w*L*H / step time.
"""
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


def workload(args):
    return dict(name="S", n=args.n, N=args.candidates, n_w=args.n_w, L=args.L, H=1)


def algorithmic_flops_per_candidate(m, n, d, H=1):
    """m*H*[n^2 (triangular solve) + 2n (mean) + 2n (square-sum) + n*(3d + c_K)]."""
    return m * H * (n * n + 2 * n + 2 * n + n * (3 * d + C_K))


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 100 ms while the timed region runs (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max, self.err = index, False, [], set(), None, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
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
                "reasons": sorted(self.reasons), "samples": len(self.sm), **({"error": self.err} if self.err else {})}


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


def run_reference(args):
    from bopl_b200.synthetic import make_problem
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    n_sample = args.cpu_sample
    p = make_problem("S", n=w["n"], N=n_sample, n_w=w["n_w"], L=w["L"], utility=args.utility)
    value, t = cpu_oracle_run(p, n_sample, max(1, args.steps), args.warmup)
    cores = os.cpu_count()
    sample = "%d of the %d candidates per step (n=%d, m=3, n_w=%d, L=%d)" % (n_sample, w["N"], w["n"], w["n_w"], w["L"])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "S: d=10 m=3 n=%d N=%d n_w=%d L=%d H=1 Matern52-ARD %s utility" % (w["n"], w["N"], w["n_w"], w["L"], args.utility)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
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
    N = w["N"]
    p = make_problem("S", n=w["n"], N=16, n_w=w["n_w"], L=w["L"], utility=args.utility)
    d, m = p["d"], p["m"]
    state = state_from_hyperparameters(p["X"], p["Y"], p["kinds"], p["variance"], p["lengthscale"], p["noise"], want_winv=False,
                                       want_linv=False)
    model = MultiOutputGP.from_state(state, device=local_rank, posterior_kernel=args.posterior)
    eng = model.engine
    theta = p["theta"]
    udist = UtilityDistribution(support=theta, prob_dist=np.ones(len(theta)) / len(theta), use_full_support=False,
                                prior_sample_generator=lambda k, seed=None: theta[:k])
    util = Utility(func=None, parameter_distribution=udist, kind=args.utility)
    acq = uEI(model, BoxSpace([(0.0, 1.0)] * d), utility=util)
    acq.W_samples = p["Z"]
    acq.utility_parameter_samples = theta

    # this rank's block of the global candidate matrix (rows rank*N ... rank*N+N-1), pinned on the host
    Xh = torch.empty((N, d), dtype=torch.float64).pin_memory()
    Xh.numpy()[:] = np.random.RandomState(3000 + rank).uniform(size=(N, d))
    acq_h = torch.empty(N, dtype=torch.float64).pin_memory()
    index_offset = rank * N

    # device-resident inputs for `value`
    Xd = Xh.to(dev)
    Zd, thd = eng.to_dev(p["Z"]), eng.to_dev(theta)
    wd = eng.to_dev(np.full(len(theta), 1.0 / len(theta)))
    best = eng.incumbent_dev(args.utility, thd)[:1].contiguous()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    scale = 1.0 / w["n_w"]
    trsm_ms = []

    def step(timed):
        flush.fill_(0.0)                                     # 256 MB write: evicts L2 (126 MB) between steps
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        mean, var = eng.posterior_dev(Xd, 0)                 # kernels (1)+(2): fused K* + TRSM + mean + variance
        if timed:
            e1.record()
            trsm_ms.append((e0, e1))
        a = eng.uei_mc_dev(mean, var, Zd, args.utility, thd, wd, best, scale)  # kernel (3)
        val, idx = eng.topk_dev(a, K_ANCHOR, index_offset)
        if world > 1:
            idx_c = idx.cpu().numpy()
            sharding.allgather_topk(val.cpu().numpy(), idx_c, Xh.numpy()[idx_c - index_offset], K_ANCHOR, None, dev)
        return val, idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the fp64 DMMA kernel on the same inputs, outside the timed region: its launch time and the largest relative difference of the
    # int8 kernel's variance from it over a whole wave of candidates (rank 0 of a 1-GPU run only)
    fp64_ms, parity_rel = None, None
    if args.posterior != "fp64" and world == 1 and not args.no_fp64_leg:
        m64 = MultiOutputGP.from_state(state, device=local_rank, posterior_kernel="fp64")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        m64.engine.posterior_dev(Xd, 0)
        ev[0].record()
        mean64, var64 = m64.engine.posterior_dev(Xd, 0)
        ev[1].record()
        torch.cuda.synchronize()
        fp64_ms = float(ev[0].elapsed_time(ev[1]))
        mean8, var8 = eng.posterior_dev(Xd, 0)
        parity_rel = float(((var8 - var64).abs() / var64.abs()).max().item())
        del m64, mean64, var64, mean8, var8
        torch.cuda.empty_cache()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        val, idx = step(True)
    t1.record()
    barrier()
    launches = eng.launch_count - launches0                 # this library's kernels only (the L2-flush fill is torch's)
    elapsed = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    trsm = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in trsm_ms]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(trsm, op=dist.ReduceOp.MAX)
    ms_step = float(elapsed.item()) / args.steps
    trsm_ms_avg = float(trsm.item())

    # end to end through the public API with HOST buffers: pinned candidates in, scores + anchors out
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    acq.top_candidates(Xh, K_ANCHOR, index_offset)          # warm (allocates the staging buffers)
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        v, i, _ = acq.top_candidates(Xh, K_ANCHOR, index_offset, want_acq=True, out=acq_h)
        if world > 1:
            sharding.allgather_topk(v, i, Xh.numpy()[i - index_offset], K_ANCHOR, None, dev)
    barrier()
    e2e_t = torch.tensor([(time.perf_counter() - te) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    evals_per_step = float(world) * N * w["n_w"] * w["L"] * w["H"]
    out = None
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
        peak_tf = float(peaks["fp64_dgemm_tflops"])
        flops_launch = algorithmic_flops_per_candidate(m, w["n"], d) * float(N)
        fp64_equiv = flops_launch / (trsm_ms_avg * 1e-3) * 1e-12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "trsm_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_int8" if args.posterior != "fp64" else "dram_bytes_per_launch")
        if args.posterior == "fp64":
            roof = {"bound": "tensor", "kernel": "posterior_trsm_t_kernel (fused K*, FP64 DMMA blocked triangular solve, mean, variance)",
                    "achieved": fp64_equiv, "peak": peak_tf, "unit": "TFLOP/s", "frac": fp64_equiv / peak_tf,
                    "traffic": traffic, "kernel_ms": trsm_ms_avg, "flops_per_launch": flops_launch,
                    "peak_source": "profiles/r01_fp64_peaks.json: cuBLAS DGEMM 8192^3 measured on this pool's B200 "
                                   "(MEASURED_PEAKS.json has no FP64 figure; DMMA issue peak 37.0)"}
        else:
            # the triangular-solve update runs as exact int8 digit products on tcgen05: S (S + 1) / 2 products per fp64 multiply-add
            S = 6 if args.posterior == "int8" else 7
            pairs = S * (S + 1) // 2
            int8_ops = 2.0 * pairs * m * float(N) * w["n"] * w["n"] / 2.0
            i8 = json.load(open(os.path.join(ROOT, "profiles", "r01_int8_peaks.json")))
            peak_i8 = float(i8["int8_tops_N256"])
            ach = int8_ops / (trsm_ms_avg * 1e-3) * 1e-12
            roof = {"bound": "tensor",
                    "kernel": "posterior_ozaki_kernel<%d> (fused K*, blocked triangular solve with the n^2/2 update as %d exact int8 digit "
                              "products per fp64 multiply-add on tcgen05.mma kind::i8 / TMEM, fp64 diagonal solve on DMMA, mean, variance)" % (S, pairs),
                    "achieved": ach, "peak": peak_i8, "unit": "TFLOP/s", "frac": ach / peak_i8,
                    "unit_note": "int8 tensor tera-ops/s (2 per multiply-add); algorithmic ops = 2 * %d * m * N * n^2/2" % pairs,
                    "traffic": traffic, "kernel_ms": trsm_ms_avg, "int8_ops_per_launch": int8_ops, "flops_per_launch": flops_launch,
                    "fp64_equivalent": {"achieved": fp64_equiv, "peak": peak_tf, "unit": "TFLOP/s", "frac": fp64_equiv / peak_tf,
                                        "note": "the same algorithmic fp64 flops against the cuBLAS DGEMM peak the fp64 kernel is bound by"},
                    "peak_source": "profiles/r01_int8_peaks.json: back-to-back tcgen05.mma kind::i8 M=128 N=256 K=32 on all SMs "
                                   "(tools/int8_umma_probe.cu); the FP64 pipe time-shares with it (profiles/r01_fp64_vs_int8.json)"}
            # The FP64 pipe and the int8 tensor pipe time-share on this part (profiles/r01_fp64_vs_int8.json), so the launch is bounded by the
            # SUM of its int8 time at the int8 peak and its fp64 time at the fp64 peak: K*, mean, square-sum and the 64-row diagonal solves
            # stay in fp64 (m N [64 n + 4 n + n (3d + c_K)] flops), the rest of the n^2 triangular solve runs as digit products.
            fp64_part = m * float(N) * (64.0 * w["n"] + 4.0 * w["n"] + w["n"] * (3 * d + C_K))
            t_i8, t_f64 = int8_ops / (peak_i8 * 1e12) * 1e3, fp64_part / (float(peaks["microbench"]["dmma_acc8_tpb256_tflops"]) * 1e12) * 1e3
            roof["time_shared_bound"] = {"int8_ms": t_i8, "fp64_ms": t_f64, "bound_ms": t_i8 + t_f64, "frac": (t_i8 + t_f64) / trsm_ms_avg,
                                         "fp64_flops_per_launch": fp64_part, "fp64_peak": float(peaks["microbench"]["dmma_acc8_tpb256_tflops"]),
                                         "note": "measured launch time against (int8 ops at the int8 peak) + (fp64 flops at the DMMA issue peak, the higher of the two fp64 peaks)"}
            if fp64_ms is not None:
                roof["fp64_kernel"] = {"kernel": "posterior_trsm_t_kernel", "kernel_ms": fp64_ms,
                                       "frac_of_dgemm_peak": flops_launch / (fp64_ms * 1e-3) * 1e-12 / peak_tf,
                                       "int8_vs_fp64_max_rel_var": parity_rel}
        out = {
            "metric": METRIC, "value": evals_per_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64",
            "dtype_note": "every result is f64; with --posterior int8 / int8x7 the triangular solve's n^2/2 update is carried out as exact int8 "
                          "digit products with s32 accumulation on 47- / 55-bit fixed-point splits of its f64 operands (DESIGN.md 3.1b)",
            "data": "synthetic",
            "config": {"workload": "S: d=10 m=3 n=%d N=%d candidates per GPU, n_w=%d L=%d H=1 Matern52-ARD noise=1e-6 %s utility, top-%d anchors"
                                   % (w["n"], N, w["n_w"], w["L"], args.utility, K_ANCHOR),
                       "posterior_kernel": args.posterior,
                       "sharding": "candidates row-sharded, %d per rank; all-gather of top-%d records only" % (N, K_ANCHOR),
                       "l2": ("working set (L 3x33.5 MB + V panels 310 MB + candidates 84 MB) > 126 MB L2, and a 256 MB flush write between steps"
                              if args.posterior == "fp64" else
                              "working set (factor digit images 3x%.1f MB + solved-digit panels %d MB + candidates 84 MB) > 126 MB L2, "
                              "and a 256 MB flush write between steps" % ((12.2, 232) if args.posterior == "int8" else (14.2, 271)))},
            "clocks": sampler.summary(),
            "e2e": {"value": evals_per_step / float(e2e_t.item()), "unit": UNIT,
                    "h2d_bytes_per_step": int(N * d * 8 + (p["Z"].size + theta.size + 2 * len(theta)) * 8),
                    "d2h_bytes_per_step": int(N * 8 + K_ANCHOR * 16), "steps": e2e_steps,
                    "api": "uEI.top_candidates -> bopl_uei_host (pinned host candidates -> scores + top-20 on the host)"},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            pc = make_problem("S", n=w["n"], N=args.cpu_sample, n_w=w["n_w"], L=w["L"], utility=args.utility)
            v, t = cpu_oracle_run(pc, args.cpu_sample, 1, 0)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": "%d of the %d candidates, one pass of the NumPy/SciPy (LAPACK dtrtrs) restatement, %.1f s"
                                             % (args.cpu_sample, N, t)}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bopl_b200", choices=["bopl_b200", "reference"])
    ap.add_argument("--candidates", type=int, default=2 ** 20)
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--n_w", type=int, default=64)
    ap.add_argument("--L", type=int, default=32)
    ap.add_argument("--utility", default="linear", choices=["linear", "quadratic", "exponential"])
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--posterior", default="int8", choices=["int8", "int8x7", "fp64"],
                    help="posterior kernel: int8 = 6-digit split on the int8 tensor pipe (default), int8x7 = 7 digits, fp64 = DMMA kernel")
    ap.add_argument("--no-fp64-leg", action="store_true", help="skip the fp64-kernel comparison launch before the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
