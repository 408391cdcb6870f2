#!/usr/bin/env python
"""bench.py -- log-marginal-likelihoods/sec of the GP hot path (BASELINE.json metric).

Workload at N=1 (``config.workload``): BASELINE.json configs[1] -- ``EQ().stretch(l) + s2*Delta()``, n=16384, d=8,
fp64, one logpdf = fused kernel-matrix build -> Cholesky (+ fused triangular solve, log-det) -> finish.
Synthetic inputs: x ~ N(0,1)^{n x d}, l = 2.0, s2 = 0.1, y ~ N(0,1)^n, seed = 2 + rank.

A "step" is one logpdf evaluation through the public API (``GP(k)(x, noise).logpdf(y)``).
  value  : steps/s with x, y already resident in HBM (CUDA-event timed, max over ranks)
  e2e    : the same call with HOST (pinned) x, y: H2D of the inputs and D2H of the scalar inside the timed region
  roofline: in-situ event timing of the dominant kernel -- by default the int8-slice emulation GEMM of the trailing updates
            (tcgen05.mma.kind::i8) vs the int8 GEMM throughput measured in-run; with --precision fp64 the DMMA GEMM vs the
            DMMA peak measured in-run
  native_fp64: the same step with B.precision = "fp64" (all-DMMA), and the relative difference of the two log-pdfs
  cpu_baseline: the NumPy/SciPy oracle (the reference cannot be imported here) on the host cores: ONE call at the full
            n = 16384 on the SAME inputs (no extrapolation); its log-pdf gives parity_vs_oracle_rel for every precision mode
  gpu_library_baseline: the reference's implicit GPU path -- torch eager exp / cholesky / solve_triangular (cuBLAS + cuSOLVER)
            composed as lab.torch composes it -- on the same B200, same inputs, same run
  posterior_solve: f | (f(x), y) at m = 4096 test points (marginals, full covariance) in GFLOP/s
  sharded_c3: BASELINE configs[2] -- B = 512 independent fp32 GPs of n = 2048 batch-sharded over the ranks with ONE scalar
            all-reduce (dist.sharded_logpdf): strong scaling, the only data-path collective the hot path has
N > 1 (torchrun): the headline is "replicas only" -- a single dense Cholesky does not shard; every rank
evaluates its own independent problem (e.g. a hyper-parameter sweep), no data-path collective; value = N*K / max-rank time.

``--impl reference`` times the oracle port of the reference's CPU path on the box's host cores AT n = 16384 (rank 0 only);
the number of timed calls is capped so the run ends within a few minutes, and the cap is printed.
"""
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

N_FULL, D, ELL, S2 = 16384, 8, 2.0, 0.1
METRIC = "log-marginal-likelihoods/sec (n=16384 fp64 EQ)"
UNIT = "logpdf/s"


def make_inputs(seed, n=N_FULL):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D))
    y = rng.standard_normal(n)
    return x, y


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference's numpy path
# ----------------------------------------------------------------------------------------------------------------------
SPEC = ("sum", ("stretched", ELL, ("eq",)), ("scaled", S2, ("delta",)))


def oracle_logpdf(x, y):
    from oracle import gp_oracle as O

    return float(O.fdd_logpdf(SPEC, x, None, y))


def oracle_phases(x, y):
    """One oracle logpdf, timed in its two phases: (kernel-matrix build [O(n^2)], Cholesky + solve + log-det [O(n^3)])."""
    from oracle import gp_oracle as O

    t0 = time.perf_counter()
    K = O.kernel_matrix(SPEC, x)
    t1 = time.perf_counter()
    O.normal_logpdf(None, K, y)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm should still use every core BLAS can take."""
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_threads():
    try:
        import threadpoolctl

        ts = [d.get("num_threads", 1) for d in threadpoolctl.threadpool_info() if d.get("user_api") == "blas"]
        return max(ts) if ts else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def cost(n):
    """Algorithmic flops of one logpdf: n^3/3 + n^2 k + n^2 (3d + 8)."""
    return n**3 / 3.0 + n * n + n * n * (3 * D + 8)


WORKLOAD = "EQ().stretch(2.0)+0.1*Delta(), n=16384, d=8, fp64: kernel build + Cholesky + logpdf"
REF_BUDGET_S = 240.0  # whole --impl reference run: 1 warm-up + as many timed calls as fit (at least 2)


def reference_arm(args, rank):
    """The reference's CPU path (oracle port) at the FULL configuration, n = 16384, on every host core BLAS can use.
    One call is ~45 s on 64 threads, so `--steps 20 --warmup 5` cannot be honoured within a few minutes: the arm does 1
    warm-up call and as many timed calls as fit REF_BUDGET_S (>= 2), and reports exactly what it did -- nothing is
    extrapolated."""
    if rank != 0:
        return
    use_all_host_threads()
    x, y = make_inputs(2)
    t0 = time.perf_counter()
    tb_w, tr_w = oracle_phases(x, y)  # warm-up (also pages BLAS in)
    t_w = time.perf_counter() - t0
    steps = int(max(2, min(args.steps, (REF_BUDGET_S - t_w) // max(t_w, 1e-3))))
    tb = tr = 0.0
    per = []
    for _ in range(steps):
        b, r = oracle_phases(x, y)
        tb += b / steps
        tr += r / steps
        per.append(b + r)
    t_full = tb + tr
    value = 1.0 / t_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "cap": f"one oracle logpdf at n={N_FULL} takes {t_full:.1f} s on these cores: {steps} timed call(s) + 1 warm-up instead of "
               f"{args.steps} + {args.warmup} (budget {REF_BUDGET_S:.0f} s); every call is the full configuration, nothing is extrapolated",
        "config": {"workload": WORKLOAD, "n": N_FULL, "d": D,
                   "parallelism": "host cores (numpy/scipy oracle restatement of the reference path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": f"{steps} x oracle logpdf at the full n={N_FULL}, d={D} ({t_full:.2f} s each: kernel build "
                                   f"{tb:.2f} s + Cholesky/solve/log-det {tr:.2f} s); per-call seconds {[round(t, 2) for t in per]}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


ARITHMETIC = {
    "auto": "fp64 storage and results; kernel build, leaf factorisations, panel solves, finish in native fp64 (FP64 pipe / DMMA); "
            "the K=512 trailing updates emulated on the int8 tensor cores: operands split error-free into 7 signed 7-bit "
            "slices (49 bits), exact int32 slice products (tcgen05.mma.kind::i8), fp64 recombination -- parity_vs_oracle_rel "
            "is the measured distance of this run's log-pdf from the CPU oracle on the same inputs; parity bar 1e-10",
    "fp64": "native fp64 everywhere (DMMA tensor cores for every GEMM-shaped update)",
}
ARITHMETIC["int8x7"] = ARITHMETIC["auto"]
ARITHMETIC["int8x8"] = ARITHMETIC["auto"].replace("7 signed 7-bit slices (49 bits)", "8 signed 7-bit slices (56 bits >= fp64's 53)")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


def kernel_capture(name):
    """One `ncu --set full` capture of a kernel, summarised in profiles/r02_kernel_summaries.json by tools/ncu_summarise.py
    (committed).  bench.py cannot run ncu on itself; `traffic` is therefore the capture's number, labelled as such."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_kernel_summaries.json")) as fh:
            return json.load(fh).get(name)
    except Exception:
        return None


def int8_peak_in_run(dev):
    """Dense int8 tensor-core throughput of the library GEMM (torch._int_mm -> cuBLASLt) measured in this run: the
    SECONDARY denominator (the primary is 2 x MEASURED_PEAKS.json's bf16 entry)."""
    import torch

    try:
        n = 8192
        a = torch.randint(-100, 100, (n, n), device=dev, dtype=torch.int8)
        b = torch.randint(-100, 100, (n, n), device=dev, dtype=torch.int8)
        for _ in range(3):
            torch._int_mm(a, b)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:  # pragma: no cover
        return None


# BASELINE configs[2]
C3_B, C3_N, C3_D, C3_NOISE, C3_EPS = 512, 2048, 8, 0.1, 1e-6


def gpu_arm(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import stheno_b200 as S
    from stheno_b200 import ops
    from stheno_b200.dist import shard_bounds, sharded_logpdf

    S.B.epsilon = 1e-12
    S.B.precision = args.precision
    dev = torch.device("cuda", local_rank)
    x_np, y_np = make_inputs(2 + rank)
    x_dev = torch.as_tensor(x_np, device=dev)
    y_dev = torch.as_tensor(y_np, device=dev)
    x_host = torch.as_tensor(x_np).pin_memory()
    y_host = torch.as_tensor(y_np).pin_memory()
    kernel = S.EQ().stretch(ELL) + S2 * S.Delta()

    def step_resident():
        return S.GP(kernel)(x_dev, None).logpdf(y_dev)

    def step_e2e():
        # public API with HOST buffers: inputs are copied H2D, the scalar result comes back D2H (numpy-like usage)
        lp = S.GP(kernel)(x_host, None).logpdf(y_host)
        return float(lp)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, out, (t0, t1)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        lp = step_resident()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ops.launch_count(reset=True)
    ms, lp, (t0, t1) = timed(step_resident, args.steps)
    launches = ops.launch_count(reset=True)
    clocks = sampler.stop(t0, t1)

    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)

    # the same step in the other precision modes, for the record
    def other_mode(mode, note):
        chosen = S.B.precision
        S.B.precision = mode
        try:
            for _ in range(2):
                step_resident()
            k = min(args.steps, 5)
            ms_m, lp_m, _ = timed(step_resident, k)
        finally:
            S.B.precision = chosen
        return {"value": world * k / (ms_m * 1e-3), "unit": UNIT, "ms_per_step": ms_m / k, "logpdf": float(lp_m),
                "logpdf_rel_diff": abs(float(lp) - float(lp_m)) / abs(float(lp_m)), "note": note}

    native = eight = None
    if S.B.precision != "fp64":
        native = other_mode("fp64", "B.precision='fp64': DMMA trailing updates (python bench.py --precision fp64 makes it the headline); "
                                    "logpdf_rel_diff = headline log-pdf vs this one")
    if S.B.precision in ("auto", "int8x7"):
        eight = other_mode("int8x8", "B.precision='int8x8': the same emulation with 8 slices = 56-bit operands (>= fp64's 53), product "
                                     "error ~1e-15 like the DMMA kernel itself (python bench.py --precision int8x8)")

    # ------------------------------------------------------------------------------------------------------------------
    # BASELINE configs[2] through the one collective of the path: B = 512 x n = 2048 fp32, batch-sharded, scalar all-reduce
    # ------------------------------------------------------------------------------------------------------------------
    def sharded_c3_leg():
        S.B.epsilon = C3_EPS
        try:
            lo, hi = shard_bounds(C3_B, world, rank)
            # problem b is generated from seed 3000 + b: the global batch is the same whatever the number of ranks
            xs_, ys_ = [], []
            for b in range(lo, hi):
                g = np.random.default_rng(3000 + b)
                xs_.append(g.standard_normal((C3_N, C3_D), dtype=np.float32))
                ys_.append(g.standard_normal((C3_N, 1), dtype=np.float32))
            xl = torch.as_tensor(np.stack(xs_), device=dev)
            yl = torch.as_tensor(np.stack(ys_), device=dev)
            make = lambda xb: S.GP(S.EQ())(xb, C3_NOISE)
            run = lambda: sharded_logpdf(make, xl, yl, reduce="sum", presharded=True)
            local = lambda: make(xl).logpdf(yl).sum()
            for _ in range(2):
                tot = run()
            k = max(2, min(args.steps, 5))
            ms_c, tot, _ = timed(run, k)
            ms_l, loc, _ = timed(local, k)  # the same share without the collective (max over ranks)
            # the collective against an independent route: all_gather of the local sums, added on the host in fp64
            loc_v = torch.as_tensor([float(loc)], device=dev, dtype=torch.float64)
            parts = [loc_v]
            if dist is not None:
                parts = [torch.zeros_like(loc_v) for _ in range(world)]
                dist.all_gather(parts, loc_v)
            gathered = float(sum(float(p_.item()) for p_ in parts))
            # parity of this rank's first problem against the oracle (fp64 CPU) -- 1e-4 is north_star's fp32 bar
            from oracle import gp_oracle as O

            ref0 = float(O.fdd_logpdf(("eq",), xs_[0].astype(np.float64), C3_NOISE, ys_[0].astype(np.float64), eps=C3_EPS))
            got0 = float(make(xl[:1]).logpdf(yl[:1]).reshape(-1)[0])
            par = torch.as_tensor([abs(got0 - ref0) / abs(ref0)], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(par, op=dist.ReduceOp.MAX)
            flops = C3_B * (C3_N ** 3 / 3.0 + C3_N ** 2 * (3 * C3_D + 9))
            return {
                "workload": f"B={C3_B} independent EQ GPs, n={C3_N}, d={C3_D}, fp32, noise {C3_NOISE}, epsilon {C3_EPS}: "
                            "kernel build + batched Cholesky + logpdf, summed",
                "scaling": "strong", "n_gpus": world, "batch_per_rank": hi - lo,
                "collective": "ONE all_reduce(SUM) of a single scalar per step (torch.distributed / NCCL); no other data-path exchange",
                "value": C3_B * k / (ms_c * 1e-3), "unit": "batch logpdf/s", "ms_per_step": ms_c / k, "steps": k,
                "ms_per_step_without_collective": ms_l / k,
                "collective_overhead_frac": (ms_c - ms_l) / ms_c,
                "tflops_fp32_equivalent": flops / (ms_c / k * 1e-3) / 1e12,
                "sum_logpdf": float(tot), "sum_via_all_gather_fp64": gathered,
                "sum_rel_diff": abs(float(tot) - gathered) / abs(gathered),
                "parity_vs_oracle_rel_max_over_ranks": float(par.item()), "parity_bar": 1e-4,
            }
        finally:
            S.B.epsilon = 1e-12

    sharded = None
    if not args.no_c3:
        try:
            sharded = sharded_c3_leg()
        except Exception as exc:  # the headline must survive a failure of a secondary leg
            sharded = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    if rank == 0:
        # roofline leg: dominant kernel timed in situ with events on its own stream (look-ahead off for these steps: with
        # it on, the update kernels share the SMs with the side-stream panel kernels and their event-bracketed durations
        # overlap, which would not be a per-kernel figure)
        prof_steps = min(args.steps, 3)
        os.environ["GPK_NO_LOOKAHEAD"] = "1"
        ops.gemm_profile(True)
        for _ in range(prof_steps):
            step_resident()
        d_ms, d_flops, d_launches = ops.gemm_profile_read(0)
        o_ms, o_flops, o_launches = ops.gemm_profile_read(1)
        ops.gemm_profile(False)
        del os.environ["GPK_NO_LOOKAHEAD"]
        slices = ops._oz_slices(True)  # the benchmarked factorisation: noise 0.1 of the variance -> well conditioned by construction
        mp = measured_peaks()
        timing_note = ("CUDA events around every launch of the kernel, inside whole logpdf steps run with GPK_NO_LOOKAHEAD=1: the same "
                       "launches in the same order (pairs of 512-panels, far updates with K = 1024) on ONE stream, so that launches do "
                       "not share SMs and the bracketed durations are per-kernel figures (the timed region of `value` overlaps them "
                       "with the panel factorisations on side streams)")
        if o_launches > 0:
            # int8-slice emulation kernel: the work it does is S (S + 1) / 2 int8 GEMMs per fp64-equivalent GEMM
            n_prod = slices * (slices + 1) // 2
            eq_tflops = o_flops / (o_ms * 1e-3) / 1e12
            achieved = eq_tflops * n_prod
            if mp is not None:
                peak, peak_src = 2.0 * float(mp["bf16_tflops"]), ("2 x MEASURED_PEAKS.json bf16_tflops (driver-measured cuBLAS bf16 burst; "
                                                                  "int8 tcgen05 runs at twice the bf16 rate) -- of measured")
            else:
                peak, peak_src = 2.0 * 1590.0, "2 x 1.59 PFLOP/s (B200_PROFILING.md fallback; MEASURED_PEAKS.json absent) -- of fallback"
            in_run = int8_peak_in_run(dev)
            cap = kernel_capture(f"oz_gemm_kernel<{slices}>")
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s (int8)",
                "frac": achieved / peak, "traffic": (cap or {}).get("dram_bytes"),
                "traffic_source": (cap or {}).get("source", "no committed capture of this kernel (profiles/r02_kernel_summaries.json)"),
                "traffic_algorithmic_bytes": (cap or {}).get("algorithmic_bytes"),
                "kernel": f"gpk::oz_gemm_kernel<{slices}> (UTCIMMA = tcgen05.mma.kind::i8; the trailing SYRK updates of the Cholesky "
                          f"-- K = 1024 per pair of panels for the far part, K = 512 inside a pair -- as {n_prod} exact int8 slice "
                          f"products per fp64 product)",
                "peak_source": peak_src,
                "peak_sustained": (2.0 * float(mp["bf16_tflops_sustained"])) if mp and "bf16_tflops_sustained" in mp else None,
                "frac_vs_sustained": (achieved / (2.0 * float(mp["bf16_tflops_sustained"]))) if mp and "bf16_tflops_sustained" in mp else None,
                "peak_in_run_cublaslt_int8": in_run, "frac_vs_in_run_cublaslt_int8": (achieved / in_run) if in_run else None,
                "fp64_equivalent_tflops": eq_tflops, "int8_products_per_fp64_product": n_prod,
                "launches_per_step": o_launches / prof_steps, "kernel_ms_per_step": o_ms / prof_steps,
                "kernel_share_of_step": (o_ms / prof_steps) / (ms / args.steps),
                "timing": timing_note,
                "whole_step_tflops_fp64_equivalent": cost(N_FULL) / (ms / args.steps * 1e-3) / 1e12,
            }
        else:
            peak = ops.probe_dmma_tflops()
            achieved = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else None
            cap = kernel_capture("gemm_nt_f64_v3_kernel")
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved and peak > 0 else None, "traffic": (cap or {}).get("dram_bytes"),
                "traffic_source": (cap or {}).get("source", "no committed capture of this kernel"),
                "kernel": "gpk::gemm_nt_f64_v3_kernel<32,2> (DMMA.8x8x4; the K>=512 trailing SYRK updates of the Cholesky)",
                "peak_source": "fp64 DMMA issue peak measured in-run (gpk_probe_dmma_tflops); MEASURED_PEAKS.json has no fp64 entry",
                "launches_per_step": d_launches / prof_steps, "kernel_ms_per_step": d_ms / prof_steps,
                "timing": timing_note,
                "whole_step_tflops": cost(N_FULL) / (ms / args.steps * 1e-3) / 1e12,
            }

        # --------------------------------------------------------------------------------------------------------------
        # posterior solve: f | (f(x), y) at m = 4096 test points, factor cached (as the reference
        # caches K_x's factor, observations.py:127-141); flops = n^2 m (TRSM) + 2 n m (mean) + [2 n m | n m^2] (variance)
        # --------------------------------------------------------------------------------------------------------------
        def posterior_leg():
            m = 4096
            xs_dev = torch.as_tensor(np.random.default_rng(22).standard_normal((m, D)), device=dev)
            f = S.GP(kernel)
            post = f | (f(x_dev), y_dev)
            n = N_FULL
            out = {"m": m, "n": n, "note": "factor of K_x cached by the conditioning (not timed); each step = kernel rows "
                                           "k(x*, x) + triangular solve + mean/variance reduction"}
            for name, fn, fl in (
                ("marginals", lambda: post(xs_dev).marginals(), n * n * m + 2 * n * m + 2 * n * m),
                ("full_covariance", lambda: (lambda mv: (mv[0], S.B.dense(mv[1])))(post(xs_dev).mean_var), n * n * m + 2 * n * m + n * m * m),
            ):
                for _ in range(2):
                    fn()
                k = 5
                ms_p, _, _ = timed(fn, k) if dist is None else (None, None, None)
                if ms_p is None:  # N > 1: rank 0 alone times this leg (no barrier inside)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(k):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms_p = e0.elapsed_time(e1)
                out[name] = {"ms": ms_p / k, "gflops": fl / (ms_p / k * 1e-3) / 1e9, "flops": fl}
            return out

        # --------------------------------------------------------------------------------------------------------------
        # the reference's implicit GPU path on the same B200: lab.torch composes torch eager ops = cuBLAS / cuSOLVER kernels
        # ("the existing Blackwell kernels" the hand-written path has to beat, not only the CPU)
        # --------------------------------------------------------------------------------------------------------------
        def library_leg():
            n = N_FULL

            def eager():
                xs_ = x_dev / ELL                                            # k.stretch(l)
                n1 = (xs_ * xs_).sum(-1)
                d2 = n1[:, None] + n1[None, :] - 2.0 * (xs_ @ xs_.T)           # B.pw_dists2 (GEMM expansion, d > 1)
                K = torch.exp(-0.5 * d2)                                     # EQ
                K = K + S2 * torch.eye(n, dtype=K.dtype, device=dev)         # + s2 * Delta()(x) (identity for the same object)
                K = K + 1e-12 * torch.eye(n, dtype=K.dtype, device=dev)      # B.reg
                L = torch.linalg.cholesky(K)                                 # B.cholesky
                a = torch.linalg.solve_triangular(L, y_dev[:, None], upper=False)
                return -0.5 * (2.0 * torch.log(torch.diagonal(L)).sum() + n * np.log(2 * np.pi) + (a * a).sum())

            for _ in range(2):
                v = eager()
            k = 5
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                v = eager()
            e1.record()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / k
            return {"value": 1e3 / ms_l, "unit": UNIT, "ms_per_step": ms_l, "logpdf": float(v),
                    "logpdf_rel_diff": abs(float(v) - float(lp)) / abs(float(lp)),
                    "what": "torch eager on the same GPU and inputs, composed as the reference's lab.torch backend composes the path: "
                            "pw_dists2 by GEMM expansion, exp, + noise, + epsilon, torch.linalg.cholesky, solve_triangular, log-det "
                            "(cuBLAS / cuSOLVER sm_100 kernels; none of this repo's kernels)",
                    "speedup_of_headline": (1e3 / ms_l) and value / world / (1e3 / ms_l)}

        posterior = library = None
        try:
            posterior = posterior_leg()
        except Exception as exc:
            posterior = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        try:
            library = library_leg()
        except Exception as exc:
            library = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

        # CPU baseline: the oracle at the FULL n = 16384 on rank 0's inputs, one call, every host core (N = 1 only: the scaling
        # runs would otherwise spend most of their wall-clock here)
        cpu_baseline, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import gp_oracle as O

            use_all_host_threads()
            xw, yw = make_inputs(7, 2048)
            O.fdd_logpdf(SPEC, xw, None, yw)  # page BLAS in
            tc0 = time.perf_counter()
            K = O.kernel_matrix(SPEC, x_np)
            tc1 = time.perf_counter()
            lp_ref = float(O.normal_logpdf(None, K, y_np))
            tc2 = time.perf_counter()
            del K
            cpu_baseline = {
                "value": 1.0 / (tc2 - tc0), "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                "sample": f"1 x oracle logpdf at the full n={N_FULL}, d={D}, same inputs as the GPU arm ({tc2 - tc0:.2f} s: kernel build "
                          f"{tc1 - tc0:.2f} s + Cholesky/solve/log-det {tc2 - tc1:.2f} s); not extrapolated",
                "logpdf": lp_ref,
            }
            r_ = lambda v: abs(float(v) - lp_ref) / abs(lp_ref)
            parity = {"bar": 1e-10, "oracle_logpdf": lp_ref, S.B.precision: r_(lp)}
            if native is not None:
                parity["fp64"] = r_(native["logpdf"])
            if eight is not None:
                parity["int8x8"] = r_(eight["logpdf"])
            if library is not None and "logpdf" in library:
                parity["gpu_library_baseline"] = r_(library["logpdf"])
        elif world > 1:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                            "sample": "not run at N > 1 (rank 0 at N = 1 only); see the N = 1 line and --impl reference"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "arithmetic": ARITHMETIC.get(S.B.precision, S.B.precision), "precision_mode": S.B.precision,
            "parity_vs_oracle_rel": parity,
            "native_fp64": native, "emulated_8_slices": eight,
            "config": {"workload": WORKLOAD,
                       "parallelism": "single GPU" if world == 1 else f"replicas only x{world} (independent problems, no data-path collective)",
                       "l2": "working set 2.1 GB per step >> 126 MB L2 (no flush needed)", "n": N_FULL, "d": D},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 8 + y_host.numel() * 8),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "gpu_library_baseline": library,
            "posterior_solve": posterior,
            "sharded_c3": sharded,
            "logpdf": float(lp),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp64", "int8x6", "int8x7", "int8x8", "tf32x3"],
                    help="stheno_b200.B.precision for the timed steps (auto = int8x7 emulation of the large fp64 updates)")
    ap.add_argument("--no-c3", action="store_true", help="skip the sharded BASELINE configs[2] leg")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~1 min CPU oracle call at n=16384 (N=1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
