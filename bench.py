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
  cpu_baseline: the NumPy/SciPy oracle (the reference cannot be imported here) on the host cores, bounded sample
N > 1 (torchrun): "replicas only" -- a single dense Cholesky does not shard; every rank evaluates its own
independent problem (e.g. a hyper-parameter sweep), no data-path collective; value = N*K / max-rank time.

``--impl reference`` times the oracle port of the reference's CPU path on the box's host cores (rank 0 only).
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


def scale_to_full(t_build, t_rest, n):
    """Extrapolate a sample at size n to n = 16384: the build scales with n^2, the factorisation with n^3."""
    r = N_FULL / float(n)
    return t_build * r**2 + t_rest * r**3


def pick_sample_n(budget_s):
    """Largest n in {16384, 8192, 4096, 2048} whose predicted oracle time fits the per-step budget."""
    x, y = make_inputs(2, 2048)
    oracle_phases(x, y)  # warm BLAS up
    tb, tr = oracle_phases(x, y)
    for n in (16384, 8192, 4096, 2048):
        r = n / 2048.0
        if tb * r**2 + tr * r**3 <= budget_s:
            return n
    return 2048


def sample_text(steps, n_s, dt):
    txt = f"{steps} x oracle logpdf at n={n_s}, d={D} ({dt:.3f} s each)"
    if n_s != N_FULL:
        txt += f", extrapolated to n={N_FULL}: kernel build x{(N_FULL / n_s) ** 2:.0f} (n^2), Cholesky+solve x{(N_FULL / n_s) ** 3:.0f} (n^3)"
    return txt


def reference_arm(args, rank):
    if rank != 0:
        return
    use_all_host_threads()
    total_budget = 150.0
    n_s = pick_sample_n(total_budget / (args.steps + args.warmup))
    x, y = make_inputs(2, n_s)
    for _ in range(args.warmup):
        oracle_phases(x, y)
    tb = tr = 0.0
    for _ in range(args.steps):
        b, r = oracle_phases(x, y)
        tb += b / args.steps
        tr += r / args.steps
    t_full = scale_to_full(tb, tr, n_s)
    value = 1.0 / t_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "EQ().stretch(2.0)+0.1*Delta(), n=16384, d=8, fp64: kernel build + Cholesky + logpdf",
                   "parallelism": "host cores (numpy/scipy oracle restatement of the reference path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                         "sample": sample_text(args.steps, n_s, tb + tr)},
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
            "slices (49 bits), exact int32 slice products (tcgen05.mma.kind::i8), fp64 recombination -- log-pdf agrees "
            "with the all-DMMA path to ~1e-13 relative (see native_fp64.logpdf_rel_diff); parity bar 1e-10",
    "fp64": "native fp64 everywhere (DMMA tensor cores for every GEMM-shaped update)",
}
ARITHMETIC["int8x7"] = ARITHMETIC["auto"]

# one ncu --set full capture of the emulation kernel (profiles/r01_ncu_oz_gemm_details.csv): lower, M = N = 8192, K = 512
OZ_TRAFFIC_SAMPLE = {"from": "profiles/r01_ncu_oz_gemm_details.csv (ncu --set full, one launch: lower, M=N=8192, K=512, 7 slices)",
                     "dram_bytes": 520.2e6, "algorithmic_bytes": 7 * 8192 * 512 + 2 * 8 * (8192 * 8192 // 2)}


def alt_int8_peak(achieved):
    """The same fraction against 2 x the driver-written bf16 number of MEASURED_PEAKS.json (int8 runs at twice the bf16 rate)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        alt = 2.0 * float(mp["bf16_tflops"])
        return {"peak_2x_bf16_measured_peaks_json": alt, "frac_vs_2x_bf16_measured_peaks_json": achieved / alt}
    except Exception:
        return {}


def int8_peak_tops(dev):
    """Dense int8 tensor-core throughput measured in-run with the library GEMM (torch._int_mm -> cuBLASLt), the int8
    counterpart of MEASURED_PEAKS.json's bf16 entry; falls back to 2 x that entry."""
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
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, "cuBLASLt int8 GEMM 8192^3 (torch._int_mm), best of 10, measured in-run"
    except Exception as exc:  # pragma: no cover
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as fh:
                return 2.0 * json.load(fh)["bf16_tflops"], f"2 x MEASURED_PEAKS.json bf16_tflops (torch._int_mm failed: {exc})"
        except Exception:
            return 4500.0, "nominal dense int8 peak (no measurement available)"


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

    for _ in range(max(args.warmup, 3)):
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

    # the same step on the native fp64 tensor-core path (B.precision = "fp64") and with 8 slices (56-bit operands, i.e. no
    # fewer bits than fp64's 53), for the record
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
        slices = ops._oz_slices()
        if o_launches > 0:
            # int8-slice emulation kernel: the work it does is S (S + 1) / 2 int8 GEMMs per fp64-equivalent GEMM
            n_prod = slices * (slices + 1) // 2
            peak, peak_src = int8_peak_tops(dev)
            eq_tflops = o_flops / (o_ms * 1e-3) / 1e12
            achieved = eq_tflops * n_prod
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s (int8)",
                "frac": (achieved / peak) if peak else None, "traffic": None,
                "kernel": f"gpk::oz_gemm_kernel<{slices}> (UTCIMMA = tcgen05.mma.kind::i8; the K=512 trailing SYRK updates of "
                          f"the Cholesky as {n_prod} exact int8 slice products per fp64 product)",
                "peak_source": peak_src,
                "fp64_equivalent_tflops": eq_tflops, "int8_products_per_fp64_product": n_prod,
                **alt_int8_peak(achieved),
                "launches_per_step": o_launches / prof_steps, "kernel_ms_per_step": o_ms / prof_steps,
                "traffic_sample": OZ_TRAFFIC_SAMPLE,
                "whole_step_tflops_fp64_equivalent": cost(N_FULL) / (ms / args.steps * 1e-3) / 1e12,
            }
        else:
            peak = ops.probe_dmma_tflops()
            achieved = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else None
            roofline = {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved and peak > 0 else None, "traffic": None,
                "kernel": "gpk::gemm_nt_f64_v3_kernel<32,2> (DMMA.8x8x4; the K>=512 trailing SYRK updates of the Cholesky)",
                "peak_source": "fp64 DMMA issue peak measured in-run (gpk_probe_dmma_tflops); MEASURED_PEAKS.json has no fp64 entry",
                "launches_per_step": d_launches / prof_steps, "kernel_ms_per_step": d_ms / prof_steps,
                "traffic_sample": {"from": "profiles/r01_ncu_gemm_f64_v3_details.csv (ncu --set full, one launch: lower, M=N=8192, K=1024)",
                                   "dram_bytes": 727.4e6, "algorithmic_bytes": 604.0e6},
                "whole_step_tflops": cost(N_FULL) / (ms / args.steps * 1e-3) / 1e12,
            }
        # bounded CPU baseline sample: the oracle at the largest n that fits ~25 s
        use_all_host_threads()
        n_s = pick_sample_n(25.0)
        xs_, ys_ = make_inputs(2, n_s)
        tb, tr = oracle_phases(xs_, ys_)
        cpu_baseline = {
            "value": 1.0 / scale_to_full(tb, tr, n_s), "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": sample_text(1, n_s, tb + tr),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "arithmetic": ARITHMETIC.get(S.B.precision, S.B.precision), "precision_mode": S.B.precision,
            "native_fp64": native, "emulated_8_slices": eight,
            "config": {"workload": "EQ().stretch(2.0)+0.1*Delta(), n=16384, d=8, fp64: kernel build + Cholesky + logpdf",
                       "parallelism": "single GPU" if world == 1 else f"replicas only x{world} (independent problems, no data-path collective)",
                       "l2": "working set 2.1 GB per step >> 126 MB L2 (no flush needed)", "n": N_FULL, "d": D},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 8 + y_host.numel() * 8),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
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
