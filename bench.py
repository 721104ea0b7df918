#!/usr/bin/env python
"""bench.py -- QR GFLOP/s (2mn^2 - 2n^3/3) of the B200-native hot path, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extra] [--no-e2e] [--no-cpu] [--config1-full]

Workloads (BASELINE.json configs).  The headline line is config 2: 16384 x 16384 blocked Householder QR on one GPU
(at N > 1 every rank factors its own matrix: replicas only).  The same JSON object carries, under the keys
`tsqr`, `batched` and (N > 1) `caqr`: config 3, 8 388 608 x 64 row-partitioned over the N ranks (local step: the Gram leaf on
tcgen05, gram_umma.cu, with the Householder leaf timed beside it) with the R factors
combined over peer memory (one hop into rank 0 up to 8 ranks; strong scaling; `tsqr.efficiency_vs_ideal` from the in-run per-rank leaf time); config 4,
65 536 independent 64 x 64 matrices split across ranks; config 5, CAQR with 16384 rows per rank, with its three
acceptance numbers.  `--no-extra` drops those, `--no-e2e` / `--no-cpu` drop the host-buffer and CPU legs.
`cpu_baseline` also times config 1 (the reference's own CPU case, 512 x 512) as BASELINE.md section 3 specifies;
`--config1-full` runs its six-minute full-Q leg instead of the bounded sample.

One JSON line is printed by rank 0.  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` is the same metric through the reference-facing legacy call mmqr(host buffers) with
the H2D/D2H copies inside the timed region.  `--impl reference` times the reference's own CPU
mmqr (oracle/_ref, unmodified qr.c) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version banner there)

METRIC = "QR GFLOP/s (2mn^2-2n^3/3)"


def qr_flops(m: int, n: int) -> float:
    return 2.0 * m * n * n - 2.0 * n ** 3 / 3.0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (unmodified qr.c -> oracle/_ref), 1 thread
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(m: int, n: int, steps: int, warmup: int):
    import oracle
    kind = "reference" if oracle.Ref.available(4, 2) else "port"
    A = oracle.rand_matrix(m, n, 12)
    if kind == "reference":
        ref = oracle.Ref(4, 2)            # qr.c as shipped: PR=4, PC=2 (qr.c:12-13)
        run = lambda: ref.mmqr(A)
    else:
        port = oracle.Port()
        run = lambda: port.mmqr(A, 4, 2)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return kind, dt, qr_flops(m, n) / dt / 1e9


def cpu_config1(full: bool) -> dict:
    """BASELINE config 1 / BASELINE.md section 3: the reference's CPU path on the 512 x 512 srand(12) matrix, 1 core:
    mmqr alone with PR=4/PC=2 (qr.c as shipped) and PR=64/PC=8, and mmqr + explicitQR (full Q and R).  explicitQR is
    O(m^3) per reflector (qr.c:415-429): 512^2 with PR=64/PC=8 takes about six minutes, so the default run times that
    leg on a bounded 304 x 304 sample (PR=64/PC=4, legal: 304 = 64 + 4*60) and `full` runs the named size."""
    import numpy as np
    import oracle
    from oracle import metrics
    out = {"cores": 1, "host_cores": os.cpu_count()}
    A = oracle.rand_matrix(512, 512, 12)
    for PR, PC in ((4, 2), (64, 8)):
        if not oracle.Ref.available(PR, PC):
            continue
        ref = oracle.Ref(PR, PC)
        t0 = time.perf_counter(); rv, tau = ref.mmqr(A); dt = time.perf_counter() - t0
        out[f"mmqr_512x512_pr{PR}_pc{PC}"] = {"seconds": dt, "gflops": qr_flops(512, 512) / dt / 1e9,
                                               "gram_error": metrics.gram_error(A, rv)}
    m, PR, PC = (512, 64, 8) if full else (304, 64, 4)
    if oracle.Ref.available(PR, PC):
        ref = oracle.Ref(PR, PC)
        As = oracle.rand_matrix(m, m, 12)
        t0 = time.perf_counter(); rv, tau = ref.mmqr(As); t1 = time.perf_counter()
        Q, R = ref.explicitQR(rv, tau); t2 = time.perf_counter()
        out[f"mmqr_explicitQR_{m}x{m}_pr{PR}_pc{PC}"] = {
            "mmqr_seconds": t1 - t0, "explicitQR_seconds": t2 - t1, "gflops": qr_flops(m, m) / (t2 - t0) / 1e9,
            "residual_fro": float(np.linalg.norm(Q.astype(np.float64) @ R.astype(np.float64) - As)),
            "backward_over_n_eps": metrics.backward_error(As, Q, R),
            "note": "the named config-1 size" if full else "bounded sample of config 1 (512x512 full Q: ~6 min, see profiles/r02_config1_cpu.json)"}
    return out


def main_reference(args, rank: int):
    if rank != 0:
        return
    m = n = args.ref_size
    kind, dt, gf = cpu_reference_run(m, n, args.steps, args.warmup)
    sample = f"{m}x{n} uniform[0,1) srand(12) matrix (bounded sample of the 16384x16384 workload), mmqr only, PR=4/PC=2 as shipped"
    line = {"impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": "square 16384x16384 fp32 blocked Householder QR (config 2)",
                                                             "sample": sample},
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def main_ours(args, rank: int, world: int, local_rank: int):
    # stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, C stdio of the legacy entry
    # points) write to fd 1, so park fd 1 on stderr for the run and emit the line on the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("cuda-qr_b200")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pkg.Context(local_rank)
    ctx.use_torch_stream()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(step_fn, restore_fn, steps, warmup):
        """W untimed + K timed steps; each step bracketed by CUDA events on the launching stream; inputs are
        restored between steps outside the events (as the reference does, qr.cu:785-787)."""
        for _ in range(warmup):
            restore_fn(); step_fn()
        barrier()
        evs = []
        wall0 = time.perf_counter()
        for _ in range(steps):
            restore_fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step_fn(); e1.record()
            evs.append((e0, e1))
        barrier()
        wall = time.perf_counter() - wall0
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        return max_over_ranks(ms), wall

    out = {}
    # ---------------- square 16384^2 (config 2): the headline workload --------------------------
    m = n = args.size
    g = torch.Generator(device=dev).manual_seed(12 + rank)
    A0 = pkg.colmajor(m, n, device=dev)
    A0.copy_(torch.rand((m, n), device=dev, generator=g))            # uniform[0,1), the reference's distribution
    A = pkg.colmajor(m, n, device=dev)
    tau = torch.zeros(n, device=dev)
    restore = lambda: A.copy_(A0)
    step = lambda: ctx.geqrf(A, tau)
    restore(); step(); torch.cuda.synchronize()                       # sizes the workspace outside any timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms, wall = timed_steps(step, restore, args.steps, args.warmup)
    launches = (ctx.launch_count() - l0) // max(1, args.steps + args.warmup) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    flops = qr_flops(m, n)
    value = world * flops / (ms * 1e-3) / 1e9

    # residual ||A - QR|| / ||A|| of the last timed step (outside the timed region), rank 0
    resid = None
    if rank == 0:
        R = pkg.colmajor(n, n, device=dev)
        ctx.extract_r(A, R)
        QR = pkg.colmajor(m, n, device=dev)
        QR.copy_(R)
        ctx.apply_q(A, tau, QR, trans=False)
        ctx.synchronize()
        num = torch.zeros((), device=dev, dtype=torch.float64)
        den = torch.zeros((), device=dev, dtype=torch.float64)
        for c0 in range(0, n, 2048):                                   # blocked fp64 norms: bounded memory
            d = (A0[:, c0:c0 + 2048].double() - QR[:, c0:c0 + 2048].double())
            num += (d * d).sum(); den += (A0[:, c0:c0 + 2048].double() ** 2).sum()
        resid = float((num / den).sqrt())
        del R, QR

    # per-kernel-class profile of one more step (CUDA events inside the library), rank 0
    roof = None
    if rank == 0:
        restore(); torch.cuda.synchronize()
        ctx.profile_begin(); step(); prof = ctx.profile_end()
        nn = prof["gemm_nn"]
        tot_ms = sum(v["ms"] for v in prof.values())
        ach = nn["flops"] / (nn["ms"] * 1e-3) / 1e12 if nn["ms"] > 0 else 0.0
        tensor_mode = ctx.get_option(pkg.OPT_GEMM) == 1
        peak = pk["bf16_sustained"] / 2.0 / 3.0 if tensor_mode else 74.0
        roof = {"bound": "tensor", "kernel": "umma_gemm_kernel<256,MN> (C -= V X, tcgen05 3xTF32)" if tensor_mode else "gemm_nn_simt_kernel",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_basis": (f"{pk['source']}: bf16 sustained {pk['bf16_sustained']} TF/s / 2 (tf32 rate) / 3 (hi*hi + lo*hi + hi*lo passes)"
                               if tensor_mode else "nominal fp32 SIMT 74 TF/s"),
                "frac_of_bf16_measured": 3.0 * ach / pk["bf16_sustained"] * 2.0 if tensor_mode else None,
                "algorithmic_flops_per_step": nn["flops"], "launches_per_step": nn["launches"],
                "avg_launch_ms": nn["ms"] / max(1, nn["launches"]), "share_of_step": nn["ms"] / tot_ms if tot_ms else None,
                "hbm_GBps_algorithmic": nn["bytes"] / (nn["ms"] * 1e-3) / 1e9 if nn["ms"] > 0 else None,
                "by_class_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                "by_class_launches": {k: v["launches"] for k, v in prof.items()},
                # algorithmic TFLOP/s of every class over its own event brackets (the two streams overlap, so the class
                # times do not add up to the step), and the whole step against the same roof
                "by_class_tflops": {k: round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) for k, v in prof.items() if v["ms"] > 0 and v["flops"] > 0},
                "by_class_flops": {k: v["flops"] for k, v in prof.items() if v["flops"] > 0},
                "gemm_tn_frac": (prof["gemm_tn"]["flops"] / (prof["gemm_tn"]["ms"] * 1e-3) / 1e12 / peak) if prof["gemm_tn"]["ms"] > 0 else None,
                # inside the factorisation the trailing GEMMs own 116 of the 148 SMs (the panel chain holds the other 32):
                # the same achieved rate against the roof of the SMs the kernel actually runs on, for orientation only
                "frac_of_116_sm_partition_roof": ach / (peak * 116.0 / 148.0),
                "step_tflops": flops / (ms * 1e-3) / 1e12, "step_frac": flops / (ms * 1e-3) / 1e12 / peak}
        # measured DRAM traffic of the dominant kernel from the committed ncu --set full capture (one launch at the
        # first-block shape; `achieved` above averages all launches of the step, whose shapes shrink)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["gemm_nn"]
            roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roof["traffic_detail"] = {"shape": tr["shape"], "algorithmic_bytes": tr["algorithmic_bytes"],
                                      "ratio": (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["algorithmic_bytes"],
                                      "source": tr["source"]}
        except Exception:
            pass

    # e2e: legacy mmqr(host buffers) -- H2D, factor, D2H inside the timed region, rank-local, pinned host memory
    e2e = None
    if not args.no_e2e:
        hostA = torch.empty((n, m), dtype=torch.float32, pin_memory=True)  # column-major m x n
        hostA.copy_(A0.t())
        hnp = hostA.numpy().T                                            # F-contiguous view
        htau = __import__("numpy").empty(pkg.tau_size(m, n), dtype="float32")
        pkg.mmqr(hnp, htau)                                              # warm (cudaMalloc pools, page faults)
        ts = []
        for _ in range(args.e2e_steps):
            hostA.copy_(A0.t())
            barrier()
            t0 = time.perf_counter(); pkg.mmqr(hnp, htau); ts.append(time.perf_counter() - t0)
        dt = max_over_ranks(sum(ts) / len(ts))
        e2e = {"value": world * flops / dt / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": m * n * 4,
               "d2h_bytes_per_step": m * n * 4 + n * 4, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "api": "mmqr(float* mat, float* tau, int m, int n) on pinned host buffers (qr.cu:475 signature)"}
        del hostA
    del A, A0
    torch.cuda.empty_cache()

    # ---------------- TSQR 8M x 64 (config 3), row-partitioned + NCCL R-tree --------------------
    if not args.no_extra:
        out["tsqr"] = bench_tsqr(args, pkg, ctx, torch, dist, dev, rank, world, timed_steps, pk)
        out["batched"] = bench_batched(args, pkg, ctx, torch, dev, rank, world, timed_steps, pk)
        if world > 1:
            out["caqr"] = bench_caqr(args, pkg, ctx, torch, dist, dev, rank, world, timed_steps, pk)

    cpu = None
    if rank == 0 and not args.no_cpu:
        kind, dt, gf = cpu_reference_run(args.ref_size, args.ref_size, 1, 0)
        cpu = {"value": gf, "unit": "GFLOP/s", "cores": 1, "kind": kind, "seconds": dt,
               "sample": f"{args.ref_size}x{args.ref_size} srand(12) uniform[0,1) matrix (bounded sample of the 16384x16384 workload: a rate-vs-rate comparison "
                         f"across sizes, not the same input), reference mmqr only (qr.c as shipped, PR=4/PC=2), 1 of {os.cpu_count()} host cores",
               "config1": cpu_config1(args.config1_full)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (tcgen05 3xTF32 split, fp32 accumulate)" if ctx.get_option(pkg.OPT_GEMM) == 1 else "f32",
                "data": "synthetic",
                "config": {"workload": f"square {m}x{n} fp32 blocked Householder QR (BASELINE config 2)" + (" x N replicas" if world > 1 else ""),
                           "panel": 64, "outer_block": ctx.get_option(pkg.OPT_OUTER_BLOCK), "input": "uniform[0,1) Philox seed 12+rank, column-major lda=m",
                           "l2": "input 1 GiB > 126 MB L2; restored from a pristine copy before every step"},
                "residual": resid, "clocks": clocks, "gpu_launches": launches, "wall_s_timed_region": wall,
                "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
        line.update(out)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _traffic(key):
    """measured DRAM bytes of one launch at the bench shape, from the committed ncu --set full capture"""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[key]
        return tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        return None


def bench_tsqr(args, pkg, ctx, torch, dist, dev, rank, world, timed_steps, pk):
    dt = importlib.import_module("cuda-qr_b200.dist_tsqr")
    m_total, n = args.tsqr_rows, 64
    m_loc = m_total // world
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    A_loc = pkg.colmajor(m_loc, n, device=dev)
    A_loc.copy_(torch.rand((m_loc, n), device=dev, generator=g))
    ts = dt.DistTSQR(pkg, ctx, n, rank, world, dev)
    rtree = "single GPU"
    if world > 1:
        if os.environ.get("CQR_BENCH_RTREE", "peer") == "peer":
            ts.enable_peer()                                # cqr_tsqr_dist_r: one tree kernel per rank, NVLink stores into the receiver's slab
            rtree = "peer-memory kernel (cudaIpc slabs, cqr_tsqr_dist_r)"
        else:
            rtree = "ncclSend/ncclRecv + cqr_stack_qr per level (torch.distributed)"
    step = lambda: ts.factor(A_loc, keep_q=False)          # R-only variant: A is read once, never written
    step(); torch.cuda.synchronize()
    ms, _ = timed_steps(step, lambda: None, max(args.steps, 10), args.warmup)
    flops = qr_flops(m_total, n)
    res = {"workload": f"tall-skinny {m_total}x{n} fp32 TSQR, R-only, row-partitioned over {world} GPU(s), binary R-tree (config 3)", "rtree": rtree,
           "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms, "scaling": "strong", "n_gpus": world}
    bytes_alg = 4.0 * m_loc * n
    ach = bytes_alg / (ms * 1e-3) / 1e9
    leaf = ctx.get_option(pkg.OPT_FLAT_TSQR)
    gram = leaf == pkg.TSQR_LEAF_GRAM
    if gram:
        kernel = ("gram_kernel (R = chol(A^T A): error-free bf16 slices on tcgen05 kind::f16, exact fp32 accumulation in TMEM, fp64 "
                  "reduction) + gram_finish_kernel (fp64 Cholesky, condition gate) + the gated Householder leaf's empty launches")
        note = ("per-GPU algorithmic bytes 4*m_loc*n over the whole step (leaf + finish + gate launches + cross-GPU tree); A is read "
                "once by TMA, nothing of size m is written")
    else:
        kernel = "tsqr_flat_r_kernel (warp-resident flat-tree Householder leaf, A read once) + tile_qr_kernel<8> tree"
        note = ("per-GPU algorithmic bytes 4*m_loc*n over the whole step (leaves + tree + NCCL hops); SIMT Householder is FMA-issue bound "
                "at this shape (32 flop/B): the fp32 FMA floor is 2mn^2 / (148 SMs x 128 lanes x 2 x clock) ~ 1.0 ms = 33 % of the HBM roof")
    res["leaf"] = "gram (tcgen05, gated Householder fallback)" if gram else {0: "tile", 1: "flat", 2: "mma", 3: "pair"}.get(leaf, str(leaf))
    res["roofline"] = {"bound": "hbm", "kernel": kernel, "achieved": ach,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                       "traffic": _traffic("tsqr_gram" if gram else "tsqr_flat") if world == 1 else None, "note": note}
    if gram:
        bound, householder = ctx.tsqr_gram_info()
        res["gram_leaf"] = {"cond_bound": bound, "householder_fallback_ran": householder, "bound_max": 32768,
                            "tensor_work": "3 bf16 slices: one M=128 N=192 K=16 tcgen05.mma per 16 rows = 6 m n^2 MAC, 96 clk each: 35-40 % of the tensor pipe at this rate"}
    # the arithmetic rate of the Householder formula against the fp32 FMA ceiling (148 SMs x 128 lanes x 2 x max clock): the
    # Householder leaf is bound by it long before HBM (32 flop/B); the Gram leaf does its contraction on the tensor pipe instead
    fma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    res["roofline"]["fp32_fma"] = {"achieved": flops / world / (ms * 1e-3) / 1e12, "peak": fma_peak, "unit": "TFLOP/s per GPU",
                                   "frac": flops / world / (ms * 1e-3) / 1e12 / fma_peak,
                                   "note": "2mn^2 Householder flops over the step time; not what the Gram leaf executes" if gram else ""}
    # in-run scaling record: this rank's local TSQR alone (leaf + on-GPU tree, no cross-GPU step) and, on rank 0, the
    # whole matrix on one GPU; efficiency = t_1 / (N t_N)
    R_loc = pkg.colmajor(n, n, device=dev)
    ms_local, _ = timed_steps(lambda: ctx.tsqr_r(A_loc, R_loc), lambda: None, 10, 3)
    res["local_ms"] = ms_local
    res["cross_gpu_ms"] = ms - ms_local
    if gram:   # the Householder flat-tree leaf on the same rows, same run (what the gate falls back to)
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_FLAT)
        ms_hh, _ = timed_steps(lambda: ctx.tsqr_r(A_loc, R_loc), lambda: None, 5, 2)
        Rh = torch.triu(R_loc.double())
        ctx.set_option(pkg.OPT_FLAT_TSQR, pkg.TSQR_LEAF_GRAM)
        ctx.tsqr_r(A_loc, R_loc); torch.cuda.synchronize()
        Rg = torch.triu(R_loc.double())
        sg = torch.sign(torch.diagonal(Rh)); sg[sg == 0] = 1
        res["householder_leaf"] = {"local_ms": ms_hh, "speedup": ms_hh / ms_local,
                                   "r_rel_diff_vs_gram_leaf": float((Rh * sg[:, None] - Rg).norm() / Rg.norm())}
    if world > 1:
        t1 = torch.zeros(1, device=dev, dtype=torch.float64)
        if rank == 0:
            A_full = pkg.colmajor(m_total, n, device=dev)
            A_full.copy_(torch.rand((m_total, n), device=dev, generator=g))
            for _ in range(3):
                ctx.tsqr_r(A_full, R_loc)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for e0, e1 in ev:
                e0.record(); ctx.tsqr_r(A_full, R_loc); e1.record()
            torch.cuda.synchronize()
            t1[0] = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            del A_full
        dist.all_reduce(t1)
        res["one_gpu_ms_same_run"] = float(t1.item())
        res["efficiency_vs_ideal"] = float(t1.item()) / (world * ms)
        res["leaf_efficiency"] = float(t1.item()) / (world * ms_local)
    if world == 1:
        # implicit-Q variant and thin-Q expansion (north_star item 4) on the same matrix: A is overwritten by the
        # reflectors, so it is restored from a copy outside the timed region
        A_keep = A_loc.clone()
        Rq = pkg.colmajor(n, n, device=dev)
        Q = pkg.colmajor(m_loc, n, device=dev)
        fq = lambda: ctx.tsqr_factor(A_keep, Rq)
        fq(); ctx.tsqr_form_q(Q); torch.cuda.synchronize()
        ms_f, _ = timed_steps(fq, lambda: A_keep.copy_(A_loc), 5, 2)
        ms_q, _ = timed_steps(lambda: ctx.tsqr_form_q(Q), lambda: None, 5, 2)
        eye = torch.eye(n, device=dev, dtype=torch.float64)
        res["implicit_q"] = {"factor_ms": ms_f, "form_thin_q_ms": ms_q,
                             "orthogonality_over_n_eps": float((Q.t().double() @ Q.double() - eye).norm()) / (n * 2.0 ** -23),
                             "algorithmic_bytes": {"factor": 8.0 * m_loc * n, "form_q": 8.0 * m_loc * n}}
        del A_keep, Q
    # Gram check of the combined R against the distributed A: A^T A = sum over ranks of A_loc^T A_loc
    G = A_loc.t().double() @ A_loc.double()
    if world > 1:
        dist.all_reduce(G)
    if rank == 0:
        Rd = torch.triu(ts.R.double())
        res["gram_error"] = float((Rd.t() @ Rd - G).norm() / G.norm())
    return res


def bench_caqr(args, pkg, ctx, torch, dist, dev, rank, world, timed_steps, pk):
    """BASELINE config 5: row-partitioned CAQR, 16384 x 4096 per rank (131072 x 4096 at 8 GPUs), R-only timing."""
    dc = importlib.import_module("cuda-qr_b200.dist_caqr")
    m_loc, n = args.caqr_rows_per_gpu, args.caqr_cols
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    A0 = pkg.colmajor(m_loc, n, device=dev)
    A0.copy_(torch.rand((m_loc, n), device=dev, generator=g))
    A = pkg.colmajor(m_loc, n, device=dev)
    cq = dc.DistCAQR(pkg, ctx, m_loc, n, rank, world, dev, kb=args.caqr_kb)
    step = lambda: cq.factor(A)
    restore = lambda: A.copy_(A0)
    restore(); step(); torch.cuda.synchronize()
    ms, _ = timed_steps(step, restore, max(min(args.steps, 5), 3), args.warmup)
    m_total = m_loc * world
    flops = qr_flops(m_total, n)
    res = {"workload": f"rectangular {m_total}x{n} fp32 CAQR, {m_loc} rows per GPU over {world} GPU(s): local blocked Householder + "
                       f"tcgen05 trailing update, R / top-row all_gather per {args.caqr_kb}-column block (config 5)",
           "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms, "scaling": "weak", "n_gpus": world,
           "nvlink_bytes_per_rank_per_step": None}
    cq.bytes_exchanged = 0
    restore(); step(); torch.cuda.synchronize()
    res["nvlink_bytes_per_rank_per_step"] = int(cq.bytes_exchanged)
    G = A0.t().double() @ A0.double()
    dist.all_reduce(G)
    R = pkg.colmajor(n, n, device=dev)
    if rank == 0:
        cq.extract_r(A, R)
        Rd = torch.triu(R.double())
        res["gram_error"] = float((Rd.t() @ Rd - G).norm() / G.norm())
    # the three acceptance numbers of north_star on this configuration: thin Q through the local and tree reflectors
    # (DistCAQR.form_q), backward error and orthogonality in fp64 over all ranks' rows, R against the Gram matrix
    Q = pkg.colmajor(m_loc, n, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cq.form_q(A, Q); e1.record(); torch.cuda.synchronize()
    res["form_thin_q_ms"] = e0.elapsed_time(e1)
    dist.broadcast(R.t(), src=0)                              # colmajor(n, n): R.t() is the contiguous storage
    Rt = torch.triu(R).double()
    num = torch.zeros((), device=dev, dtype=torch.float64)
    den = torch.zeros((), device=dev, dtype=torch.float64)
    for r0 in range(0, m_loc, 4096):                          # blocked fp64 products: bounded memory
        d = A0[r0:r0 + 4096].double() - Q[r0:r0 + 4096].double() @ Rt
        num += (d * d).sum(); den += (A0[r0:r0 + 4096].double() ** 2).sum()
    GQ = torch.zeros((n, n), device=dev, dtype=torch.float64)
    for r0 in range(0, m_loc, 4096):
        q = Q[r0:r0 + 4096].double()
        GQ += q.t() @ q
    pack = torch.stack([num, den])
    dist.all_reduce(pack); dist.all_reduce(GQ)
    if rank == 0:
        eps = 2.0 ** -23
        res["backward_error_over_n_eps"] = float((pack[0] / pack[1]).sqrt()) / (n * eps)
        res["orthogonality_over_n_eps"] = float((GQ - torch.eye(n, device=dev, dtype=torch.float64)).norm()) / (n * eps)
        res["residual"] = float((pack[0] / pack[1]).sqrt())
    return res


def bench_batched(args, pkg, ctx, torch, dev, rank, world, timed_steps, pk):
    batch_total, m, n = args.batch, 64, 64
    batch = batch_total // world
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    A0 = torch.rand((batch, n, m), device=dev, generator=g)
    A = torch.empty_like(A0)
    tau = torch.zeros((batch, n), device=dev)
    step = lambda: ctx.geqrf_batched(A, tau)
    restore = lambda: A.copy_(A0)
    restore(); step(); torch.cuda.synchronize()
    ms, _ = timed_steps(step, restore, max(args.steps, 10), args.warmup)
    flops = batch_total * qr_flops(m, n)
    bytes_alg = 2.0 * batch * m * n * 4
    ach = bytes_alg / (ms * 1e-3) / 1e9
    return {"workload": f"batched {batch_total} x ({m}x{n}) fp32 QR, one CTA per matrix, split over {world} GPU(s) (config 4)",
            "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms, "scaling": "strong", "n_gpus": world,
            "roofline": {"bound": "hbm", "kernel": "batched_qr_warp_kernel (one warp per matrix)", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"], "traffic": _traffic("batched_warp") if world == 1 else None,
                         "note": "fp32 SIMT Householder: FMA-issue bound (10.7 flop/B), not HBM bound; see DESIGN.md section 5"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="square workload edge (config 2: 16384)")
    ap.add_argument("--tsqr-rows", type=int, default=8388608)
    ap.add_argument("--caqr-rows-per-gpu", type=int, default=16384)
    ap.add_argument("--caqr-cols", type=int, default=4096)
    ap.add_argument("--caqr-kb", type=int, default=256, help="CAQR outer block width (columns per cross-GPU tree step)")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--ref-size", type=int, default=1536, help="edge of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config1-full", action="store_true", help="time config 1's full-Q leg at 512x512 (about six minutes of CPU)")
    ap.add_argument("--config1-only", action="store_true", help="print only the config-1 CPU timings (no GPU work)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config1_only:
        if rank == 0:
            print(json.dumps({"config1": cpu_config1(args.config1_full)}), flush=True)
        return
    if args.impl == "reference":
        main_reference(args, rank)
        return
    main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
