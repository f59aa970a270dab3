#!/usr/bin/env python
"""bench.py — GGML_OP_MUL_MAT over block-quantized weights on B200 (contract: see the task statement / DESIGN.md §4).

Headline workload (BASELINE.json configs[1]): Q4_K 4096 -> 11008 mat-vec, n_batch = 1 (Llama-7B FFN shape).
  step            one sweep over NBUF distinct weight matrices (NBUF x 25.4 MB > 2 x the 126 MB L2, so every
                  mat-vec streams its weights from HBM), i.e. NBUF mat-vecs
  value           GB/s of quantized weight bytes processed, whole job (all ranks), inputs resident in HBM,
                  CUDA-event timed on the launching stream, max over ranks
  e2e             the same metric through the C ABI with HOST buffers (ggml_b200_mul_mat_host): every mat-vec
                  copies the activation vector from pinned host memory and the result back, synchronously
  roofline        algorithmic bytes (weights + x + y) / mean device time per launch, vs the measured HBM peak
  cpu_baseline    the unmodified reference's ggml-cpu backend (oracle/_ref) on this box's host cores, same shape
  --impl reference  the reference arm: ggml-cpu via its own public API, same metric/config
N > 1: weak scaling — every rank owns one row-shard (M rows) of an (N*M) x K matrix and the output slices are
exchanged with an NCCL all-gather (the path's only exchange step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WL = {"type": "q4_K", "type_id": 12, "K": 4096, "M": 11008, "N": 1}
METRIC = "mul_mat_q4_K_4096x11008_n1_weight_throughput"
UNIT = "GB/s"


def weight_bytes(K, M):
    return (K // 256) * 144 * M


def algorithmic_bytes(K, M, N):
    return weight_bytes(K, M) + K * N * 4 + M * N * 4


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


NBUF = 13          # distinct weight matrices per step: 13 x 25.4 MB = 330 MB > 2 x the 126 MB L2 (and > the host's last-level cache)


def make_weights(torch, nbuf, K, M, seed):
    """NBUF distinct packed Q4_K matrices: arbitrary code bytes, finite small fp16 d/dmin per superblock."""
    rb = (K // 256) * 144
    gen = torch.Generator(device="cuda").manual_seed(seed)
    out = []
    for _ in range(nbuf):
        w = torch.randint(0, 256, (M * (K // 256), 144), dtype=torch.uint8, device="cuda", generator=gen)
        d = (torch.rand((M * (K // 256), 2), device="cuda", generator=gen) * 2e-3 + 1e-4).to(torch.float16).view(torch.uint8)
        w[:, 0:4] = d.reshape(-1, 4)
        out.append(w.reshape(-1))
    assert out[0].numel() == rb * M
    return out


def cpu_baseline(sample_s=12.0):
    """The reference's ggml-cpu MUL_MAT on the host cores (oracle/_ref), bounded sample of the same workload."""
    from oracle import oracle as O
    ref = O.Ref()
    K, M, N, t = WL["K"], WL["M"], WL["N"], WL["type_id"]
    rng = np.random.default_rng(1234)
    W = O.random_blocks(t, M * K // 256, rng)
    X = np.random.default_rng(5678).uniform(-1, 1, K * N).astype(np.float32)
    best = None
    hw = ref.hw_threads()
    nw = NBUF                                   # same sweep as the GPU arm: 13 distinct matrices (330 MB), larger than the host's LLC
    for threads in sorted({hw, max(1, hw // 2), max(1, hw // 4)}, reverse=True):
        _, s = ref.mul_mat_sweep(t, W, X, M, N, K, nw, threads=threads, iters=2, warmup=1)          # probe
        iters = max(1, int(sample_s / 3 / max(s * nw, 1e-6)))
        _, s = ref.mul_mat_sweep(t, W, X, M, N, K, nw, threads=threads, iters=iters, warmup=1)
        if best is None or s < best[0]:
            best = (s, threads, iters * nw)
    s, threads, n = best
    return {"value": weight_bytes(K, M) / s / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{n} MUL_MAT nodes (q4_K {K}x{M}, n=1; {nw} distinct weight tensors per graph = 330 MB sweep, persistent threadpool), "
                      f"ggml-cpu {'native' if ref.native else 'x86-64-v3'} build, {s * 1e6:.1f} us per mat-vec",
            "us_per_matvec": s * 1e6}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, M, N = WL["K"], WL["M"], WL["N"]
    from oracle import oracle as O
    ref = O.Ref()
    t = WL["type_id"]
    rng = np.random.default_rng(1234)
    W = O.random_blocks(t, M * K // 256, rng)
    X = np.random.default_rng(5678).uniform(-1, 1, K * N).astype(np.float32)
    # the reference's threadpool is used with the thread count that serves it best on this box (all hardware threads, or one per
    # physical core: oversubscribing the SMT siblings of a 2-socket host slows ggml-cpu's spin barriers down by several x)
    hw = ref.hw_threads()
    best_s, threads = None, hw
    nbuf = NBUF
    for th in sorted({hw, max(1, hw // 2), max(1, hw // 4)}, reverse=True):
        _, sp = ref.mul_mat_sweep(t, W, X, M, N, K, nbuf, threads=th, iters=3, warmup=1)
        if best_s is None or sp < best_s:
            best_s, threads = sp, th
    # one step = one sweep of NBUF mat-vecs over NBUF distinct weight tensors (330 MB > the host's last-level cache), like the GPU arm;
    # fewer timed steps if the requested K would take > ~90 s
    per_step = nbuf
    steps = args.steps
    if best_s * nbuf * (steps + args.warmup) > 90.0:
        steps = max(1, int(90.0 / (best_s * nbuf)) - args.warmup)
    _, s = ref.mul_mat_sweep(t, W, X, M, N, K, nbuf, threads=threads, iters=steps, warmup=args.warmup, e2e=True)
    val = weight_bytes(K, M) / s / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": s * per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 x int4 (q4_K x q8_K), f32 accumulate",
            "data": "synthetic", "config": {"workload": f"q4_K {K}x{M} mat-vec n_batch=1 (BASELINE.json configs[1])", "mat_vecs_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{per_step} mat-vecs over {per_step} distinct weight tensors per step through ggml_backend_graph_compute on ggml-cpu ({'native' if ref.native else 'x86-64-v3'} build), tensor_set/get included"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of CUDA-graph replay")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: 'fused' = the mat-vec kernel stores its rows into every peer's y over NVLink and signals flags; "
                         "'nccl' = ncclAllGather of the slices after the kernel (the baseline)")
    ap.add_argument("--cpu-baseline-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_child:
        print(json.dumps(cpu_baseline()), flush=True)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import ggml_b200 as g
    g.lib()                                    # fail loudly if the CUDA library is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, M, N, t = WL["K"], WL["M"], WL["N"], WL["type_id"]
    wb = weight_bytes(K, M)
    nbuf = NBUF                                                     # 13 x 25.4 MB = 330 MB > 2 x L2
    Ws = make_weights(torch, nbuf, K, M, seed=1234 + rank)
    X = torch.rand(N * K, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5678)) * 2 - 1
    Yloc = torch.empty((1, 1, N, M), dtype=torch.float32, device="cuda")
    Yall = torch.empty((world, N * M), dtype=torch.float32, device="cuda") if world > 1 else None
    assert g.mul_mat_plan(t, M, N, K) == g.MM_GEMV, "headline workload must run on the TMA mat-vec kernel"

    fused = world > 1 and args.exchange == "fused"
    use_graph = (world == 1 or fused) and not args.no_graph
    ex = None
    # Like the reference's own perf harness (tests/test-backend-ops.cpp eval_perf, :657-660: the op node repeated in one
    # graph on the same inputs, one output tensor per node) the mat-vecs of a sweep are INDEPENDENT: fixed activation vector,
    # one output buffer per weight matrix.  Weights and activations are final before the timed region, so the launches are
    # flagged SRC0_STATIC | SRC1_STATIC and consecutive launches may overlap (programmatic dependent launch).
    # `dependent_chain` re-times the sweep with every launch waiting for the previous one (one shared output, activations
    # treated as produced by the preceding kernel): the latency-bound lower bracket.
    Ys = [torch.empty((1, 1, N, M), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
    F_IND = g.MM_SRC0_STATIC | g.MM_SRC1_STATIC
    F_DEP = g.MM_SRC0_STATIC
    if fused:
        # same independence rules as N = 1: every op of the sweep has its own gathered-y slot on every rank; the peer stores
        # and flag publication of op i overlap the streaming of op i+1; one stream-wait at the end of the sweep
        ex = g.PeerExchange(world * M, rank, world, rank * M, slots=nbuf)
        margs = [g.mul_mat_args(t, Ws[i], X, Yloc, M, N, K, flags=F_IND) for i in range(nbuf)]

    def sweep(dependent=False):
        for i in range(nbuf):
            if fused:
                ex.mul_mat_gather(margs[i], slot=i)                                    # compute + NVLink peer stores + flag publish
                if i == nbuf - 1:
                    ex.wait()                                                          # every rank's slices of the whole sweep have landed
            elif world > 1:
                g.mul_mat(t, Ws[i], X, M, N, K, out=Yloc, flags=F_DEP)
                dist.all_gather_into_tensor(Yall, Yloc.view(-1))
            elif dependent:
                g.mul_mat(t, Ws[i], X, M, N, K, out=Yloc, flags=F_DEP)
            else:
                g.mul_mat(t, Ws[i], X, M, N, K, out=Ys[i], flags=F_IND)

    sweep()                                                          # first-launch setup outside capture
    sweep(True)
    torch.cuda.synchronize()
    graph = None
    launches_per_step = None
    if use_graph:
        c0 = g.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sweep()
        launches_per_step = g.launch_count() - c0
        step = graph.replay
    else:
        step = sweep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    c0 = g.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launched = (g.launch_count() - c0) if not use_graph else launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    n_mv = nbuf * args.steps
    value = world * n_mv * wb / (ms * 1e-3) / 1e9
    us_per_launch = ms * 1e3 / n_mv
    dep_us = None
    if world == 1:
        if use_graph:
            gdep = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gdep):
                sweep(True)
            dstep = gdep.replay
        else:
            dstep = lambda: sweep(True)
        for _ in range(args.warmup):
            dstep()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            dstep()
        d1.record()
        torch.cuda.synchronize()
        dep_us = d0.elapsed_time(d1) * 1e3 / n_mv

    # ---- e2e: host buffers through the C ABI (rank-local; aggregated like `value`).  One step = one call of
    # ggml_b200_mul_mat_host_batch: upload the step's input x from pinned host memory, the sweep's NBUF mat-vecs, download the
    # NBUF results into pinned host memory, synchronise.  (The per-op variant ggml_b200_mul_mat_host is timed too.)
    import ctypes as C
    L = g.lib()
    L.ggml_b200_mul_mat_host_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    xh = torch.empty(N * K, dtype=torch.float32).pin_memory()
    xh.copy_(X.cpu())
    yh = torch.empty((nbuf, N * M), dtype=torch.float32).pin_memory()
    arr = (g.MulMatArgs * nbuf)()
    for i in range(nbuf):
        ai = g.mul_mat_args(t, Ws[i], X, Ys[i], M, N, K, flags=F_IND)
        C.memmove(C.addressof(arr[i]), C.addressof(ai), C.sizeof(g.MulMatArgs))
    hdst = (C.c_void_p * nbuf)(*[yh[i].data_ptr() for i in range(nbuf)])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def e2e_step():
        g.check(L.ggml_b200_mul_mat_host_batch(arr, nbuf, C.c_void_p(xh.data_ptr()), hdst, stream), "ggml_b200_mul_mat_host_batch")

    for _ in range(5):
        e2e_step()
    barrier()
    e2e_steps = max(20, min(args.steps, 300))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    a1 = g.mul_mat_args(t, Ws[0], X, Yloc, M, N, K, flags=g.MM_SRC0_STATIC)
    for i in range(10):
        g.check(L.ggml_b200_mul_mat_host(C.byref(a1), C.c_void_p(xh.data_ptr()), C.c_void_p(yh[0].data_ptr()), stream), "ggml_b200_mul_mat_host")
    t1 = time.perf_counter()
    for i in range(200):
        a1.src0 = Ws[i % nbuf].data_ptr()
        g.check(L.ggml_b200_mul_mat_host(C.byref(a1), C.c_void_p(xh.data_ptr()), C.c_void_p(yh[0].data_ptr()), stream), "ggml_b200_mul_mat_host")
    per_op_us = (time.perf_counter() - t1) / 200 * 1e6
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_n = e2e_steps * nbuf
    e2e_val = world * e2e_n * wb / e2e_s / 1e9

    if rank != 0:
        if world > 1:
            dist.barrier()
            if ex is not None:
                ex.close()
            dist.destroy_process_group()
        return

    if world > 1:
        dist.barrier()
    peak, peak_src = measured_peaks()
    achieved = algorithmic_bytes(K, M, N) / (us_per_launch * 1e-6) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "r01_gemv_q4k_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8 x int4 (q4_K weights x q8_K activations, dp4a), f32 accumulate", "data": "synthetic",
        "config": {"workload": f"q4_K {K}x{M} mat-vec n_batch=1 (BASELINE.json configs[1])", "mat_vecs_per_step": nbuf,
                   "l2_policy": f"inputs larger than L2: {nbuf} distinct weight matrices ({nbuf * wb / 1e6:.0f} MB) visited round-robin",
                   "launch": "CUDA-graph replay of one sweep" if use_graph else "plain stream launches",
                   "independence": "mat-vecs of a sweep are independent ops (fixed x, one y per weight matrix) as in the reference's eval_perf; "
                                   "consecutive launches may overlap via programmatic dependent launch",
                   "dependent_chain": None if dep_us is None else {"us_per_matvec": dep_us, "GBps": wb / dep_us / 1e3,
                                                                   "note": "every launch waits for the previous kernel (x treated as its output, shared y)"},
                   "parallelism": (f"row-shard x{world}, exchange fused into the mat-vec kernel (NVLink peer stores + flags)" if fused else
                                   f"row-shard x{world} + NCCL all-gather of output slices") if world > 1 else "single GPU"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": K * N * 4, "d2h_bytes_per_step": M * N * 4 * nbuf,
                "us_per_step": e2e_s / e2e_steps * 1e6, "us_per_matvec": e2e_s / e2e_n * 1e6,
                "api": "ggml_b200_mul_mat_host_batch: per step, pinned host x -> device, the sweep's 13 mat-vecs, 13 results -> pinned host, stream sync",
                "per_op_call_us": per_op_us, "per_op_api": "ggml_b200_mul_mat_host (upload, one mat-vec, download, sync per op)"},
        "gpu_launches": int(launched),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "mmvq_sb_kernel<Q4_K>", "us_per_launch": us_per_launch,
                     "algorithmic_bytes_per_launch": algorithmic_bytes(K, M, N), "peak_source": peak_src},
    }
    if world == 1 and not args.no_cpu_baseline:
        # the baseline is a report, never a reason to lose the GPU line: run the reference in a child with a deadline
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-baseline-child"], capture_output=True, text=True, timeout=150)
            line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        if ex is not None:
            ex.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
