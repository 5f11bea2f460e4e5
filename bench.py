#!/usr/bin/env python
"""bench.py — LSIGF edge·feature ops/s on B200 (BASELINE.json `metric`), with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload er1m|cfg2|cfg3|cfg4|er2m] [--impl reference]

A "step" is one LSIGF forward (alegnn/utils/graphML.py:83-176 semantics) over one synthetic batch:
  value  = E * nnz * (K-1) * B * G / t_step     (one op = one multiply-add of one non-zero of S with one feature
           column for one hop), inputs resident in HBM in the reference's [B,G,N] layout, timed with
           CUDA events around exactly K steps after a barrier + synchronize, max over ranks.
  e2e    = same metric through the public API (gnn_b200.LSIGF) with pinned HOST x and y: every step's H2D and D2H copy
           inside the timed region, overlapped ACROSS steps on two copy streams (class E2EPipeline).
  roofline = the shift kernel (spmm_hop_kernel): algorithmic bytes per launch (gather model) divided by
           its average duration measured live with CUDA events around every hop launch inside the timed region
           (events recorded by the library on the launching stream), against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline = the reference's dense torch.matmul algorithm (oracle/lsigf_oracle.py:lsigf_dense_torch, a port: the
           reference is Python and cannot travel to the GPU box) on this box's host cores, bounded sample.
Default workload = the configuration the north_star target is quoted on: ER N=1M, avgDeg=32, K=5, G=F=64, B=1, fp32.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: graph, N, deg, E, K, G, F, B, seed
    "er1m": dict(graph="er", N=1_000_000, deg=32, E=1, K=5, G=64, F=64, B=1, seed=1),
    "er2m": dict(graph="er", N=2_000_000, deg=32, E=1, K=5, G=64, F=64, B=1, seed=5),
    "cfg2": dict(graph="er", N=100_000, deg=16, E=1, K=5, G=64, F=64, B=32, seed=2),
    "cfg3": dict(graph="knn", N=1682, deg=10, E=1, K=5, G=64, F=64, B=32, seed=3),
    "cfg4": dict(graph="er", N=200_000, deg=16, E=4, K=3, G=32, F=32, B=32, seed=4),
    "tiny": dict(graph="er", N=20_000, deg=16, E=1, K=5, G=64, F=64, B=1, seed=9),
    # stochastic block model, 1000 communities of 1000 nodes, ~80 % of the edges inside a community (avgDeg ~ 32);
    # nodes are numbered community by community, so gathers have the locality a real graph ordering would give
    "sbm1m": dict(graph="sbm", N=1_000_000, deg=32, E=1, K=5, G=64, F=64, B=1, seed=6, communities=1000, intra=0.8),
}


def describe(w):
    return "%s N=%d avgDeg=%d E=%d K=%d G=%d F=%d B=%d fp32" % (
        {"er": "Erdos-Renyi", "knn": "kNN-like", "sbm": "SBM(%d communities)" % w.get("communities", 0)}[w["graph"]], w["N"], w["deg"], w["E"], w["K"], w["G"], w["F"], w["B"])


def make_gso(w):
    from gnn_b200 import graphs
    if w["graph"] == "er":
        return graphs.er_gso(w["N"], w["deg"], seed=w["seed"], E=w["E"])
    if w["graph"] == "sbm":
        C, n = w["communities"], w["N"] // w["communities"]
        p_in = w["deg"] * w["intra"] / (n - 1)
        p_out = w["deg"] * (1 - w["intra"]) / (w["N"] - n)
        return graphs.sbm_gso(w["N"], C, p_in, p_out, seed=w["seed"], E=w["E"])
    return graphs.knn_like_gso(w["N"], w["deg"], seed=w["seed"])


def hop_algorithmic_bytes(nnz, N, C, s=4):
    """Gather model, per hop and per S_e: col idx + value per nnz, rowptr, one neighbour row of C
    columns per non-zero, one result row per node."""
    return nnz * (4 + s) + (N + 1) * 8 + nnz * C * s + N * C * s


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML every 10 ms; nvidia-smi fallback)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index = index
        self.sm = []
        self.sm_max = None
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)  # all GPUs of the box are visible: NVML index == CUDA index
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        for k, b in bits.items():
            if r & b:
                self.reasons.add(k)

    def _sample_smi(self):
        import subprocess
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [v.strip() for v in out.strip().split(",")]
        if len(f) >= 6:
            self.sm.append(float(f[0]))
            self.sm_max = float(f[1])
            for i, k in enumerate(self.NAMES):
                if f[2 + i].lower().startswith("active"):
                    self.reasons.add(k)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: dense torch port of the reference algorithm on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_dense_sample(w, n_dense, reps):
    """Reference algorithm (dense E x N x N GSO, K-1 batched GEMMs + contraction) at N = n_dense with the workload's
    avgDeg/K/G/F/B.  Returns (ops_per_s, seconds_per_forward, nnz)."""
    import lsigf_oracle as orc
    from gnn_b200 import graphs
    torch.set_num_threads(os.cpu_count() or 1)
    ww = dict(w, N=n_dense)
    gso = graphs.er_gso(n_dense, w["deg"], seed=w["seed"], E=w["E"]) if w["graph"] != "knn" else \
        graphs.knn_like_gso(n_dense, w["deg"], seed=w["seed"])
    S = gso.to_dense().float()
    g = torch.Generator().manual_seed(0)
    bound = 1.0 / np.sqrt(w["G"] * w["K"])
    h = (torch.rand(w["F"], w["E"], w["K"], w["G"], generator=g) * 2 - 1) * bound
    b = (torch.rand(w["F"], 1, generator=g) * 2 - 1) * bound
    x = torch.randn(w["B"], w["G"], n_dense, generator=g)
    with torch.no_grad():
        orc.lsigf_dense_torch(h, S, x, b)  # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.lsigf_dense_torch(h, S, x, b)
            ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    ops = ww["E"] * (gso.nnz() / ww["E"]) * (w["K"] - 1) * w["B"] * w["G"]
    return ops / t, t, gso.nnz()


def cpu_sparse_sample(w, gso, reps=2):
    """The same filter at the workload's FULL size with torch.sparse CSR x dense products on the host cores
    (oracle/lsigf_oracle.py:lsigf_sparse_torch).  NOT reference code — the reference has no sparse path; it shows what a
    CPU could do with the sparse formulation."""
    import warnings
    import lsigf_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    bound = 1.0 / np.sqrt(w["G"] * w["K"])
    h = (torch.rand(w["F"], w["E"], w["K"], w["G"], generator=g) * 2 - 1) * bound
    b = (torch.rand(w["F"], 1, generator=g) * 2 - 1) * bound
    x = torch.randn(w["B"], w["G"], w["N"], generator=g)
    csr = [(r, c, v.astype(np.float32)) for (r, c, v) in gso.csr]
    ts = []
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        prepared = orc.prepare_sparse_torch(csr, w["N"], torch.float32)     # format conversion: once, untimed
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.lsigf_sparse_torch(h, csr, x, b, prepared=prepared)
            ts.append(time.perf_counter() - t0)
    t = float(min(ts))
    ops = float(gso.nnz()) * (w["K"] - 1) * w["B"] * w["G"]
    return {"value": ops / t, "unit": "edge-feature-op/s", "cores": os.cpu_count(), "kind": "port-sparse (not reference code)",
            "sample": "full workload (N=%d, nnz=%d), torch.sparse CSR, best of %d forwards of %.2f s" % (w["N"], gso.nnz(), reps, t)}


def pick_dense_n(w):
    # dense work ~ 2*(K-1)*B*G*E*N^2 flop; keep one forward to a few seconds on a multi-core host
    flop_budget = 6e11
    n = int(np.sqrt(flop_budget / (2.0 * max(w["K"] - 1, 1) * w["B"] * w["G"] * w["E"])))
    return int(min(w["N"], max(1024, min(n, 16384))))


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_dense = pick_dense_n(w)
    t_start = time.time()
    # one "step" = one dense forward on the bounded sample
    ops_s, t, nnz = cpu_dense_sample(w, n_dense, reps=max(1, args.steps))
    line = {
        "impl": "reference", "metric": "LSIGF edge-feature ops/s", "value": ops_s, "unit": "edge-feature-op/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(w), "note": "CPU dense torch.matmul algorithm of the reference (port)"},
        "cpu_baseline": {"value": ops_s, "unit": "edge-feature-op/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "dense GSO at N=%d (nnz=%d), same avgDeg/K/G/F/B as the workload; the dense "
                                   "algorithm is O(N^2) and cannot hold N=%d" % (n_dense, nnz, w["N"])},
        "e2e": {"value": ops_s, "unit": "edge-feature-op/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t_start,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class E2EPipeline:
    """End-to-end steps with the copies overlapped across steps: while step i computes, step i+1's input goes host ->
    device and step i-1's result device -> host on two copy streams (PCIe is full duplex).  Every step still copies ITS
    input from pinned host memory and ITS result back to pinned host memory; device inputs and host outputs are
    double-buffered, ordering is by CUDA events, nothing is skipped or cached."""

    def __init__(self, dev, xh, yh_shape, compute):
        self.xh, self.compute = xh, compute
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.xd = [torch.empty(xh.shape, dtype=xh.dtype, device=dev) for _ in range(2)]
        self.yh = [torch.empty(yh_shape, dtype=xh.dtype).pin_memory() for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def step(self):
        k = self.i & 1
        self.i += 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_cmp[k])            # the compute of two steps ago has finished reading xd[k]
            self.xd[k].copy_(self.xh, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        cur.wait_event(self.ev_in[k])
        y = self.compute(self.xd[k])                         # contiguous device result
        self.ev_cmp[k].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[k])
            self.yh[k].copy_(y, non_blocking=True)
        y.record_stream(self.s_out)


class _JsonOnlyStdout:
    """The driver reads ONE JSON line from stdout; NCCL / libraries print banners there.  Everything written to fd 1
    while this is active goes to stderr; `emit` writes the final line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())


def run_gpu_arm(args, w):
    import torch.distributed as dist
    import gnn_b200
    from gnn_b200 import _cabi
    out_fd = _JsonOnlyStdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("B200GF_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    E, K, G, F, B, N = w["E"], w["K"], w["G"], w["F"], w["B"], w["N"]
    tdt = torch.float64 if args.dtype == "f64" else torch.float32      # the reference's examples run in float64
    es = 8 if args.dtype == "f64" else 4
    gso = make_gso(w).astype(tdt)
    nnz_e = gso.nnz() // E
    ops_per_step = float(gso.nnz()) * (K - 1) * B * G
    g = torch.Generator().manual_seed(0)
    bound = 1.0 / np.sqrt(G * K)
    h = ((torch.rand(F, E, K, G, generator=g) * 2 - 1) * bound).to(dev, tdt)
    b = ((torch.rand(F, 1, generator=g) * 2 - 1) * bound).to(dev, tdt)
    peak, peak_src = measured_peak_gbs()
    out = {}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    if world == 1:
        x = torch.randn(B, G, N, generator=g).to(dev, tdt)      # reference layout, resident in HBM
        plan = gso.plan(dev)
        fwd = lambda: gnn_b200.LSIGF(h, gso, x, b)              # noqa: E731  (layout conversion inside the step)
        hops = E * (K - 1)
        cap = hops * (args.steps + args.warmup)
        lib.b200gf_profile_hops(plan.handle, cap)
        with torch.no_grad(), ClockSampler(local) as clk:
            ms = timed(fwd, args.steps, args.warmup)
        hop_ms = ctypes_floats(lib, plan, cap)[hops * args.warmup:]   # launches inside the timed region only
        lib.b200gf_profile_hops(plan.handle, 0)
        launches_per_step = 1 + 1 + E * (K - 1) + 1             # to_node_major, pack_taps, hops, tap_contract
        # end-to-end through the public API with pinned host buffers
        xh = torch.randn(B, G, N, generator=g).to(tdt).pin_memory()
        pipe = E2EPipeline(dev, xh, (B, F, N), lambda xd: gnn_b200.to_feature_major(gnn_b200.LSIGF(h, gso, xd, b)))
        yh = pipe.yh[0]
        with torch.no_grad():
            ms_e2e = timed(pipe.step, args.steps, 3)
        # forward + backward (reported beside the headline; SURVEY.md §8d asks for both)
        xg = x.clone().requires_grad_(True)
        hg = h.clone().requires_grad_(True)
        bg = b.clone().requires_grad_(True)
        dy = torch.randn(B, F, N, generator=g).to(dev, tdt)

        def fwd_bwd():
            xg.grad = hg.grad = bg.grad = None
            gnn_b200.LSIGF(hg, gso, xg, bg).backward(dy)

        ms_fb = timed(fwd_bwd, max(3, args.steps // 2), 2)
        out["fwd_bwd"] = {"ms_per_step": ms_fb, "unit": "edge-feature-op/s",
                          "value": float(gso.nnz()) * (K - 1) * B * (G + F) / (ms_fb * 1e-3),
                          "note": "forward hops on B*G columns + backward hops on B*F columns per step"}
        C = B * G
        hop_bytes = hop_algorithmic_bytes(nnz_e, N, C, es)
        hop_avg_ms = float(np.mean(hop_ms)) if len(hop_ms) else float("nan")
        achieved = hop_bytes / (hop_avg_ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": load_ncu_traffic(args.workload), "kernel": "spmm_hop_kernel",
                           "bytes_per_launch": hop_bytes, "ms_per_launch": hop_avg_ms, "launches_timed": len(hop_ms),
                           "peak_source": peak_src, "kernel_share_of_step": float(np.sum(hop_ms)) / (ms * args.steps)}
        out["e2e"] = {"value": ops_per_step / (ms_e2e * 1e-3), "unit": "edge-feature-op/s",
                      "h2d_bytes_per_step": xh.numel() * es, "d2h_bytes_per_step": yh.numel() * es, "ms_per_step": ms_e2e}
        out["clocks"] = clk.summary()
        parallelism = "single"
    else:
        from gnn_b200.distributed import PartitionedLSIGF
        part = PartitionedLSIGF(gso, mode=args.mode, device=dev, fused=False if args.no_fused else None, fence=args.fence)
        if args.mode == "nodes":
            x_local = torch.randn(part.rows_per_rank, B * G, generator=torch.Generator().manual_seed(rank)).to(dev, tdt)
        else:
            g0, g1 = part.feature_slice(G)
            x_local = torch.randn(N, B * (g1 - g0), generator=torch.Generator().manual_seed(rank)).to(dev, tdt)
        fwd = lambda: part.forward(h, x_local, b, B=B)          # noqa: E731
        use_graph = args.graph and part.fused and args.fence == "flags" and args.mode == "features"
        if use_graph:
            with torch.no_grad():
                fwd = part.graphed(h, x_local, b, B=B)
        hops = E * (K - 1)
        cap = hops * (args.steps + args.warmup)
        lib.b200gf_profile_hops(part.plan.handle, cap)
        with torch.no_grad(), ClockSampler(local) as clk:
            ms = timed(fwd, args.steps, args.warmup)
        hop_ms = ctypes_floats(lib, part.plan, cap)[hops * args.warmup:]
        lib.b200gf_profile_hops(part.plan.handle, 0)
        # e2e: every rank copies its shard in from pinned host memory and its result rows back
        xh = x_local.cpu().pin_memory()
        if use_graph:
            def compute(xd):
                x_local.copy_(xd)
                return fwd().contiguous()
        else:
            compute = lambda xd: part.forward(h, xd, b, B=B).contiguous()   # noqa: E731
        pipe = E2EPipeline(dev, xh, (part.rows_per_rank, B * F), compute)
        yh = pipe.yh[0]
        with torch.no_grad():
            ms_e2e = timed(pipe.step, args.steps, 3)
        if args.bwd:   # opt-in: forward + the collective backward (dh, db all-reduced; dx sharded like x)
            dy_rows = torch.randn(part.rows_per_rank, B * F, generator=torch.Generator().manual_seed(100 + rank)).to(dev, tdt)

            def fwd_bwd():
                part.forward(h, x_local, b, B=B)
                part.backward(h, x_local, dy_rows, B=B, want_db=True)

            with torch.no_grad():
                ms_fb = timed(fwd_bwd, max(3, args.steps // 2), 2)
            out["fwd_bwd"] = {"ms_per_step": ms_fb, "unit": "edge-feature-op/s",
                              "value": float(gso.nnz()) * (K - 1) * B * (G + F) / (ms_fb * 1e-3),
                              "note": "partitioned forward + backward (NCCL exchanges in the backward)"}
        # my kernels per rank and step: pack_taps, split_w, tc_contract, the hops, and the scatter of the k = 0 slice
        launches_per_step = (E * (K - 1) + 3 + (1 if part.fused else 0)) * world
        if args.mode == "nodes":
            c_loc, nnz_loc, rows_loc = B * G, part.local_nnz // E, part.rows_per_rank
        else:
            g0, g1 = part.feature_slice(G)
            c_loc, nnz_loc, rows_loc = B * (g1 - g0), nnz_e, N
        hop_bytes = hop_algorithmic_bytes(nnz_loc, rows_loc, c_loc, es)
        hop_avg_ms = float(np.mean(hop_ms)) if len(hop_ms) else float("nan")
        achieved = hop_bytes / (hop_avg_ms * 1e-3) / 1e9
        out["roofline"] = None if not hop_ms else {
                           "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": None, "kernel": "spmm_hop_kernel (rank 0 shard: %d rows x %d columns, %d nnz)" %
                                                        (rows_loc, c_loc, nnz_loc),
                           "bytes_per_launch": hop_bytes, "ms_per_launch": hop_avg_ms, "launches_timed": len(hop_ms),
                           "peak_source": peak_src, "kernel_share_of_step": float(np.sum(hop_ms)) / (ms * args.steps)}
        out["e2e"] = {"value": ops_per_step / (ms_e2e * 1e-3), "unit": "edge-feature-op/s",
                      "h2d_bytes_per_step": xh.numel() * es * world, "d2h_bytes_per_step": yh.numel() * es * world,
                      "ms_per_step": ms_e2e}
        out["clocks"] = clk.summary()
        parallelism = "%s-partition x%d%s%s" % (args.mode, world,
                                                " (fused hop+NVLink scatter, %s fence)" % args.fence if part.fused else "",
                                                ", CUDA graph" if use_graph else "")

    if rank == 0:
        line = {
            "metric": "LSIGF edge-feature ops/s", "value": ops_per_step / (ms * 1e-3), "unit": "edge-feature-op/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": describe(w).replace("fp32", "fp64" if args.dtype == "f64" else "fp32"), "name": args.workload, "nnz": gso.nnz(), "parallelism": parallelism,
                       "l2": "inputs larger than L2 (x and every z_k are %d MB each; no flush needed)" %
                             (N * B * G * es // 2 ** 20) if N * B * G * es > 126 * 2 ** 20 else
                             "working set fits L2: numbers are L2-warm",
                       "ops_per_step": ops_per_step},
            "gpu_launches": launches_per_step * args.steps,
        }
        line.update(out)
        if world == 1 and not args.no_cpu_baseline:
            n_dense = pick_dense_n(w)
            ops_s, t, nnz_d = cpu_dense_sample(w, n_dense, reps=3)
            line["cpu_baseline"] = {"value": ops_s, "unit": "edge-feature-op/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "reference dense torch.matmul algorithm at N=%d (nnz=%d), 3 forwards of "
                                              "%.2f s; same avgDeg/K/G/F/B" % (n_dense, nnz_d, t)}
            if w["B"] * w["G"] <= 256:   # second CPU line at the FULL graph size: sparse torch restatement, not reference code
                try:
                    line["cpu_sparse_baseline"] = cpu_sparse_sample(w, gso)
                except Exception as exc:  # never let the extra baseline break the bench line
                    line["cpu_sparse_baseline"] = {"error": str(exc)[:200]}
        out_fd.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ctypes_floats(lib, plan, n):
    import ctypes
    buf = (ctypes.c_float * n)()
    got = lib.b200gf_profile_read(plan.handle, buf, n)
    return [float(buf[i]) for i in range(max(0, got))]


def load_ncu_traffic(workload):
    """dram bytes per hop launch from the committed ncu capture of this workload (profiles/), or null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="er1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="features", choices=["nodes", "features"],
                    help="multi-GPU sharding (DESIGN.md §4): feature columns (default) or node rows")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="arithmetic type (headline: f32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fence", default="flags", choices=["flags", "nccl"],
                    help="multi-GPU fused path: peer flags in symmetric memory (default) or a 4-byte NCCL all-reduce")
    ap.add_argument("--graph", action="store_true", help="multi-GPU fused path: replay the step as a CUDA graph")
    ap.add_argument("--no-fused", action="store_true", help="multi-GPU: NCCL all-to-all instead of the fused NVLink scatter")
    ap.add_argument("--bwd", action="store_true", help="multi-GPU: also time forward + partitioned backward (fwd_bwd key)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
    else:
        run_gpu_arm(args, w)


if __name__ == "__main__":
    main()
