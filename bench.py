#!/usr/bin/env python
"""bench.py -- GNN forward (gnn::model::predict) throughput in undirected edges/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode exact|fast] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): R-MAT, Graph500
parameters (0.57, 0.19, 0.19, 0.05), scale 20, edge factor 16, symmetrised and
de-duplicated, integer weights 1..200, trained GNN_VC weights (tests/golden).  For
N > 1 GPUs the scale grows by log2(N) (weak scaling: fixed work per GPU; N = 8 is
scale 23, 8.4 M vertices / ~128 M edges, BASELINE configs[3]); the graph is cut
into work-balanced vertex ranges and the 16-float rows are exchanged between
stages over NCCL.  A "step" is one whole forward: x -> scores.

One JSON line on stdout (rank 0).  `value` is device-resident throughput, `e2e`
goes through the host-buffer C ABI call (pinned H2D of x, D2H of the scores).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gnn_forward_edges_per_sec"
UNIT = "edges/s"


def rmat_scale_for(n_gpus: int) -> int:
    return 20 + max(0, int(round(np.log2(n_gpus))))


def algorithmic_bytes(n: int, m: int):
    """fp32 rows, uint32 ids/offsets, no-reuse gather model, per stage."""
    s0 = 8 * m + 80 * n
    s1 = 68 * m + 140 * n
    s2 = 68 * m + 80 * n
    return s0, s1, s2


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region, via NVML (5 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), or the oracle port
# ------------------------------------------------------------------------------------------
def reference_workload(args):
    """The graph the reference arm times: the arm's own config when one CPU step of it stays within
    seconds (R-MAT scale <= 20, ~4 s per predict on 16 cores), else a bounded sample of the same
    generator (scale 20 stands in for the scale-21..23 graphs of the 2/4/8-GPU runs)."""
    import torch
    from gnn_mwvc_b200 import graphs
    want = args.scale if args.scale else rmat_scale_for(args.gpus)
    scale = min(want, 20)
    # generated where the GPU arm generates it (torch's CUDA and CPU generators draw different
    # streams): on the GPU box both arms time the very same graph
    g = graphs.rmat_graph(scale, 16, seed=42, device="cuda" if torch.cuda.is_available() else "cpu")
    name = f"rmat_scale{want}_ef16"
    sample = None if scale == want else f"R-MAT scale {scale} ef 16 (n={g.n}, E={g.n_edges}) stands in for scale {want}"
    return g, name, sample


def blas_kernel(ref) -> str:
    """'OpenBLAS 0.3.15 ... Haswell MAX_THREADS=128' -> the kernel set OpenBLAS picked for this CPU."""
    cfg = ref.blas_config().split()
    known = [w for w in cfg if w[0].isupper() and w.isalnum() and w not in ("OpenBLAS", "DYNAMIC_ARCH", "NO_AFFINITY")]
    return known[-1] if known else "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import gnn_mwvc_b200  # noqa: F401
    from gnn_mwvc_b200 import capi
    from oracle import pyoracle as po

    cores = os.cpu_count() or 1
    layers = capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz")
    text = po.layers_to_text(layers)
    g, wl_name, sample_note = reference_workload(args)
    rp, col, W, NW = g.numpy()
    eu, ev = g.edges_numpy()
    scale = 200.0
    x = W.astype(np.float32) / np.float32(scale)
    if po.REF_SO.exists():
        # the TIMED reference runs the kernel set OpenBLAS itself selects on this host (the pin to
        # "Prescott" is the checker's, see oracle/pyoracle.py), with every host thread
        ref = po.Reference(threads=cores, coretype=None)
        h = ref.model(text)
        gh = ref.graph_create(g.n, eu, ev, W)          # resident, as GNN_VC holds its graph
        kind, kernel = "reference", blas_kernel(ref)

        def step():
            ref.predict_on(h, gh, x, scale)
            return ref.last_seconds                    # predict() only
    else:
        orc = po.Oracle()
        h = orc.parse(text)
        kind, cores, kernel = "port", 1, "plain C restatement"

        def step():
            t = time.perf_counter()
            orc.predict(h, rp, col, W, NW, x, scale)
            return time.perf_counter() - t
    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = float(np.sum(secs))
    value = g.n_edges * args.steps / total
    sample = (sample_note or f"the whole workload graph (n={g.n}, E={g.n_edges})") + f", predict() only, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "vertices": g.n, "edges": g.n_edges, "nnz": g.nnz, "sample": sample,
                   "mode": f"reference CPU (OpenBLAS kernel {kernel}, {cores} threads)" if kind == "reference" else "oracle port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "blas_kernel": kernel},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def cpu_baseline_leg(layers, g_cpu, budget_s: float = 25.0):
    """Rank 0, N=1: the reference CPU path on a bounded sample of the SAME graph the GPU arm timed
    (a few predict() calls on it, ~10-30 s of CPU work), OpenBLAS' own kernel choice, all host threads.
    Runs in a child process: OpenBLAS fixes its kernel set when it is first loaded."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.savez(Path(td) / "g.npz", eu=g_cpu["eu"], ev=g_cpu["ev"], w=g_cpu["w"], n=g_cpu["n"])
        code = f"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, {str(ROOT)!r})
import gnn_mwvc_b200
from gnn_mwvc_b200 import capi
from oracle import pyoracle as po
z = np.load({str(Path(td) / 'g.npz')!r})
n, eu, ev, W = int(z['n']), z['eu'], z['ev'], z['w']
layers = capi.load_model_npz({str(ROOT / 'tests' / 'golden' / 'mwvc_model.npz')!r})
text = po.layers_to_text(layers)
x = W.astype(np.float32) / np.float32(200.0)
cores = os.cpu_count() or 1
secs, t_all = [], time.perf_counter()
if po.REF_SO.exists():
    ref = po.Reference(threads=cores, coretype=None)
    h = ref.model(text); gh = ref.graph_create(n, eu, ev, W)
    kind = 'reference'; cfg = ref.blas_config()
    while len(secs) < 2 or (time.perf_counter() - t_all < {budget_s} and len(secs) < 6):
        ref.predict_on(h, gh, x, 200.0); secs.append(ref.last_seconds)
else:
    from gnn_mwvc_b200 import graphs
    import torch
    g = graphs.graph_from_edges(n, torch.from_numpy(eu.astype(np.int64)), torch.from_numpy(ev.astype(np.int64)), torch.from_numpy(W.astype(np.int64)))
    rp, col, W2, NW = g.numpy()
    orc = po.Oracle(); h = orc.parse(text); kind, cores, cfg = 'port', 1, 'plain C restatement'
    while len(secs) < 2 or (time.perf_counter() - t_all < {budget_s} and len(secs) < 4):
        t = time.perf_counter(); orc.predict(h, rp, col, W2, NW, x, 200.0); secs.append(time.perf_counter() - t)
print(json.dumps(dict(secs=secs, kind=kind, cores=cores, cfg=cfg)))
"""
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "failed", "sample": r.stderr[-300:]}
    d = json.loads(r.stdout.strip().splitlines()[-1])
    secs = d["secs"]
    best = float(np.median(secs[1:])) if len(secs) > 1 else secs[0]
    e = int(len(g_cpu["eu"]))
    cfgw = d["cfg"].split()
    kernel = next((w for w in reversed(cfgw) if w[0].isupper() and w.isalnum() and w not in ("OpenBLAS", "DYNAMIC_ARCH", "NO_AFFINITY")), d["cfg"])
    return {"value": e / best, "unit": UNIT, "cores": d["cores"], "kind": d["kind"], "blas_kernel": kernel,
            "sample": f"{'R-MAT scale 20 sample of the generator' if g_cpu.get('sample') else 'the bench graph itself'} (n={g_cpu['n']}, E={e}): median of {max(len(secs) - 1, 1)} warm predict() calls "
                      f"after one cold call, {best * 1e3:.0f} ms each"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import gnn_mwvc_b200 as pkg
    from gnn_mwvc_b200 import capi, graphs
    from gnn_mwvc_b200 import dist as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG", "WARN")   # a caller's setting wins (the driver reads the rank count from INFO)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libgvc has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    mode = pkg.MODE_EXACT if args.mode == "exact" else pkg.MODE_FAST
    layers = capi.load_model_npz(ROOT / "tests" / "golden" / "mwvc_model.npz")
    scale_log2 = args.scale if args.scale else rmat_scale_for(world)

    # ---- synthetic graph, generated on the GPU (every rank builds the same graph) ----------
    t0 = time.perf_counter()
    if args.workload == "rmat":
        g = graphs.rmat_graph(scale_log2, 16, seed=42, device=dev)
        wl_name = f"rmat_scale{scale_log2}_ef16"
    elif args.workload == "grid":
        side = int(round((20_000_000 * world) ** 0.5))
        g = graphs.grid_graph(side, side, device=dev)
        wl_name = f"grid_{side}x{side}"
    elif args.workload == "isolated":      # dense chain only: no edges at all (diagnostic)
        nn = (1 << 20) * world
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        g = graphs.graph_from_edges(nn, z, z, graphs.random_weights(nn, 3, dev), name="isolated")
        wl_name = f"isolated_{nn}"
    else:
        g = graphs.er_graph(10000 * world, 50000 * world, seed=1, device=dev)
        wl_name = f"er_{g.n}_{g.n_edges}"
    n, m, e_total = g.n, g.nnz, g.n_edges
    weight_scale = 200.0
    x_full = (g.weights.to(torch.float32) / weight_scale).contiguous()
    if world > 1:
        # equal-sized, equal-work shards: vertices dealt to the shards in order of descending degree
        g, _perm = graphs.balanced_relabel(g, world)
        x_full = (g.weights.to(torch.float32) / weight_scale).contiguous()
        per = g.n // world
        bounds = [r * per for r in range(world + 1)]
    else:
        bounds = [0, n]
    shard = gdist.make_shard(g, bounds, rank, skip_isolated=world > 1)
    rp32 = shard.row_ptr.to(torch.int32).contiguous()
    # host copy of the edge list: predict() end to end (the drop-in builds a reduction_graph from it)
    # and the CPU baseline leg work on the very graph the GPU arm times, when it is small enough
    g_cpu = None
    if world == 1 and g.eu is not None and e_total <= 40_000_000:
        g_cpu = {"n": n, "eu": g.eu.cpu().numpy().astype(np.uint32), "ev": g.ev.cpu().numpy().astype(np.uint32),
                 "w": g.weights.cpu().numpy().view(np.uint32).copy()}
    g.eu = g.ev = None
    gen_s = time.perf_counter() - t0

    ctx = pkg.Context(local_rank)
    ctx.model_upload(layers)
    ctx.graph_adopt(rp32, shard.col, shard.weights, shard.nw, n_global=g.n, v_begin=shard.v_begin, v_end=shard.v_end)
    if world > 1:
        ctx.graph_set_tail(int(_perm[n - 1].item()) if n % 2 else None)
    stream = ctx.torch_stream()
    # N > 1: the stage kernels store their rows straight into the other ranks' h1/h2 over NVLink
    # (peer memory); --exchange nccl all-gathers them between the stages instead
    pr, pr_note = None, None
    if world > 1 and args.exchange == "peer":
        try:
            pr = gdist.PeerRows(ctx, g.n, bounds=bounds)
        except RuntimeError as e:          # raised on every rank alike: fall back to the collective, and say so
            pr_note = str(e)
            print(f"bench: {pr_note}; exchanging rows with NCCL all-gathers instead", file=sys.stderr, flush=True)
    h1 = pr.h1 if pr else torch.zeros(g.n, 16, device=dev)
    h2 = pr.h2 if pr else torch.zeros(g.n, 16, device=dev)
    scores = torch.zeros(shard.n_local, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    torch.cuda.synchronize()

    def forward_once():
        if world == 1:
            ctx.forward_device(x_full, weight_scale, scores, mode)
        else:
            gdist.sharded_forward(ctx.stage_device, shard, x_full, h1, h2, scores, weight_scale, mode, peer_rows=pr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            forward_once()
        barrier()
        # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        launches0 = ctx.launches
        with ClockSampler(local_rank) as clk:
            barrier()
            for i in range(args.steps):
                flush.fill_(i & 0xFF)
                starts[i].record(stream)
                forward_once()
                ends[i].record(stream)
            barrier()
        launches = ctx.launches - launches0
        step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        total_ms = float(np.sum(step_ms))

        # ---- multi-GPU: where the step goes (stage kernels vs row exchanges), CUDA events --------
        phase_ms = None
        if world > 1:
            names = ["stage0", "exchange_h1", "stage1", "exchange_h2", "stage2"]
            acc = {k: [] for k in names}
            for rep in range(5):
                flush.fill_(rep)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                ev[0].record(stream)
                ctx.stage_device(0, x_full, h1, weight_scale, mode); ev[1].record(stream)
                (pr.barrier() if pr else gdist.exchange_rows(h1, bounds, None, shard.live)); ev[2].record(stream)
                ctx.stage_device(1, h1, h2, weight_scale, mode); ev[3].record(stream)
                (pr.barrier() if pr else gdist.exchange_rows(h2, bounds, None, shard.live)); ev[4].record(stream)
                ctx.stage_device(2, h2, scores, weight_scale, mode); ev[5].record(stream)
                ev[5].synchronize()
                for i, k in enumerate(names):
                    acc[k].append(ev[i].elapsed_time(ev[i + 1]))
            phase_ms = {k: float(np.median(v)) for k, v in acc.items()}
        # ---- per-stage kernel timing for the roofline (dominant kernel = stage 1) ----------------
        stage_ms = [[], [], []]
        if world == 1:
            for rep in range(max(5, min(args.steps, 20))):
                for st, (src, dst) in enumerate(((x_full, h1), (h1, h2), (h2, scores))):
                    flush.fill_(rep & 0xFF)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    ctx.stage_device(st, src, dst, weight_scale, mode)
                    b.record(stream)
                    b.synchronize()
                    stage_ms[st].append(a.elapsed_time(b))
        # ---- the other arithmetic mode, same graph, for the record (N = 1) ----------------------------
        other = None
        if world == 1:
            omode = pkg.MODE_FAST if mode == pkg.MODE_EXACT else pkg.MODE_EXACT
            for _ in range(3):
                ctx.forward_device(x_full, weight_scale, scores, omode)
            oms = []
            for i in range(max(5, min(args.steps, 20))):
                flush.fill_(i & 0xFF)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ctx.forward_device(x_full, weight_scale, scores, omode)
                b.record(stream)
                b.synchronize()
                oms.append(a.elapsed_time(b))
            other = {"mode": "fast" if omode == pkg.MODE_FAST else "exact", "ms_per_step": float(np.mean(oms)),
                     "value": max(e_total, 1) / (float(np.mean(oms)) * 1e-3), "unit": UNIT}
        torch.cuda.synchronize()

    # max over ranks of the per-rank timed total
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = max(e_total, 1) * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer API ------------------------------------------------
    e2e = None
    x_host = x_full.cpu().numpy()
    if world == 1:
        k = max(3, min(args.steps, 50))
        # (i) CSR resident: per step only x goes up and the scores come back
        for _ in range(3):
            ctx.forward(x_host, weight_scale, mode)
        t1 = time.perf_counter()
        for _ in range(k):
            out_host = ctx.forward(x_host, weight_scale, mode)
        res_s = (time.perf_counter() - t1) / k
        # (ii) the whole predict() input per step, as the drop-in calls the ABI: the CSR sits in
        # the context's pinned host buffers (where the drop-in's extraction writes it) and is
        # uploaded -- copies, checks, degree schedule -- inside the timed region, then forward
        srp, scol, sW, sNW = ctx.graph_staging(n, m)
        srp[:] = shard.row_ptr.cpu().numpy()
        scol[:] = shard.col.cpu().numpy().view(np.uint32)
        sW[:] = shard.weights.cpu().numpy().view(np.uint32)
        sNW[:] = shard.nw.cpu().numpy().view(np.uint32)

        def cold_step():
            ctx.graph_upload(srp, scol, sW, sNW)
            return ctx.forward(x_host, weight_scale, mode)
        for _ in range(3):
            cold_step()
        t1 = time.perf_counter()
        for _ in range(k):
            out_cold = cold_step()
        e2e_s = (time.perf_counter() - t1) / k
        assert np.array_equal(out_cold.view(np.uint32), out_host.view(np.uint32))
        t1 = time.perf_counter()
        for _ in range(k):
            ctx.graph_upload(srp, scol, sW, sNW)
        up_s = (time.perf_counter() - t1) / k
        h2d = int(srp.nbytes + scol.nbytes + sW.nbytes + sNW.nbytes + 4 * n)
        c_abi = {"value": max(e_total, 1) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": int(4 * n), "ms_per_step": e2e_s * 1e3, "graph_upload_ms": up_s * 1e3,
                 "graph_upload_GBps": (h2d - 4 * n) / up_s / 1e9,
                 "what": "gvc_graph_upload() + gvc_forward() per step from pre-filled pinned staging buffers: H2D of the "
                         "packed CSR + x, id/offset checks, degree schedule, 3 fused kernels, D2H of the scores"}
        resident = {"value": max(e_total, 1) / res_s, "unit": UNIT, "ms_per_step": res_s * 1e3,
                    "h2d_bytes_per_step": int(4 * n), "d2h_bytes_per_step": int(4 * n),
                    "what": "gvc_forward() only: H2D of x, 3 fused kernels, D2H of the scores; CSR uploaded once"}
        assert np.isfinite(out_host).all()
        # (iii) THE end-to-end number: gnn::model::predict(in, out, reduction_graph) through the
        # reference's own C++ interface (the call src/GNN_VC.cpp:192 makes), on a reduction_graph built
        # from the same edge list: adjacency read through begin(u)/end(u)/W/NW, staged, uploaded,
        # compacted and scheduled on the device, forward, scores back in the host matrix -- every step.
        e2e = None
        if g_cpu is not None:
            from gnn_mwvc_b200 import dropin as gdrop
            try:
                dr = gdrop.Dropin()
            except FileNotFoundError as ex:
                dr = None
                print(f"bench: {ex}", file=sys.stderr)
            if dr is not None:
                os.environ["GVC_MODE"] = args.mode
                os.environ["GVC_DEVICE"] = str(local_rank)
                dm = dr.model(gdrop.model_text(layers))
                dg = dr.graph(n, g_cpu["eu"], g_cpu["ev"], g_cpu["w"])
                first = None
                for i in range(3):
                    out_pred = dr.predict(dm, dg, x_host, weight_scale)
                    first = first if first is not None else dr.last_seconds
                secs = []
                for _ in range(k):
                    out_pred = dr.predict(dm, dg, x_host, weight_scale)
                    secs.append(dr.last_seconds)
                pred_s = float(np.mean(secs))
                assert np.array_equal(out_pred.view(np.uint32), out_host.view(np.uint32)), "predict() != C ABI forward"
                e2e = {"value": max(e_total, 1) / pred_s, "unit": UNIT, "h2d_bytes_per_step": int(4 * m + 16 * n + 4 * n),
                       "d2h_bytes_per_step": int(4 * n), "ms_per_step": pred_s * 1e3, "first_call_ms": first * 1e3,
                       "what": "gnn::model::predict(in, out, reduction_graph) through the reference's C++ interface "
                               "(host/gvc_dropin_capi.cpp), per step: adjacency read from the reduction_graph through begin(u)/end(u), "
                               "streamed through a ring of pinned slots (edge span + per-vertex ranges, weights, x), packed CSR, "
                               "checks and degree schedule built on the device, 3 fused kernels, D2H of the scores "
                               "into the host matrix; wall clock of predict()",
                       "c_abi": c_abi, "csr_resident": resident}
                dr.graph_destroy(dg)
                dr.model_destroy(dm)
        if e2e is None:
            e2e = dict(c_abi, c_abi=None, csr_resident=resident,
                       what=c_abi["what"] + " (predict() leg not run: graph too large for a host reduction_graph "
                                            "in the bench's time, or the drop-in binding is not built)")
    else:
        # every rank takes its shard's CSR from pinned host buffers, uploads it (copies, checks, degree
        # schedule), copies x, runs the three stages with the two exchanges and reads its scores back
        xp = torch.from_numpy(x_host).pin_memory()
        sp = torch.empty(shard.n_local, dtype=torch.float32).pin_memory()
        xd = torch.empty_like(x_full)
        srp, scol, sW, sNW = ctx.graph_staging(shard.n_local, shard.nnz)
        srp[:] = shard.row_ptr.cpu().numpy()
        scol[:] = shard.col.cpu().numpy().view(np.uint32)
        sW[:] = shard.weights.cpu().numpy().view(np.uint32)
        sNW[:] = shard.nw.cpu().numpy().view(np.uint32)
        tail = int(_perm[n - 1].item()) if n % 2 else None
        want_scores = scores.clone()
        k = max(3, min(args.steps, 50))
        with torch.cuda.stream(stream):
            def e2e_step(with_upload):
                if with_upload:
                    ctx.graph_upload(srp, scol, sW, sNW, n_global=g.n, v_begin=shard.v_begin, v_end=shard.v_end)
                    ctx.graph_set_tail(tail)
                xd.copy_(xp, non_blocking=True)
                gdist.sharded_forward(ctx.stage_device, shard, xd, h1, h2, scores, weight_scale, mode, peer_rows=pr)
                sp.copy_(scores, non_blocking=True)
                stream.synchronize()
            times = {}
            for with_upload in (False, True):
                for _ in range(3):
                    e2e_step(with_upload)
                barrier()
                t1 = time.perf_counter()
                for _ in range(k):
                    e2e_step(with_upload)
                barrier()
                t = torch.tensor([(time.perf_counter() - t1) / k], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times[with_upload] = float(t.item())
        assert torch.equal(scores, want_scores), "scores changed after re-uploading the shard"
        csr = torch.tensor([srp.nbytes + scol.nbytes + sW.nbytes + sNW.nbytes], device=dev, dtype=torch.float64)
        dist.all_reduce(csr)
        e2e_s = times[True]
        e2e = {"value": e_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(csr.item()) + int(4 * n) * world,
               "d2h_bytes_per_step": int(4 * n), "ms_per_step": e2e_s * 1e3,
               "what": "per rank and step: gvc_graph_upload_shard() of its shard's CSR from pinned host memory (copies, "
                       "id/offset checks, degree schedule), pinned H2D of x (replicated), 3 fused kernels + 2 NCCL row "
                       "exchanges, D2H of its score slice; wall clock, max over ranks",
               "csr_resident": {"value": e_total / times[False], "unit": UNIT, "ms_per_step": times[False] * 1e3,
                                "h2d_bytes_per_step": int(4 * n) * world, "d2h_bytes_per_step": int(4 * n),
                                "what": "the same without the per-step shard upload"}}

    # ---- N > 1: the sharded scores against a 1-GPU forward of the same graph (outside any timed region)
    parity = None
    if world > 1:
        torch.cuda.synchronize()
        gathered = gdist.gather_scores(scores, bounds)
        one = pkg.Context(local_rank)
        one.model_upload(layers)
        one.graph_adopt(g.row_ptr.to(torch.int32).contiguous(), g.col, g.weights, g.nw)
        one.graph_set_tail(int(_perm[n - 1].item()) if n % 2 else None)
        one_scores = torch.empty(g.n, device=dev)
        one.forward_device(x_full, weight_scale, one_scores, mode)
        one.sync()
        torch.cuda.synchronize()
        if mode == pkg.MODE_EXACT:
            same = bool(torch.equal(gathered.view(torch.int32), one_scores.view(torch.int32)))
        else:   # fast mode splits hub sums by shard-local chunk lists: equal within the fast tolerance
            same = bool(((gathered - one_scores).abs() <= 1e-4 * one_scores.abs().clamp_min(1e-30)).all())
        ok = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity = {"equal": bool(ok.item()), "compared": "bit for bit" if mode == pkg.MODE_EXACT else "within 1e-4 relative",
                  "checksum_sharded": int(gathered.view(torch.int32).to(torch.int64).sum().item()),
                  "checksum_1gpu": int(one_scores.view(torch.int32).to(torch.int64).sum().item())}
        one.close()
        del one_scores, gathered

    if rank == 0:
        peak, peak_src = load_peaks()
        b0, b1, b2 = algorithmic_bytes(n, m)
        roof = None
        if world == 1 and stage_ms[1]:
            t_s1 = float(np.mean(stage_ms[1])) * 1e-3
            ach = b1 / t_s1 / 1e9
            traffic = None
            tp = ROOT / "profiles" / "traffic.json"
            if tp.exists():
                traffic = json.loads(tp.read_text()).get(wl_name, {}).get(f"stage1_{args.mode}")   # ncu, per launch, this workload
            roof = {"bound": "hbm", "kernel": f"stage_kernel<1,{args.mode}> (graph layer w=16 + 35->32->32->16)",
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": b1,
                    "avg_launch_ms": t_s1 * 1e3,
                    "stage_ms": [float(np.mean(s)) for s in stage_ms],
                    "forward_frac": (b0 + b1 + b2) / (total_ms / args.steps * 1e-3) / 1e9 / peak}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if g_cpu is None:      # a workload too large for the CPU leg's budget: a scale-20 sample of the generator
                gs = graphs.rmat_graph(20, 16, seed=42, device=dev)
                g_cpu = {"n": gs.n, "eu": gs.eu.cpu().numpy().astype(np.uint32), "ev": gs.ev.cpu().numpy().astype(np.uint32),
                         "w": gs.weights.cpu().numpy().view(np.uint32).copy(), "sample": True}
            cpu = cpu_baseline_leg(layers, g_cpu)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "vertices": n, "edges": e_total, "nnz": m, "mode": args.mode,
                       "weights": "trained GNN_VC model (tests/golden/mwvc_model.npz)",
                       "l2": "256 MiB flush write between timed steps; working set (CSR + rows) also exceeds the 126 MB L2",
                       "sharding": "single GPU" if world == 1 else f"{world} equal vertex ranges of the relabelled graph (vertices dealt to the shards by descending degree: equal counts, equal nnz), " + ("the stage kernels store every 16-float row into the h buffers of the ranks that own a neighbour of its vertex, over NVLink (CUDA IPC peer memory), a one-element NCCL all-reduce as barrier after stages 0 and 1" if pr else "NCCL all-gather of the 16-float rows of the non-isolated vertices after stages 0 and 1" + (f" [{pr_note}]" if pr_note else "")),
                       "graph_generation_s": round(gen_s, 2)},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "phase_ms": phase_ms, "other_mode": other, "parity_vs_1gpu": parity,
            "step_ms_min_med_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))],
        }
        emit(line)
    # ---- teardown in dependency order, then a normal interpreter exit (exit hooks must run) ----------
    torch.cuda.synchronize()
    if pr is not None:
        h1 = h2 = None
        pr.close()
    del h1, h2, scores, flush, x_full, shard, rp32, g
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def own_stdout():
    """ONE JSON line on stdout is the contract, and libraries print there too (NCCL's version
    banner, for one): keep the real stdout for emit(), send everything else to stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    ap.add_argument("--workload", default="rmat", choices=["rmat", "grid", "er", "isolated"])
    ap.add_argument("--scale", type=int, default=0, help="override the R-MAT scale")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: rows delivered by the stage kernels through peer memory, or NCCL all-gathers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    own_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main() or 0)
