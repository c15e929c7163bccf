#!/usr/bin/env python
"""bench.py — batched HNSW search QPS at recall@10 >= 0.95 on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                         reference arm: the CPU path on the box's host cores

A "step" = one pass of the hot path (Hnsw::search, lib.rs:352-383) over one batch of `--batch` synthetic queries.
Workload at N=1 = BASELINE.json configs[1]: 1M x 128 f32 "SIFT-shaped" synthetic points, M=32, ef_construction=100,
ef_search=100 (raised only if recall@10 < 0.95), batch = 10k queries.  N>1: every rank holds a replica of the index
and searches its own batch (queries are independent objects -> no data-path collective; "scaling": "weak").
`value` = queries/s with queries + outputs resident in HBM; `e2e` = the same through idb_search_batch_f32 with pinned
HOST buffers (H2D + D2H inside the timed region).  The index (512 MB points + 308 MB adjacency) is far larger than the
126 MB L2 and every step uses a different query batch, so no L2 flush is needed between iterations.
torch is used only as plumbing: device buffers, CUDA events, torch.distributed, and the brute-force recall reference.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))

from tests import datagen  # noqa: E402  (seeded synthetic data shared with the tests)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE K1 launch at the headline config, from the committed ncu --set full
    capture of the current kernel (profiles/k1_ncu_traffic.json names the capture it was read from)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")))
        return float(d["dram_bytes_per_launch"]), d["source"]
    except Exception:  # noqa: BLE001
        return None, None


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def host_threads():
    n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return n


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md), via NVML every ~10 ms."""

    def __init__(self, gpu_index):
        self.rows, self.gpu, self._stop, self.th, self.h, self.nv = [], gpu_index, False, None, None, None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h, self.nv = nv.nvmlDeviceGetHandleByIndex(idx), nv
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:  # noqa: BLE001
            log("clock sampler unavailable:", e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), sm, r, pw))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def window(self, t0, t1):
        out = [r for r in self.rows if t0 <= r[0] <= t1]
        return out or [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or self.rows[-3:]

    def stop(self):
        self._stop = True

    def summarize(self, rows):
        if not rows or self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({k for k, bit in names.items() for r in rows if r[2] & bit})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[3] for r in rows)}


def make_workload(args):
    gen = datagen.sift_shaped if args.data == "sift" else datagen.uniform
    t = time.time()
    pts = gen(args.n, args.dim, 1)
    log(f"generated {args.n} x {args.dim} {args.data} points in {time.time() - t:.1f}s")
    return pts, gen


def graph_cache_path(args):
    key = f"{args.n}-{args.dim}-{args.data}-{args.M}-{args.efc}-{args.graph}-{args.seed}-v2"
    d = os.environ.get("IDB_CACHE", os.path.join(ROOT, "gpurun_cache"))
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "graph_" + hashlib.sha1(key.encode()).hexdigest()[:16] + ".npz")


def permute_points(pts, ids):
    p = np.empty_like(pts)
    p[ids] = pts  # points[pid] = rows[orig]   (lib.rs:262-270)
    return p


def obtain_graph(args, pts, device):
    """Returns (points in PointId order, zero, upper list, how).  Setup, never timed.
    The graph (not the points, which are regenerated from the seed) is cached under gpurun_cache/ (git-ignored)."""
    from instant_distance_b200 import _abi

    cp = graph_cache_path(args)
    if os.path.exists(cp) and not args.no_cache:
        z = np.load(cp)
        upper = [z[f"u{i}"] for i in range(int(z["n_upper"]))]
        log(f"graph loaded from cache {cp}")
        return permute_points(pts, z["ids"]), z["zero"], upper, str(z["how"]) + " [cached]"
    how = None
    if args.graph == "gpu":
        try:
            t = time.time()
            ix, ids = _abi.Index.build(pts, M=args.M, ef_construction=args.efc, ef_search=args.ef, seed=args.seed, device=device)
            _, zero, upper = ix.export_graph()
            ix.close()
            how = f"GPU Builder::build ({time.time() - t:.1f}s)"
        except _abi.IdbError as e:
            if e.status != _abi.ERR_UNSUPPORTED:
                raise
            log("GPU build unavailable:", e)
    if how is None:
        from oracle import oracle as O  # setup only: the reference algorithm builds the graph the GPU then searches

        T = host_threads()
        t = time.time()
        ix, ids = O.build(pts, M=args.M, ef_construction=args.efc, ef_search=args.ef, seed=args.seed, threads=T)
        g = ix.export()
        zero, upper = g.zero, g.upper
        how = f"oracle (reference algorithm) threaded build, {T} threads ({time.time() - t:.1f}s)"
    log("graph:", how)
    if not args.no_cache:
        try:
            np.savez(cp, ids=ids, zero=zero, n_upper=len(upper), how=how, **{f"u{i}": u for i, u in enumerate(upper)})
        except Exception as e:  # cache is best effort
            log("cache write failed:", e)
    return permute_points(pts, ids), zero, upper, how


def brute_force_topk_torch(points_dev, queries, k):
    import torch

    q = torch.from_numpy(queries).to(points_dev.device)
    pn = (points_dev * points_dev).sum(1)
    out = []
    for s in range(0, q.shape[0], 256):
        qq = q[s:s + 256]
        d = pn[None, :] - 2.0 * (qq @ points_dev.T) + (qq * qq).sum(1)[:, None]
        out.append(torch.topk(d, k, dim=1, largest=False).indices.cpu())
    return torch.cat(out).numpy()


def recall_at_k(ids, truth, k=10):
    hit = 0
    for a, b in zip(ids[:, :k], truth[:, :k]):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / (k * len(ids))


def algorithmic_bytes(counters, dim, M, k):
    """B(q) = vec + sum_layers[n_expand_l * row_bytes_l + n_dist_l * vec] + k*8."""
    vec = dim * 4
    c = counters.astype(np.float64)
    per_q = vec + c[:, 0] * (M * 4) + c[:, 1] * vec + c[:, 2] * (2 * M * 4) + c[:, 3] * vec + k * 8
    return per_q


def run_reference(args, rank, world):
    """Reference arm: the CPU path (oracle restatement; the Rust reference cannot be built here) on the host cores."""
    if rank != 0:
        return
    from oracle import oracle as O

    pts, gen = make_workload(args)
    dev_ok = False
    try:
        from instant_distance_b200 import _abi

        dev_ok = _abi.lib().idb_device_count() > 0
    except Exception:
        pass
    if not dev_ok and args.graph == "gpu":
        args.graph = "oracle"
    p, zero, upper, how = obtain_graph(args, pts, 0)
    ix = O.from_graph(O.Graph(p, zero, upper, args.M, args.ef))
    T = host_threads()
    sample = min(args.batch, args.ref_sample)
    qs = [gen(sample, args.dim, 1000 + s) for s in range(args.warmup + args.steps)]
    for s in range(args.warmup):
        ix.search(qs[s], ef_search=args.ef, k=10, threads=T)
    t0 = time.perf_counter()
    for s in range(args.warmup, args.warmup + args.steps):
        ix.search(qs[s], ef_search=args.ef, k=10, threads=T)
    dt = time.perf_counter() - t0
    qps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "batched QPS at recall@10>=0.95 (1M x 128 f32)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, how),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": T, "kind": "port",
                         "sample": f"{sample} queries per step x {args.steps} steps, same graph/ef, {T} threads, one Search per thread"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, how):
    return {"workload": f"{args.n} x {args.dim} f32 {args.data}-shaped synthetic, M={args.M}, ef_construction={args.efc}, "
                        f"ef_search={args.ef}, batch={args.batch} queries/step, k=10",
            "graph": how, "l2": "index >> 126 MB L2 and a fresh query batch per step (no flush needed)",
            "parallelism": f"replica x{args.gpus}, queries sharded" if args.gpus > 1 else "1 GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=10_000)
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--efc", type=int, default=100)
    ap.add_argument("--M", type=int, default=32)
    ap.add_argument("--seed", type=int, default=20260923)
    ap.add_argument("--data", default="sift", choices=["sift", "uniform"])
    ap.add_argument("--graph", default="gpu", choices=["gpu", "oracle"])
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--recall-sample", type=int, default=1000)
    ap.add_argument("--ref-sample", type=int, default=10_000)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)  # timing rule: at least 3 untimed warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from instant_distance_b200 import _abi

    if _abi.lib().idb_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        args.no_cache = True  # every rank builds its own replica on its own GPU; no shared cache file
    pts, gen = make_workload(args)
    p, zero, upper, how = obtain_graph(args, pts, local_rank)
    ix = _abi.Index.from_graph(p, zero, upper, args.M, args.ef, device=local_rank)
    ix.set_profiling(True)
    stream = torch.cuda.ExternalStream(ix.stream, device=local_rank)

    # ---- recall gate: raise ef_search until recall@10 >= 0.95 on a sample --------------------------------------
    pdev = torch.from_numpy(p).cuda()
    rq = gen(args.recall_sample, args.dim, 999)
    truth = brute_force_topk_torch(pdev, rq, 10)
    del pdev
    torch.cuda.empty_cache()
    recall = 0.0
    for ef in [args.ef, 128, 160, 200, 256, 320, 400, 512]:
        if ef < args.ef:
            continue
        ids, _, _ = ix.search(rq, ef_search=ef, k=10)
        recall = recall_at_k(ids, truth)
        log(f"recall@10 = {recall:.4f} at ef_search = {ef}")
        if recall >= 0.95:
            args.ef = ef
            break
    else:
        log("WARNING: recall@10 < 0.95 even at ef_search = 512")
        args.ef = 512

    # ---- device-resident arm ("value") -------------------------------------------------------------------------
    total = args.warmup + args.steps
    nq, k = args.batch, 10
    host_q = [gen(nq, args.dim, 5000 + 977 * rank + s) for s in range(total)]
    dq = [torch.from_numpy(q).cuda() for q in host_q]
    d_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    d_dist = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_len = torch.empty((nq,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def step(s):
        ix.search_device(dq[s].data_ptr(), nq, args.ef, k, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr())

    for s in range(args.warmup):
        step(s)
    ix.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches, alg_bytes = [], 0, []
    t_wall0 = time.time()
    ev0.record(stream)
    for s in range(args.warmup, total):
        step(s)
        launches += 2  # K1 search_kernel + its (normally idle) overflow-retry launch; the control-block memset is not a kernel
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.summarize(sampler.window(t_wall0, t_wall1))

    # per-launch K1 duration + algorithmic bytes, measured on a second pass (event sync per step would perturb pass 1)
    for s in range(args.warmup, total):
        step(s)
        ms, _ = ix.last_kernel_ms()
        kernel_ms.append(ms)
        alg_bytes.append(float(algorithmic_bytes(ix.last_counters(nq), args.dim, args.M, k).sum()))

    if world > 1:
        t = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    qps = world * nq * args.steps / (dev_ms / 1e3)

    # ---- e2e arm: public host API, pinned host buffers, H2D + D2H inside the timed region ----------------------
    L = _abi.lib()

    def pinned(nbytes, dtype, shape):
        ptr_ = C.c_void_p()
        _abi.check(L.idb_host_alloc(nbytes, C.byref(ptr_)))
        buf = (C.c_char * nbytes).from_address(ptr_.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr_

    # every step's input already sits in its own PINNED host buffer (as the contract describes); results land in pinned buffers
    hqs = []
    for s_ in range(total):
        hq_, _p = pinned(nq * args.dim * 4, np.float32, (nq, args.dim))
        hq_[...] = host_q[s_]
        hqs.append((hq_, _p))
    hid, hid_p = pinned(nq * k * 4, np.uint32, (nq, k))
    hds, hds_p = pinned(nq * k * 4, np.float32, (nq, k))
    hln, hln_p = pinned(nq * 4, np.uint32, (nq,))

    def e2e_step(s):
        _abi.check(L.idb_search_batch_f32(ix._h, _abi.ptr(hqs[s][0], C.c_float), nq, args.ef, k, _abi.ptr(hid, C.c_uint32),
                                          _abi.ptr(hds, C.c_float), _abi.ptr(hln, C.c_uint32)))

    for s in range(args.warmup):
        e2e_step(s)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, total):
        e2e_step(s)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = world * nq * args.steps / e2e_s
    e2e_ids = hid.copy()

    # ---- CPU baseline (rank 0, N=1 leg only) + parity spot check against it --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:  # reported at N=1 only
        from oracle import oracle as O

        T = host_threads()
        ox = O.from_graph(O.Graph(p, zero, upper, args.M, args.ef))
        sample = min(nq, args.ref_sample)
        qs = host_q[total - 1][:sample]
        ox.search(qs[:256], ef_search=args.ef, k=k, threads=T)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            o_ids, o_dist, o_len = ox.search(qs, ef_search=args.ef, k=k, threads=T)
        cdt = time.perf_counter() - t0
        same = bool((o_ids == e2e_ids[:sample]).all())
        cpu = {"value": sample * reps / cdt, "unit": "queries/s", "cores": T, "kind": "port",
               "sample": f"{sample} queries x {reps} passes of the last step's batch, same graph, ef_search={args.ef}, "
                         f"{T} threads (one Search per thread); GPU ids identical to this run: {same}"}
        if not same:
            log("PARITY FAILURE: GPU ids differ from the oracle on the same graph")
    sampler.stop()

    peak, peak_src = measured_peaks()
    k_ms = float(np.mean(kernel_ms))
    ach = float(np.mean(alg_bytes)) / (k_ms / 1e3) / 1e9
    if rank == 0:
        line = {
            "metric": "batched QPS at recall@10>=0.95 (1M x 128 f32)", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, how), "recall_at_10": recall, "ef_search": args.ef,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * args.dim * 4,
                    "d2h_bytes_per_step": nq * k * 8 + nq * 4},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic()[0] if args.n == 1_000_000 and args.dim == 128 and args.batch == 10_000 else None,
                         "traffic_source": ncu_traffic()[1],
                         "kernel": "search_kernel (K1 search_layer)", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": float(np.mean(alg_bytes)), "peak_source": peak_src},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
