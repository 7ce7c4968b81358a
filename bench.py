#!/usr/bin/env python
"""bench.py — boosting iterations/sec of the LightGBM-on-Spark training hot path on N B200s.

Contract (driver): `python bench.py --gpus N --steps K --warmup W`; for N>1 it is launched under
torch.distributed.run (one rank per GPU).  A step = one boosting iteration (LGBM_BoosterUpdateOneIter
through the C ABI) of LightGBMClassifier(binary) on the synthetic 100M x 512 dense matrix that
BASELINE.json's metric is quoted on; rows are split evenly over the ranks (strong scaling), the per-split
histogram reduction is an NCCL int64 allreduce.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference path (the oracle; the real lightgbmlib
3.2.110 cannot be built or installed here — BASELINE.md §2) on a bounded row sample of the same workload.
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

SEED = 2025
KIND_BINARY = 1
CONFIG_ITERS = 100        # numIterations default of the estimator (LightGBMParams.scala:318-322): amortisation base for ingestion
DS_PARAMS = "max_bin=255 is_pre_partition=True bin_construct_sample_cnt=200000 num_threads=0"


def booster_params(num_machines):
    return ("metric= boost_from_average=true is_pre_partition=True boosting_type=gbdt tree_learner=data_parallel top_k=20 "
            "num_iterations=100 learning_rate=0.1 num_leaves=31 max_bin=255 bagging_fraction=1.0 pos_bagging_fraction=1.0 "
            "neg_bagging_fraction=1.0 bagging_freq=0 bagging_seed=3 early_stopping_round=0 feature_fraction=1.0 max_depth=-1 "
            "min_sum_hessian_in_leaf=0.001 num_machines=%d verbosity=-1 lambda_l1=0.0 lambda_l2=0.0 metric= "
            "min_gain_to_split=0.0 max_delta_step=0.0 min_data_in_leaf=20 objective=binary num_threads=0 is_unbalance=false" % num_machines)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_dataset(capi, n_local, F, row_start, ingest):
    """Synthesise this rank's row shard on the device in chunks and ingest it.
    ingest == 'host': every chunk is first staged in pinned HOST memory (untimed) and then pushed through
    LGBM_DatasetPushRows from the host pointer (H2D + binning timed by the library with CUDA events) —
    the streaming analogue of Spark rows arriving at the task.  ingest == 'device': pushed from HBM."""
    sample_rows = capi.sample_indices(n_local, 200000, 1)
    sample, _ = capi.synthetic_rows((sample_rows.astype(np.int64) + row_start).astype(np.int32), F, SEED, KIND_BINARY)
    ds = capi.Dataset.from_sampled_columns(sample, n_local, DS_PARAMS)
    chunk = min(n_local, max(1, (1 << 30) // (F * 4)))          # ~1 GiB of f32 per chunk
    dev_x = capi.DeviceBuffer(chunk * F * 4)
    dev_y = capi.DeviceBuffer(chunk * 4)
    label = np.empty(n_local, dtype=np.float32)
    pinned = capi.PinnedBuffer(chunk * F * 4) if ingest == "host" else None
    host_bytes = 0
    for off in range(0, n_local, chunk):
        rows = min(chunk, n_local - off)
        capi.synthetic_fill(dev_x.ptr, dev_y.ptr, row_start + off, rows, F, SEED, KIND_BINARY)
        capi.memcpy(label[off:off + rows].ctypes.data, dev_y.ptr, rows * 4)
        if pinned is not None:
            capi.memcpy(pinned.ptr, dev_x.ptr, rows * F * 4)              # stage the chunk on the host (untimed)
            ds.push_rows(pinned.ptr, off, nrow=rows, ncol=F, dtype_code=capi.DTYPE_FLOAT32)
            host_bytes += rows * F * 4
        else:
            ds.push_rows(dev_x.ptr, off, nrow=rows, ncol=F, dtype_code=capi.DTYPE_FLOAT32)
    dev_x.free(); dev_y.free()
    if pinned is not None:
        pinned.free()
    ds.set_field("label", label)
    return ds, label, host_bytes


def cpu_reference_run(n_total, F, sample_rows, steps, warmup, use_gpu_generator=True):
    """The reference's CPU path restated (oracle, OpenMP over all host cores) on a bounded row sample of the
    same synthetic workload.  Returns (iters_per_sec_on_sample, extrapolated_iters_per_sec, info)."""
    from oracle import oracle as O
    rows = np.arange(sample_rows, dtype=np.int32)
    if use_gpu_generator:
        from mmlspark_b200 import capi
        X, y = capi.synthetic_rows(rows, F, SEED, KIND_BINARY)       # generator only; no product compute on this arm
    else:
        # same distribution as the device generator (csrc/c_api.cu syn_x / syn_label), drawn with numpy so that the reference arm
        # touches none of the product's code
        rng = np.random.default_rng(SEED)
        U = rng.random((sample_rows, F), dtype=np.float32)
        m = min(F, 16)
        s = (np.sin(6.2831853 * U[:, :m]) * (1.0 + 0.1 * np.arange(m, dtype=np.float32))).sum(axis=1)
        if F >= 2:
            s += 8.0 * (U[:, 0] - 0.5) * (U[:, 1] - 0.5)
        y = (rng.random(sample_rows) < 1.0 / (1.0 + np.exp(-s))).astype(np.float32)
        X = U.astype(np.float64)
        X *= (1.0 + (np.arange(F) % 7))
        X -= (np.arange(F) % 5)
        del U
    ods = O.OracleDataset(X, DS_PARAMS).set_field("label", y)
    ob = O.OracleBooster(ods, booster_params(1))
    for _ in range(warmup):
        ob.update()
    t0 = time.perf_counter()
    for _ in range(steps):
        ob.update()
    dt = time.perf_counter() - t0
    hs, hc = ob.hist_stats()
    ips = steps / dt
    info = {"cores": int(O.lib().orc_num_threads()), "sample_rows": sample_rows, "hist_cells_per_s": hc / hs if hs > 0 else None,
            "hist_share": hs / (dt * (steps + warmup) / steps) if dt > 0 else None}
    return ips, ips * sample_rows / n_total, info


def k4_traffic(rows, feats, world):
    """roofline.traffic: DRAM bytes per K4 launch from the committed ncu capture of this shape (profiles/), else null."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_k4_dram_traffic_%dx%d.json" % (rows, feats))
    if world != 1 or not os.path.exists(path):
        return None
    try:
        return float(json.load(open(path))["bytes_per_launch_avg"])
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--features", type=int, default=512)
    ap.add_argument("--ingest", default="host", choices=["host", "device"])
    ap.add_argument("--cpu-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N, F = args.rows, args.features
    config = {"workload": "LightGBMClassifier binary, synthetic %dx%d dense f32, 255 bins, num_leaves=31, lr=0.1 (BASELINE.json configs[2], the config the metric is quoted on)" % (N, F),
              "rows": N, "features": F, "parallelism": "data_parallel x%d (rows sharded, NCCL int64 histogram allreduce)" % world,
              "l2_policy": "inputs (%.1f GB of bins per rank) are far larger than the 126 MB L2" % (N / world * F / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        ips_s, ips_x, info = cpu_reference_run(N, F, args.cpu_sample_rows, max(args.steps, 1), max(args.warmup, 1), use_gpu_generator=False)
        line = {"impl": "reference", "metric": "boosting_iters_per_sec", "value": ips_x, "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 / ips_x, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64 histograms over u8 bins (fp32 gradients)", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": ips_x, "unit": "iters/s", "cores": info["cores"], "kind": "port",
                                 "sample": "first %d rows of the same synthetic matrix, %d timed iterations at %.3f iters/s on the sample, scaled by rows (x%d/%d); "
                                           "histogram throughput %.3g cells/s" % (args.cpu_sample_rows, args.steps, ips_s, args.cpu_sample_rows, N, info["hist_cells_per_s"] or 0)},
                "e2e": {"value": ips_x, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    dist = None
    if world > 1:
        # torch first: its bundled libnccl.so.2 (2.28) must be the one the process binds; the engine's NCCL calls
        # (CommInitRank / AllReduce / AllGather) then resolve to the same library
        import torch  # noqa: F401
        import torch.distributed as dist_mod
    from mmlspark_b200 import capi
    capi.load()
    if world > 1:
        dist = dist_mod
        dist.init_process_group("gloo")          # host-side barrier / max-reduce only; training traffic is NCCL inside the library
    capi.set_device(local_rank)
    if world > 1:
        base = int(os.environ.get("MASTER_PORT", "29500")) + 512
        machines = ",".join("127.0.0.1:%d" % (base + r) for r in range(world))
        capi.network_init(machines, base + rank, 120, world)

    n_local = N // world + (1 if rank < N % world else 0)
    row_start = rank * (N // world) + min(rank, N % world)
    t_build = time.perf_counter()
    ds, label, host_bytes = build_dataset(capi, n_local, F, row_start, args.ingest)
    ingest_ms = ds.ingest_ms()
    build_s = time.perf_counter() - t_build
    bst = capi.Booster(ds, booster_params(world))
    bst.set_profile(True)

    def barrier():
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        bst.update_one_iter()
    bst.get_timing(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        bst.update_one_iter()           # returns after the iteration's tree was read back (device sync inside)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    barrier()
    tm = bst.get_timing()
    dev_ms = tm["total_ms"]
    # max over ranks
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall * 1000.0, ingest_ms, tm["hist_ms"]], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, ingest_ms_max, hist_ms = [float(x) for x in t]
        r = torch.tensor([float(tm["hist_rows"])], dtype=torch.float64)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        hist_rows_all = float(r[0])
    else:
        wall_ms, ingest_ms_max, hist_ms, hist_rows_all = wall * 1000.0, ingest_ms, tm["hist_ms"], float(tm["hist_rows"])
    if rank != 0:
        return
    steps = args.steps
    value = steps / (dev_ms / 1000.0)
    ms_per_step = dev_ms / steps
    # roofline of the dominant kernel (K4), per rank: algorithmic bytes / measured launch time
    launches = tm["hist_launches"]
    trees = steps
    rows_rank0 = float(tm["hist_rows"])
    nonroot_rows = max(rows_rank0 - trees * n_local, 0.0)
    algo_bytes = rows_rank0 * (F + 8.0) + nonroot_rows * 4.0 + launches * F * 255 * 16.0
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (tm["hist_ms"] / 1000.0) / 1e9 if tm["hist_ms"] > 0 else None
    roofline = {"bound": "hbm", "kernel": "k4_hist_build_ws<4>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "traffic": k4_traffic(args.rows, args.features, world), "launches": launches, "avg_launch_ms": tm["hist_ms"] / max(launches, 1),
                "cells_per_s": rows_rank0 * F / (tm["hist_ms"] / 1000.0) if tm["hist_ms"] > 0 else None,
                "k4_share_of_step": tm["hist_ms"] / dev_ms if dev_ms > 0 else None,
                "co_limit": "shared-memory ATOMS issue rate (profiles/r01_ubench_smem_scatter.json): 4 native 32-bit atomics per cell"}
    # e2e: per-iteration C-ABI wall time (tree read-back inside) + host ingestion amortised over the fit's iterations
    t_iter = wall_ms / steps
    e2e_ms = t_iter + ingest_ms_max / CONFIG_ITERS
    tree_bytes = 31 * 80 + 256
    e2e = {"value": 1000.0 / e2e_ms, "unit": "iters/s", "h2d_bytes_per_step": int(host_bytes / CONFIG_ITERS), "d2h_bytes_per_step": tree_bytes,
           "ingest_ms": ingest_ms_max, "ingest_GBps_host_to_bins": host_bytes / 1e9 / (ingest_ms_max / 1000.0) if ingest_ms_max > 0 and host_bytes else None,
           "note": "UpdateOneIter through the C ABI incl. tree read-back, plus LGBM_DatasetPushRows ingestion from pinned host chunks "
                   "(H2D + binning) amortised over the %d iterations of the fit; ingest mode=%s" % (CONFIG_ITERS, args.ingest)}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ips_s, ips_x, info = cpu_reference_run(N, F, args.cpu_sample_rows, 3, 1)
        cpu = {"value": ips_x, "unit": "iters/s", "cores": info["cores"], "kind": "port",
               "sample": "first %d rows of the same matrix, 3 timed iterations (%.3f iters/s on the sample), scaled by rows; hist %.3g cells/s" % (
                   args.cpu_sample_rows, ips_s, info["hist_cells_per_s"] or 0)}
    line = {"metric": "boosting_iters_per_sec", "value": value, "unit": "iters/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64 fixed-point histograms over u8 bins (fp32 gradients, fp64 split gains)", "data": "synthetic", "config": config,
            "hist_rows_x_feats_per_sec": hist_rows_all * F / (hist_ms / 1000.0) if hist_ms > 0 else None,
            "histogram_reduce": ("fused reduce-scatter+scan over NVLink peer memory (k_scan_dp)" if bst.get_info()["fused_peer_reduce"] else
                                 ("ncclAllReduce int64" if world > 1 else "none (1 rank)")),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": tm["launches"], "clocks": clocks,
            "dataset_build_s": build_s}
    print(json.dumps(line))
    if world > 1:
        capi.network_free()


if __name__ == "__main__":
    main()
