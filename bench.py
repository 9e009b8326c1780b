#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the scene_flow_constructor + scene_flow_clusterer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--pairs B]

A "step" is one pass of the hot path (fused scene-flow+mask kernel, then the GPU clusterer) over one batch of
synthetic frame pairs: by default the ZED 1280x720 sequence of BASELINE.json configs[1] (64 frames = 63 pairs per
GPU; synthetic disparity / flow / ego-motion, seed 2002 + rank).  One process per GPU; with N > 1 every rank owns its
own sequence (weak scaling), pixel data never crosses NVLink and the only collective is an NCCL all_gather of the
per-frame MovingObject records, inside the timed step.

Printed JSON (rank 0, one line):
  value      whole-job frame-pairs/s with inputs resident in HBM (device-timed, max over ranks)
  e2e        the same metric through the host-buffer C-ABI call mod_pipeline_sequence_host: pinned host inputs,
             H2D + kernels + D2H of the object lists inside the timed region
  roofline   the scene-flow+mask kernel: 41 algorithmic B/pixel / its CUDA-event duration, against
             the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU code (oracle/_ref/libmod_ref.so: its unmodified translation units compiled
             against stand-in dependency headers; "kind": "reference") timed on this box's host cores; falls back to
             the oracle restatement ("kind": "port") only where that library was not built
--impl reference times that CPU path alone (all host threads, frame-pair parallel) and prints the same line shape.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_PIXEL = 41  # read disp_now 4 + disp_prev 4 + flow 8, write xyz+v 24 + mask 1
MAX_OBJS = 64              # per-frame object record capacity of the gathered lists
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="zed_1280x720")
    ap.add_argument("--pairs", type=int, default=63, help="frame pairs per GPU per step")
    ap.add_argument("--cpu-pairs", type=int, default=16, help="frame pairs in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_traffic_per_pixel():
    """dram bytes per pixel of the scene-flow kernel from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "sceneflow_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_pixel"])
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived between host times t0 and t1 (the timed region); all if None."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, line in self.lines:
            if t0 is not None and not (t0 <= stamp <= t1):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code -- oracle/_ref/libmod_ref.so, the reference's unmodified hot-path
# translation units compiled against stand-in dependency headers (oracle/Makefile ref_full; kind "reference") --
# or, where that library was not built, the oracle restatement (kind "port")
# ---------------------------------------------------------------------------------------------------------------
def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmod_ref.so")) else "port"


def cpu_sample(cfg_name: str, n_pairs: int, seed_offset: int = 0):
    import dataclasses
    from moving_object_detector_b200 import synth
    cfg = dataclasses.replace(synth.CONFIGS[cfg_name], seed=synth.CONFIGS[cfg_name].seed + seed_offset)
    seq = synth.make_sequence(cfg, n_pairs + 1, device="cpu")
    return cfg, seq


def run_cpu_path(seq, n_threads: int, reps: int, stage_mask: int = 3):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py
    impl = oracle_py
    if cpu_kind() == "reference":
        import ref_py
        ref_py.load()
        impl = ref_py
    disp = seq["disp"].cpu().numpy()
    flows = seq["flows"].cpu().numpy()
    n_pairs = flows.shape[0]
    prev_idx = np.arange(n_pairs, dtype=np.int32)
    now_idx = prev_idx + 1
    sec, n_obj = impl.bench_pipeline(disp, prev_idx, now_idx, flows, seq["cam"], seq["transforms"], seq["dts"],
                                     seq["f"], seq["T"], seq["min_disparity"], seq["max_disparity"],
                                     n_threads=n_threads, reps=reps, stage_mask=stage_mask)
    return n_pairs / sec, sec, n_obj


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from moving_object_detector_b200 import synth
    cores = os.cpu_count() or 1
    n_pairs = cores  # one pair per host thread
    cfg, seq_full = cpu_sample(args.config, n_pairs)

    def take(seq, n):
        sub = dict(seq)
        sub["disp"], sub["flows"] = seq["disp"][:n + 1], seq["flows"][:n]
        sub["transforms"], sub["dts"] = seq["transforms"][:n], seq["dts"][:n]
        return sub

    # bounded sample: probe one step, then shrink the per-step sample so K steps + W warm-ups stay within ~2 minutes
    _, probe_sec, _ = run_cpu_path(seq_full, cores, 1)
    budget = 120.0
    if probe_sec * (args.steps + args.warmup) > budget:
        n_pairs = max(1, min(cores, int(cores * budget / (probe_sec * (args.steps + args.warmup)))))
    threads = min(cores, n_pairs)
    seq = take(seq_full, n_pairs)
    cores = threads
    for _ in range(args.warmup):
        run_cpu_path(seq, cores, 1)
    t0 = time.perf_counter()
    secs = []
    for _ in range(args.steps):
        _, sec, _ = run_cpu_path(seq, cores, 1)
        secs.append(sec)
    total = sum(secs)
    value = n_pairs * args.steps / total
    sample = (f"{n_pairs} frame pairs of {cfg.w}x{cfg.h} (seed {cfg.seed}) per step, constructor+clusterer, "
              f"{cores} std::threads, one pair per thread at a time")
    line = {
        "impl": "reference", "metric": "frame-pairs/s (scene_flow_constructor+clusterer)", "value": value, "unit": "frame-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "mpix_per_s": value * cfg.w * cfg.h / 1e6,
        "config": {"workload": f"{args.config}: ZED {cfg.w}x{cfg.h} stereo sequence, scene_flow_constructor+clusterer",
                   "pairs_per_step": n_pairs, "width": cfg.w, "height": cfg.h},
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": cpu_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t0,
    }
    emit(line)
    return 0


def bind_to_gpu_numa_node(gpu_index: int):
    """Multi-rank runs: keep this rank's threads -- and so the pinned host buffers it allocates (first touch) -- on the
    NUMA node its GPU hangs off, so that eight concurrent host->device streams do not all read through one socket.
    Returns a short description for the JSON line, or None when the topology cannot be read (then nothing changes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        sysfs = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        node = int(open(sysfs + "/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"node{node} ({len(cpus)} cpus)"
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def main_ours(args):
    import dataclasses
    import numpy as np
    import torch
    import torch.distributed as dist
    from moving_object_detector_b200 import (ClustererConfig, Context, OBJECT_DTYPE, STATS_DTYPE,
                                             SceneFlowConstructorConfig, CameraInfo, TransformMsg, synth)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version / debug lines to stdout unless told otherwise; stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    base = synth.CONFIGS[args.config]
    cfg = dataclasses.replace(base, seed=base.seed + rank)
    B = args.pairs
    W, H = cfg.w, cfg.h
    P = W * H
    seq = synth.make_sequence(cfg, B + 1, device=dev)
    disp, flows = seq["disp"], seq["flows"]
    cam = CameraInfo.pinhole(*seq["cam"])
    transforms = [TransformMsg(t[:3], t[3:]) for t in seq["transforms"]]
    sf_cfg, cl_cfg = SceneFlowConstructorConfig(), ClustererConfig()

    ctx = Context(W, H, max_batch=B, device=local_rank)
    pairs = Context.make_pairs([disp[i + 1] for i in range(B)], [disp[i] for i in range(B)],
                               [flows[i] for i in range(B)], transforms, seq["dts"], seq["f"], seq["T"],
                               seq["min_disparity"], seq["max_disparity"])
    points = torch.empty((B, P, 8), dtype=torch.float32, device=dev)
    mask = torch.empty((B, P), dtype=torch.uint8, device=dev)
    gatherer = None
    if world > 1:
        # the only collective (per-frame object lists, never pixel data): one all-gather per step on a
        # side stream, overlapped by the next step's scene-flow kernel, its buffers allocated once
        from moving_object_detector_b200.sharding import ObjectListGatherer
        gatherer = ObjectListGatherer(B, MAX_OBJS * OBJECT_DTYPE.itemsize, STATS_DTYPE.itemsize, dev)
        objs, stats = gatherer.objs, gatherer.stats
    else:
        objs = torch.zeros((B, MAX_OBJS * OBJECT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        stats = torch.zeros((B, STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    sf_events = []

    def step(record_sf: bool):
        if record_sf:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ctx.scene_flow_batched_dev(W, H, pairs, cam, sf_cfg, cl_cfg.dynamic_speed, points, mask, stream=stream)
        if record_sf:
            e1.record()
            sf_events.append((e0, e1))
        if gatherer is not None:
            gatherer.wait()  # the previous step's all-gather still reads objs / stats
        ctx.cluster_batched_dev(W, H, B, points, mask, cl_cfg, objs, stats, MAX_OBJS, stream=stream)
        if gatherer is not None:
            gatherer.launch()

    def barrier():
        if gatherer is not None:
            gatherer.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ctx.launch_count
    barrier()
    host_t0 = time.perf_counter()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step(True)
    if gatherer is not None:
        gatherer.wait()  # the last step's all-gather is inside the timed region
    t_end.record()
    barrier()
    host_t1 = time.perf_counter()
    launches = ctx.launch_count - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    sf_ms = [a.elapsed_time(b) for a, b in sf_events]
    clocks = None
    if rank == 0:
        time.sleep(0.1)
        # nvidia-smi reports every 50 ms; a sample arriving up to one period after the region still describes it
        clocks = sampler.stop(host_t0 + 0.02, host_t1 + 0.06)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_pairs = B * args.steps * world
    value = total_pairs / (elapsed_ms / 1e3)

    st = np.frombuffer(stats.cpu().numpy().tobytes(), dtype=STATS_DTYPE)
    n_objects = int(st["n_objects"].sum())
    bad = int((st["status"] != 0).sum())
    n_dynamic = int(st["n_dynamic"].sum())
    # pixels of one step the filtered residual gate re-evaluated exactly (outside the timed region)
    ctx.filter_fallbacks(reset=True)
    step(False)
    torch.cuda.synchronize()
    filter_fallbacks = ctx.filter_fallbacks(reset=True)

    # ---- e2e through the host-buffer C-ABI call -----------------------------------------------------------------
    e2e = None
    e2e_cloud = None
    if not args.no_e2e:
        import ctypes as C
        from moving_object_detector_b200._lib import Transform
        h_disp = torch.empty((B + 1, H, W), dtype=torch.float32).pin_memory()
        h_flow = torch.empty((B, H, W, 2), dtype=torch.float32).pin_memory()
        h_disp.copy_(disp)
        h_flow.copy_(flows)
        h_objs = torch.zeros((B, MAX_OBJS * OBJECT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        h_stats = torch.zeros((B, STATS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        tr_c = (Transform * B)(*[t.to_c() for t in transforms])
        dt_c = (C.c_double * B)(*[float(d) for d in seq["dts"]])
        e2e_steps = args.e2e_steps or min(args.steps, 10)

        def e2e_run(cloud_buf):
            def one():
                ctx.pipeline_sequence_host(h_disp, h_flow, cam, transforms, seq["dts"], seq["f"], seq["T"],
                                           seq["min_disparity"], seq["max_disparity"], sf_cfg, cl_cfg, h_objs, h_stats,
                                           MAX_OBJS, out_points32=cloud_buf, transforms_c=tr_c, dts_c=dt_c)
            for _ in range(3):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                one()  # synchronous: returns after the last D2H landed
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return B * e2e_steps * world / float(tt.item())

        h2d = (B + 1) * P * 4 + B * P * 8
        d2h = B * (MAX_OBJS * OBJECT_DTYPE.itemsize + STATS_DTYPE.itemsize)
        v = e2e_run(None)
        e2e = {"value": v, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "h2d_gbs_per_gpu": h2d * (v / (B * world)) / 1e9,
               "bound": "PCIe Gen5 x16 host-to-device link (55.5 GB/s measured, profiles/h2d_probe.cu)",
               "host_numa_binding": numa,
               "steps": e2e_steps, "api": "mod_pipeline_sequence_host (pinned host in, MovingObject lists out; "
               "the 32 B/px cloud stays on the GPU as it has no host subscriber)"}
        st_h = np.frombuffer(h_stats.numpy().tobytes(), dtype=STATS_DTYPE)
        assert int(st_h["n_objects"].sum()) == n_objects, "host-path and device-path object counts differ"
        if world == 1:
            h_cloud = torch.empty((B, P, 8), dtype=torch.float32).pin_memory()
            v2 = e2e_run(h_cloud)
            e2e_cloud = {"value": v2, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h + B * P * 32,
                         "note": "same call with the PointCloud2 ~scene_flow payload also downloaded"}
            del h_cloud

    # ---- CPU baseline on rank 0 (N = 1 only): bounded sample of the same workload ---------------------------------
    cpu_baseline = None
    cpu_1thread = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = max(args.cpu_pairs, cores)
        sub = {k: seq[k] for k in ("cam", "f", "T", "min_disparity", "max_disparity")}
        sub["disp"] = disp[:n_cpu + 1].cpu()
        sub["flows"] = flows[:n_cpu].cpu()
        sub["transforms"] = seq["transforms"][:n_cpu]
        sub["dts"] = seq["dts"][:n_cpu]
        v_all, _, n_obj_cpu = run_cpu_path(sub, cores, 2)
        cpu_baseline = {"value": v_all, "unit": "frame-pairs/s", "cores": cores, "kind": cpu_kind(),
                        "sample": f"first {n_cpu} pairs of the same {W}x{H} sequence, constructor+clusterer, "
                                  f"{cores} threads (pair-parallel), best of 2"}
        sub1 = dict(sub)
        n1 = min(4, n_cpu)
        sub1["disp"], sub1["flows"] = sub["disp"][:n1 + 1], sub["flows"][:n1]
        sub1["transforms"], sub1["dts"] = sub["transforms"][:n1], sub["dts"][:n1]
        v_1, _, _ = run_cpu_path(sub1, 1, 1)
        cpu_1thread = {"value": v_1, "unit": "frame-pairs/s", "cores": 1, "kind": cpu_kind(),
                       "sample": f"first {n1} pairs, 1 thread (the reference's construct()/dataCB are single-threaded)"}
        n_obj_gpu = int(st["n_objects"][:n_cpu].sum())
        assert n_obj_cpu == n_obj_gpu, f"CPU oracle found {n_obj_cpu} objects, GPU {n_obj_gpu} on the same pairs"

    if rank == 0:
        peak, peak_src = hbm_peak()
        sf_mean_ms = sum(sf_ms) / len(sf_ms)
        algo_bytes = ALGO_BYTES_PER_PIXEL * P * B
        achieved = algo_bytes / (sf_mean_ms * 1e-3) / 1e9
        tpp = measured_traffic_per_pixel()
        line = {
            "metric": "frame-pairs/s (scene_flow_constructor+clusterer)", "value": value, "unit": "frame-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic", "mpix_per_s": value * P / 1e6,
            "config": {"workload": f"{args.config}: ZED {W}x{H} stereo sequence, scene_flow_constructor+clusterer",
                       "pairs_per_step_per_gpu": B, "width": W, "height": H, "seed": base.seed,
                       "l2": "inputs larger than L2 (%.0f MB read + %.0f MB written per step)" %
                             ((B + 1) * P * 4 / 1e6 + B * P * 8 / 1e6, B * P * 33 / 1e6),
                       "params": {"dynamic_flow_diff": sf_cfg.dynamic_flow_diff, "cluster_size": cl_cfg.cluster_size,
                                  "depth_diff": cl_cfg.depth_diff, "dynamic_speed": cl_cfg.dynamic_speed,
                                  "neighbor_distance": cl_cfg.neighbor_distance},
                       "parallelism": f"frame-pair data parallel x{world}"},
            "roofline": {"bound": "hbm", "kernel": "scene_flow_mask_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "frac_of_8TBps_nominal": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": sf_mean_ms,
                         "traffic": (tpp * P * B) if tpp else None,
                         "share_of_step": sf_mean_ms / (elapsed_ms / args.steps)},
            "e2e": e2e, "e2e_with_cloud": e2e_cloud,
            "cpu_baseline": cpu_baseline, "cpu_baseline_1thread": cpu_1thread,
            "gpu_launches": int(launches),
            "clusterer": {"ms": elapsed_ms / args.steps - sf_mean_ms, "dynamic_pixels_per_step": n_dynamic,
                          "note": "step time minus the scene-flow kernel: 8 clusterer kernels, per-kernel times and "
                                  "L2 atomic sectors in profiles/r01_launches_summary.txt"},
            "filter": {"exact_fallback_pixels_per_step": filter_fallbacks,
                       "fraction_of_pixels": filter_fallbacks / float(P * B)},
            "objects_per_step": n_objects, "frames_with_error_status": bad,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result.  Libraries write there too (NCCL prints its version line with a
    plain printf, whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    claim_stdout()
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
