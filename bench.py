#!/usr/bin/env python
"""bench.py — front-end frames/s of the B200 tracking front-end (BASELINE.json metric) on config C2:
640x480 mono, ref=4 MV chaining (max_ref 3), textured plane, 64 streams batched per GPU.

One "step" = one pass of the hot path over one batch: 64 streams x 16 new frames go through
ingest -> raster (hop lists, kps, per-pixel slot grid) -> track propagation -> [join, pose, frustum, join, pose].

  python bench.py [--gpus N] [--steps K] [--warmup W]          product arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                          reference arm: the CPU restatement of the reference's
                                                                front-end on the host cores (the reference itself cannot
                                                                be built in this image)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))

from movfe import synth, types as T  # noqa: E402

W, H = 640, 480
S_PER_GPU = 64
F = int(os.environ.get("BENCH_F", 16))   # frames per stream per step (window length)
MAX_REF = 3         # ref=4 chaining -> reference indices 0..3
CPU_REPEATS = 4     # repeats of the cpu_baseline sample (about 10-20 s of CPU work)
REF_FRAMES = int(os.environ.get("BENCH_REF_FRAMES", 100))    # frames per stream of one CPU sample (reference arm / cpu_baseline)
N_BASE = 8          # distinct synthetic clips; stream s replays clip s % N_BASE (every stream is processed separately)
MAX_RECORDS = 4800
MAX_TRACKS = 8192   # the reference's tables are unbounded; the C2 tables plateau near 4000 entries
METRIC = "front_end_frames_per_s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL's version banner) write to fd 1 as well, so fd 1 is pointed at
    stderr for the whole run and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_clips(n_frames, n_base=N_BASE, with_grey=True):
    clips = []
    for b in range(n_base):
        spec = synth.Spec(W, H, n_frames=n_frames, refs=MAX_REF + 1, seed=0x5EED0002 + 977 * b, phase=0.37 * b)
        recs, off, flags = synth.make_records(spec)
        grey = synth.make_grey(spec) if with_grey else None
        clips.append(dict(spec=spec, recs=recs, off=off, flags=flags, grey=grey))
    return clips


def pack_window(clips, S, f0, f1, pinned=None):
    """Stream-major packed inputs for frames [f0,f1) of S streams (stream s replays clip s % len(clips))."""
    import torch
    per = []
    for c in clips:
        r0, r1 = c["off"][f0], c["off"][f1]
        per.append((c["recs"][r0:r1], c["off"][f0:f1 + 1] - r0, c["flags"][f0:f1]))
    n = f1 - f0
    tot = sum(len(per[s % len(per)][0]) for s in range(S))
    recs = torch.empty(max(tot, 1) * 40 + 16, dtype=torch.uint8, pin_memory=pinned is not False)
    off = torch.empty(S * n + 1, dtype=torch.int64, pin_memory=pinned is not False)
    flags = torch.empty(S * n, dtype=torch.uint8, pin_memory=pinned is not False)
    grey = None
    if clips[0]["grey"] is not None:
        grey = torch.empty((S, n, H, W), dtype=torch.uint8, pin_memory=pinned is not False)
    rv = recs.numpy()[:tot * 40].view(T.MV_RECORD)
    ov, fv = off.numpy(), flags.numpy()
    pos = 0
    for s in range(S):
        r, o, fl = per[s % len(per)]
        rv[pos:pos + len(r)] = r
        ov[s * n:(s + 1) * n] = o[:-1] + pos
        fv[s * n:(s + 1) * n] = fl
        if grey is not None:
            grey.numpy()[s] = clips[s % len(clips)]["grey"][f0:f1]
        pos += len(r)
    ov[S * n] = pos
    return dict(recs=recs, off=off, flags=flags, grey=grey, n_records=tot, n=n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md). The sampler is started before
    the warm-up steps (nvidia-smi needs about a second to produce its first line) and every sample taken between
    mark_load() and stop() - warm-up and timed steps, back to back - is kept."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path, self.t_load = device, None, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 3.0 and os.path.getsize(self.path) == 0:   # wait for the first sample
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark_load(self):
        self.t_load = time.time()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        t_end = time.time()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t_load is not None and not (self.t_load - 0.02 <= ts <= t_end + 0.02):
                    continue
                sm.append(float(f[2]))
                mx.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm ------
def cpu_frontend_sample(clips, n_frames, threads, repeats=1):
    """Times the oracle's whole front-end (raster -> extract -> joins/frustum -> pose x2) on `threads` host threads,
    one stream per thread, n_frames frames each, `repeats` times over. Returns (frames/s, seconds)."""
    from oracle import pyoracle as orc
    orc.lib()
    cam = clips[0]["spec"].camera()
    pp = T.pose_params()
    jobs = []
    for t in range(threads):
        c = clips[t % len(clips)]
        sp = c["spec"]
        off = c["off"][:n_frames + 1]
        recs = c["recs"][:off[-1]]
        # map points from the frame-0 seeds (same construction as the GPU arm)
        clip0 = orc.Clip(W, H, recs[:off[1]], off[:2], c["flags"][:1], MAX_REF)
        t0, _, _, _ = orc.extract_frame(W, H, c["flags"][0], c["grey"][0], clip0.grid(0), clip0.hops(0), clip0.kps(0),
                                        clip0.coverage(0), np.zeros(0, T.TRACK), 0, max_tracks=MAX_TRACKS)
        mp = synth.map_from_tracks(sp, t0, synth.pose_at(sp, 0))
        jobs.append((recs, off, c["flags"][:n_frames], c["grey"][:n_frames], mp, synth.pose_struct(synth.pose_at(sp, 0))))
    res = [None] * threads

    def run(i):
        recs, off, fl, grey, mp, p0 = jobs[i]
        for _ in range(repeats):
            res[i] = orc.frontend_run(W, H, recs, off, fl, grey, None, mp, p0, cam, pp, max_ref=MAX_REF, max_tracks=MAX_TRACKS,
                                      n_kf_points=len(mp) // 2)

    ths = [threading.Thread(target=run, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return threads * n_frames * repeats / dt, dt


def run_reference(args):
    """Reference arm: the reference's CPU front-end (oracle port) on all host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_frames = REF_FRAMES
    clips = make_clips(n_frames)
    vals = []
    for i in range(args.warmup + args.steps):
        fps, dt = cpu_frontend_sample(clips, n_frames, cores)
        if i >= args.warmup:
            vals.append((fps, dt))
        log("reference step %d: %.1f frames/s (%.2fs)" % (i, fps, dt))
    fps = float(np.mean([v[0] for v in vals]))
    sample = "%d streams (one per host thread) x %d frames of the C2 clip per step" % (cores, n_frames)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([v[1] for v in vals])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32/f32 raster+tracks, f64 pose", "data": "synthetic",
            "config": config_dict(1, cores), "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def grid_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one grid_kernel launch of THIS workload, from the committed
    `ncu --set full` capture (profiles/grid_kernel_traffic.json, written by scripts/ncu_traffic.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "grid_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    # the capture is taken on a shorter window (fewer frames per launch); traffic scales with the frames per launch
    return float(t["dram_bytes_per_frame_stream"]) * S_PER_GPU * F


def config_dict(n_gpus, cores=None):
    return {"workload": "C2: 640x480 mono, x264-style MV records with ref=4 chaining (max_ref 3), textured plane, descriptor gating on",
            "streams_per_gpu": S_PER_GPU, "frames_per_stream_per_step": F, "n_gpus": n_gpus, "distinct_clips": N_BASE,
            "l2": "inputs+outputs per step (~5.4 GB) are far larger than the 126 MB L2; no explicit flush", "max_tracks": MAX_TRACKS,
            "map_points_per_stream": "one per frame-0 track (~450), half of them as the reference keyframe's list"}


def bind_to_gpu_numa_node(local):
    """One process per GPU: run on (and therefore allocate pinned host buffers from) the NUMA node the GPU hangs off, so that
    the host->device copies of N ranks do not all read one socket's memory. Best effort: returns the node or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % dev).read())
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        if node < 0 or len(nodes) < 2:
            return "gpu %s node %d of %s: nothing to bind" % (dev, node, nodes)
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "gpu %s node %d: no allowed cpu there" % (dev, node)
        os.sched_setaffinity(0, cpus)
        return "gpu %s -> node %d (%d cpus)" % (dev, node, len(cpus))
    except Exception as e:      # containers without sysfs topology, older torch: run unbound
        return "unbound (%s)" % (e,)


# ------------------------------------------------------------------------------------------------ GPU arm ------
def run_product(args):
    import torch
    from movfe import lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    log("[rank %d] numa: %s" % (rank, bind_to_gpu_numa_node(local)))
    S = S_PER_GPU
    n_steps = args.warmup + args.steps
    LA = MAX_REF + 1
    n_frames = max(F * (n_steps + 2) + LA, REF_FRAMES)
    t0 = time.time()
    clips = make_clips(n_frames)
    log("[rank %d] generated %d clips x %d frames in %.1fs" % (rank, N_BASE, n_frames, time.time() - t0))
    cam = clips[0]["spec"].camera()

    def new_context(serial_raster=False):
        """Context + untimed setup: window 0 seeds the tracks; map points are built from the frame-0 tables."""
        ctx = lib.Context(S, W, H, max_records_per_frame=MAX_RECORDS, max_ref=MAX_REF, window_frames=F, max_tracks=MAX_TRACKS,
                          max_map_points=2048, has_grey=True, device=local, serial_raster=serial_raster)
        ctx.set_camera(cam, T.pose_params(), 0.5)
        win0 = pack_window(clips, S, 0, F + LA)
        ctx.push_frames(win0["n"], win0["recs"].numpy()[:win0["n_records"] * 40].view(T.MV_RECORD), win0["off"].numpy(),
                        win0["flags"].numpy(), win0["grey"].numpy())
        ctx.raster(0, F)
        ctx.extract(0, F)
        for b in range(N_BASE):
            sp = clips[b]["spec"]
            mp = synth.map_from_tracks(sp, ctx.tracks(b, 0), synth.pose_at(sp, 0))
            for s in range(b, S, N_BASE):
                ctx.set_map_points(s, mp, len(mp) // 2)
                ctx.set_pose(s, synth.pose_struct(synth.pose_at(sp, 0)))
        ctx.track_poses(0, F)
        ctx.synchronize()
        return ctx, torch.cuda.ExternalStream(ctx.stream_ptr, device=local)

    # ---- inputs of every step: host (pinned) and device-resident copies; one extra window feeds the pipelined push ----
    host, dev = [], []
    for k in range(n_steps + 1):
        f0 = F * (k + 1) + LA
        w = pack_window(clips, S, f0, f0 + F)
        host.append(w)
        if k < n_steps:
            dev.append({kk: (v.cuda(non_blocking=True) if hasattr(v, "cuda") else v) for kk, v in w.items()})
    torch.cuda.synchronize()
    h2d = int(host[0]["n_records"] * 40 + host[0]["off"].numel() * 8 + host[0]["flags"].numel() + host[0]["grey"].numel())
    d2h = int(np.zeros((S, F), T.POSE).nbytes + np.zeros((S, F), np.int32).nbytes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(ctx, ext, fn, label):
        """W warm-up steps then exactly K timed steps, CUDA events on the library's primary stream (the pose stream is
        joined into it by a device-side fence before the closing event), max over ranks."""
        sampler = ClockSampler(local)
        sampler.start()
        sampler.mark_load()
        for k in range(args.warmup):
            fn(k)
        ctx.profile_enable(True)
        ctx.profile_read(reset=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        with torch.cuda.stream(ext):
            e0.record()
        for k in range(args.warmup, n_steps):
            fn(k)
        ctx.fence()
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        wall = time.perf_counter() - t_wall
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        stage_ms, launches = ctx.profile_read(reset=True)
        ctx.profile_enable(False)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        log("[rank %d] %s: %.3f ms device, %.3f ms wall, stages %s, clocks %s" %
            (rank, label, ms, wall * 1e3, {k: round(v, 3) for k, v in stage_ms.items()}, clocks))
        return float(t[0]), float(t[1]), stage_ms, launches, clocks

    # ---- device-resident: inputs already in HBM when the timed region starts ------------------------------------------
    ctx, ext = new_context()

    def step_device(k):
        d = dev[k]
        first = F * (k + 1)
        ctx.push_frames_device(F, d["recs"].data_ptr(), d["off"].data_ptr(), d["n_records"], d["flags"].data_ptr(), d["grey"].data_ptr())
        ctx.raster(first, F)
        ctx.extract(first, F)
        if not os.environ.get("BENCH_NO_POSE"):     # development only: how much the pose chain costs the other streams
            ctx.track_poses(first, F)

    dev_ms, _, stage_ovl, launches, clocks = timed(ctx, ext, step_device, "device-resident")
    frames_total = world * S * F * args.steps
    value = frames_total / (dev_ms / 1e3)
    step_ms = dev_ms / args.steps
    ctx.close()
    if os.environ.get("BENCH_QUICK"):   # development sweeps: the device-resident headline region only
        if rank == 0:
            emit({"quick": True, "value": value, "ms_per_step": step_ms,
                              "stage_ms_per_step": {k: v / args.steps for k, v in stage_ovl.items()}})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the same steps once more with MOVFE_CFG_SERIAL_RASTER: the raster kernels of a window wait for the propagation of
    # the previous one instead of running beside it, so the CUDA events around grid_kernel time that kernel ALONE (beside
    # another kernel its event span measures the sharing, not the kernel). This pass feeds `roofline`; `value` does not use it.
    ctx, ext = new_context(serial_raster=True)
    serial_ms, _, stage_ms, _, _ = timed(ctx, ext, step_device, "device-resident, serial raster")

    # workload statistics for the whole-step byte count, sampled on stream 0 of the last window: T tracks per
    # frame, c = candidate hops per track (non-empty slots of the slot-grid cell under the track), hops per frame
    last_first = F * n_steps
    t_cnt, c_sum, c_n, hop_cnt = [], 0, 0, []
    for f in range(last_first + 1, last_first + F, 5):
        prev = ctx.tracks(0, f - 1)
        g = ctx.grid(0, f).reshape(H, W, 4)
        xs, ys = prev["pt_x"].astype(np.int64), prev["pt_y"].astype(np.int64)
        ok = (xs >= 0) & (ys >= 0) & (xs < W) & (ys < H) & ((prev["flags"] & T.TRACK_COVERAGE) == 0)
        c_sum += int((g[ys[ok], xs[ok]] >= 0).sum())
        c_n += int(ok.sum())
        t_cnt.append(len(prev))
        hop_cnt.append(int(ctx.raster_counts(0, f)[0]))
    T_mean, c_bar, hops_mean = float(np.mean(t_cnt)), c_sum / max(c_n, 1), float(np.mean(hop_cnt))
    P_corr = float(np.mean(ctx.poses(last_first, F)[1]))    # inliers of the last window (<= correspondences per solve)

    grid_bytes = S * F * W * H * 16.0                       # algorithmic bytes of the dominant HBM kernel per launch
    grid_ms = stage_ms["grid"] / max(args.steps, 1)
    grid_ms_ovl = stage_ovl["grid"] / max(args.steps, 1)
    peak, peak_src = peaks()
    achieved = grid_bytes / 1e9 / (grid_ms / 1e3)
    serial_step_ms = serial_ms / args.steps
    # whole-step view: algorithmic bytes of one step by SURVEY.md 8d's per-frame formulas, over the step time.
    #   B_raster = 40 M + 16 W H + 12 Hops;  B_prop = T (64 + 16 + 12 c + 64) + T (1 + c) 272 (descriptor gating on);
    #   B_match = 36 L + 8 L + 4 T + 4 T;    B_pose = I 20 P + P / 8 + 64, twice per frame
    n_rec = host[args.warmup]["n_records"]
    L_pts, I_pose = 450.0, float(T.pose_params()["iteration_count"])   # I: the iteration budget (an upper bound of the passes run)
    b_raster = 40.0 * n_rec + grid_bytes + 12.0 * hops_mean * S * F
    b_prop = S * F * (T_mean * (64 + 16 + 12 * c_bar + 64) + T_mean * (1 + c_bar) * 272)
    b_match = S * F * (44 * L_pts + 8 * T_mean)
    b_pose = S * F * 2 * (I_pose * 20 * P_corr + P_corr / 8 + 64)
    step_bytes = b_raster + b_prop + b_match + b_pose
    roofline = {"bound": "hbm", "kernel": "grid_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": grid_traffic(), "peak_source": peak_src, "algorithmic_bytes_per_launch": grid_bytes,
                "launch_ms": grid_ms, "kernel_share_of_step": grid_ms / serial_step_ms,
                "timed": "CUDA events on the raster stream around every grid_kernel launch of a second timed region (same steps, "
                         "MOVFE_CFG_SERIAL_RASTER: the kernel runs alone, as under ncu); in the headline region it runs beside the "
                         "previous window's propagation, see `overlapped`",
                "serial_ms_per_step": serial_step_ms,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "overlapped": {"launch_ms": grid_ms_ovl, "achieved": grid_bytes / 1e9 / (grid_ms_ovl / 1e3),
                               "stage_ms_per_step": {k: v / args.steps for k, v in stage_ovl.items()}},
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / 1e9 / (step_ms / 1e3),
                               "frac": step_bytes / 1e9 / (step_ms / 1e3) / peak,
                               "bytes": {"raster": b_raster, "propagation": b_prop, "match": b_match, "pose": b_pose},
                               "workload": {"tracks_per_frame": T_mean, "candidates_per_track": c_bar, "hops_per_frame": hops_mean,
                                            "pose_correspondences": P_corr, "records_per_frame": n_rec / float(S * F)},
                               "note": "headline region: raster of window k+1 and the pose chain of window k run on their own streams beside the propagation of window k"}}
    ctx.close()
    del dev
    torch.cuda.empty_cache()

    # ---- end to end: same steps through the host-buffer API, H2D + D2H inside the timed region -------------------------
    # Software-pipelined as a streaming caller would: while window k is computed, window k+1 is pushed (its host->device
    # copy runs on the library's copy stream), then the poses of window k are read back. Every step does one push of one
    # step's inputs from pinned memory and one device->host read of its result.
    ctx, ext = new_context()
    last = {}

    def push_host(k):
        w = host[k]
        ctx.push_frames(F, w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())

    push_host(0)

    def step_host(k):
        first = F * (k + 1)
        ctx.raster(first, F)
        ctx.extract(first, F)
        ctx.track_poses(first, F)
        push_host(k + 1)
        last["poses"], last["ninl"] = ctx.poses(first, F)      # device->host read of the step's result (synchronises)

    _, e2e_wall_ms, e2e_stage_ms, _, _ = timed(ctx, ext, step_host, "end-to-end")
    e2e_value = frames_total / (e2e_wall_ms / 1e3)
    med_inl = float(np.median(last["ninl"]))
    max_tracks_seen = max(ctx.track_count(s, F * n_steps + F - 1)[0] for s in range(0, S, max(S // N_BASE, 1)))
    ctx.close()

    if rank == 0:
        cores = os.cpu_count() or 1
        os.sched_setaffinity(0, all_cpus)        # the CPU baseline uses every host core again
        cpu_fps, cpu_dt = cpu_frontend_sample(clips, REF_FRAMES, cores, repeats=CPU_REPEATS)
        cfg = config_dict(world)
        cfg["tracks_in_last_table"] = int(max_tracks_seen)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "i32/f32 raster+tracks, f64 pose", "data": "synthetic", "config": cfg,
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                           "samples": clocks["samples"]},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_wall_ms / args.steps, "median_inliers_last_step": med_inl,
                        "h2d_gbs": h2d / 1e9 / (e2e_wall_ms / args.steps / 1e3),
                        "bound": "host->device copy of the step's inputs (records + full grey planes) over PCIe",
                        "pipelining": "push of window k+1 overlaps compute of window k; poses of window k read back every step"},
                "gpu_launches": int(sum(launches.values())), "roofline": roofline,
                "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d streams (one per host thread) x %d frames x %d repeats of the same C2 clips, %.1fs" %
                                           (cores, REF_FRAMES, CPU_REPEATS, cpu_dt)}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
