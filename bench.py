#!/usr/bin/env python
"""bench.py — front-end frames/s of the B200 tracking front-end (BASELINE.json metric).

Default (what the driver runs) = config C2: 640x480 mono, ref=4 MV chaining (max_ref 3), textured plane, 64 streams
batched per GPU. One "step" = one pass of the hot path over one batch: S streams x F new frames go through
ingest -> raster (hop lists, kps, slot resolution) -> track propagation -> [join(KF), pose, frustum, join(local), pose].
Every window the tracker receives a fresh local map (about 10^3 points per stream, built from the stream's own track
table of the previous window - the stand-in for keyframe insertion + UpdateLocalPoints, which are out of scope), so the
join / frustum / pose stages work on a few hundred correspondences per frame, as SURVEY.md 8a describes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4|C5|C1]     product arm (CUDA, through the C-ABI)
  python bench.py --impl reference ...                 reference arm: the CPU front-end on the host cores, SAME frames,
                                                       same map schedule (oracle port; raster + propagation pinned to the
                                                       reference's own sources by tests/test_ref_parity.py)
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

METRIC = "front_end_frames_per_s"
DTYPE = "i32/f32 raster+tracks, f64 pose"
N_BASE = 8          # distinct synthetic clips; stream s replays clip s % N_BASE (every stream is processed separately)
CPU_REPEATS = 1

# The workloads BASELINE.json names. C2 is the metric's configuration (what the driver runs and what `value` is quoted on).
CONFIGS = {
    "C2": dict(name="C2: 640x480 mono, x264-style MV records with ref=4 chaining (max_ref 3), textured plane, descriptor gating on",
               W=640, H=480, S=64, F=16, max_ref=3, refs=4, stereo=False, dense=False, grey=True, max_records=4800, max_tracks=8192,
               fx=320.0, map_n=1024),
    "C3": dict(name="C3: EuRoC-shaped 752x480 stereo frame-packed (left/right alternate, only left frames carry MVs, ref=2), descriptor gating on",
               W=752, H=480, S=128, F=16, max_ref=1, refs=2, stereo=True, dense=False, grey=True, max_records=5640, max_tracks=8192,
               fx=458.654, map_n=1024),
    "C4": dict(name="C4: 1920x1080 dense 4x4-partition MV fields (129 600 records per frame, ref=0), MV-only propagation of seeded 16x16 tracks",
               W=1920, H=1080, S=32, F=8, max_ref=0, refs=1, stereo=False, dense=True, grey=False, max_records=129600, max_tracks=8192,
               fx=960.0, map_n=1024),
}
F_ENV = os.environ.get("BENCH_F")


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


NOMINAL_HBM_GBS = 8000.0    # north_star states the >= 60 % target against B200's ~8 TB/s


# ------------------------------------------------------------------------------------------------ workload -----
def make_clips(cfg, n_frames, n_base=N_BASE):
    """Seeded synthetic clips. Generation (the warped grey planes) takes tens of seconds, so the arrays are cached in the
    temp directory: the driver's two arms and the profiling passes of one box reuse them."""
    import hashlib
    import pickle
    key = hashlib.sha1(repr((sorted((k, v) for k, v in cfg.items() if k not in ("S", "F", "name", "map_n")), n_frames, n_base, 3)).encode()).hexdigest()[:16]
    cache = os.path.join(tempfile.gettempdir(), "movfe_clips_%s.pkl" % key)
    if os.path.exists(cache):
        try:
            with open(cache, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
    clips = _make_clips(cfg, n_frames, n_base)
    try:
        tmp = cache + ".%d" % os.getpid()
        with open(tmp, "wb") as f:
            pickle.dump(clips, f, protocol=4)
        os.replace(tmp, cache)
    except Exception:
        pass
    return clips


def _make_clips(cfg, n_frames, n_base):
    clips = []
    for b in range(n_base):
        spec = synth.Spec(cfg["W"], cfg["H"], n_frames=n_frames, refs=cfg["refs"], seed=0x5EED0002 + 977 * b, phase=0.37 * b,
                          fx=cfg["fx"], fy=cfg["fx"], stereo=cfg["stereo"], dense4x4=cfg["dense"], start_p=cfg["dense"])
        recs, off, flags = synth.make_records(spec)
        grey = synth.make_grey(spec) if cfg["grey"] else None
        clips.append(dict(spec=spec, recs=recs, off=off, flags=flags, grey=grey))
    return clips


def pack_window(cfg, clips, S, f0, f1, pinned=True):
    """Stream-major packed inputs for frames [f0,f1) of S streams (stream s replays clip s % len(clips))."""
    import torch
    W, H = cfg["W"], cfg["H"]
    per = []
    for c in clips:
        r0, r1 = c["off"][f0], c["off"][f1]
        per.append((c["recs"][r0:r1], c["off"][f0:f1 + 1] - r0, c["flags"][f0:f1]))
    n = f1 - f0
    tot = sum(len(per[s % len(per)][0]) for s in range(S))
    recs = torch.empty(max(tot, 1) * 40 + 16, dtype=torch.uint8, pin_memory=pinned)
    off = torch.empty(S * n + 1, dtype=torch.int64, pin_memory=pinned)
    flags = torch.empty(S * n, dtype=torch.uint8, pin_memory=pinned)
    grey = None
    if clips[0]["grey"] is not None:
        grey = torch.empty((S, n, H, W), dtype=torch.uint8, pin_memory=pinned)
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
    out = dict(recs=recs, off=off, flags=flags, grey=grey, n_records=tot, n=n)
    if pinned:   # the form the decoder shim hands over: 16-byte records, packed while the side data is copied out of the AVFrame
        from movfe import lib
        recs16 = torch.empty(max(tot, 1) * 16, dtype=torch.uint8, pin_memory=True)
        lib.pack_records(rv, recs16.numpy()[:tot * 16].view(T.PACKED_RECORD))
        out["recs16"] = recs16
    return out


def local_map(cfg, spec, table, frame):
    """The local map handed to the tracker before frame `frame`+1: one point per track among the `map_n` OLDEST tracks of
    `table` (the table of `frame`; tables are ordered by age), back-projected through that frame's ground-truth pose.
    Half of the points play the reference keyframe's list. Identical in both arms: `table` is bit-identical."""
    t = table[:cfg["map_n"]]
    mp = synth.map_from_tracks(spec, t, synth.pose_at(spec, frame))
    return mp, len(mp) // 2


def pack_maps(maps, S):
    """[(mp, n_kf)] per distinct clip -> packed arrays for S streams (movfe_set_map_points_batch)."""
    pts = np.concatenate([maps[s % len(maps)][0] for s in range(S)]) if S else np.zeros(0, T.MAP_POINT)
    off = np.cumsum([0] + [len(maps[s % len(maps)][0]) for s in range(S)]).astype(np.int64)
    nkf = np.array([maps[s % len(maps)][1] for s in range(S)], np.int32)
    return np.ascontiguousarray(pts, T.MAP_POINT), off, nkf


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


def frame_plan(cfg, args):
    """Frames of one run: window 0 seeds the state, W warm-up windows, K timed windows. Both arms time frames
    [timed_from, n_proc)."""
    F, LA = cfg["F"], cfg["max_ref"] + 1
    n_steps = args.warmup + args.steps
    n_proc = F * (n_steps + 1)
    return dict(F=F, LA=LA, n_steps=n_steps, n_proc=n_proc, n_clip=n_proc + F + LA, timed_from=F * (args.warmup + 1))


def config_dict(cfg, n_gpus, plan, mode=None):
    d = {"workload": cfg["name"], "streams_per_gpu": cfg["S"], "frames_per_stream_per_step": plan["F"], "n_gpus": n_gpus,
         "distinct_clips": N_BASE, "max_tracks": cfg["max_tracks"],
         "timed_frames": "[%d, %d) of every stream" % (plan["timed_from"], plan["n_proc"]),
         "l2": "inputs+outputs of a step are far larger than the 126 MB L2; no explicit flush",
         "local_map": "refreshed every window: the %d oldest tracks of the previous window's last table, back-projected through the "
                      "ground-truth pose; half of them as the reference keyframe's list" % cfg["map_n"]}
    if mode:
        d["raster_mode"] = mode
    return d


# ------------------------------------------------------------------------------------------------ CPU arm ------
def oracle_tables_for_schedule(cfg, clips, plan):
    """Map schedule of every distinct clip from the ORACLE's track tables (the reference arm has no GPU). One pass over the
    clip per clip, tables kept at the window ends."""
    from oracle import pyoracle as orc
    W, H, F = cfg["W"], cfg["H"], plan["F"]
    out = []
    flat = np.full((H, W), 128, np.uint8)

    def one(c):
        sp = c["spec"]
        nf = plan["n_proc"] + plan["LA"]
        clip = orc.Clip(W, H, c["recs"][:c["off"][nf]], c["off"][:nf + 1], c["flags"][:nf], cfg["max_ref"])
        prev = synth.seed_tracks_lattice(sp) if cfg["dense"] else np.zeros(0, T.TRACK)
        cid = int(prev["track_id"].max()) if len(prev) else 0
        tabs = {}
        for f in range(plan["n_proc"]):
            img = c["grey"][f] if c["grey"] is not None else flat
            prev, _, cid, _ = orc.extract_frame(W, H, c["flags"][f], img, clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev,
                                                cid, max_tracks=cfg["max_tracks"])
            if f == 0 or (f + 1) % F == 0:
                tabs[f] = prev
        return tabs

    res = [None] * len(clips)
    ths = [threading.Thread(target=lambda i=i: res.__setitem__(i, one(clips[i]))) for i in range(len(clips))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for c, tabs in zip(clips, res):
        out.append(schedule_from_tables(cfg, c["spec"], tabs, plan))
    return out


def schedule_from_tables(cfg, spec, tabs, plan):
    """tabs: {frame: table} at frame 0 and at the last frame of every window -> (initial map, [(frame, map, n_kf), ...])."""
    F = plan["F"]
    init = local_map(cfg, spec, tabs[0], 0)
    sched = []
    for k in range(1, plan["n_steps"] + 1):
        mp, nkf = local_map(cfg, spec, tabs[F * k - 1], F * k - 1)
        sched.append((F * k, mp, nkf))
    return init, sched


def cpu_frontend_sample(cfg, clips, scheds, plan, threads, repeats=1):
    """Times the oracle's whole front-end (raster -> extract -> joins/frustum -> pose x2) on `threads` host threads, one
    stream per thread: every thread runs its stream from frame 0 (the state has to be built) and the clock covers frames
    [timed_from, end) - the frames the GPU arm times - from the first thread entering them to the last thread leaving.
    Returns (frames/s, seconds, per-clip results of the first `len(clips)` threads)."""
    from oracle import pyoracle as orc
    orc.lib()
    W, H = cfg["W"], cfg["H"]
    cam = clips[0]["spec"].camera()
    pp = T.pose_params()
    nf = plan["n_proc"] + plan["LA"]     # the look-ahead frames the GPU arm has pushed when it rasterises the last window
    res = [None] * threads

    def run(i):
        c = clips[i % len(clips)]
        (mp0, nkf0), sched = scheds[i % len(clips)]
        sp = c["spec"]
        seeds = synth.seed_tracks_lattice(sp) if cfg["dense"] else None
        for _ in range(repeats):
            res[i] = orc.frontend_run(W, H, c["recs"][:c["off"][nf]], c["off"][:nf + 1], c["flags"][:nf],
                                      None if c["grey"] is None else c["grey"][:nf], seeds, mp0, synth.pose_struct(synth.pose_at(sp, 0)), cam, pp,
                                      max_ref=cfg["max_ref"], max_tracks=cfg["max_tracks"], n_kf_points=nkf0, map_schedule=sched,
                                      timed_from=plan["timed_from"])

    ths = [threading.Thread(target=run, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    t0 = min(r["tail_times"][0] for r in res)
    t1 = max(r["tail_times"][1] for r in res)
    n_timed = nf - plan["timed_from"]
    return threads * n_timed / (t1 - t0), t1 - t0, res


def run_reference(args, cfg):
    """Reference arm: the reference's CPU front-end on all host cores, same config / metric / unit / frames / map schedule."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref_args = argparse.Namespace(**vars(args))
    if os.environ.get("BENCH_REF_FRAMES"):   # shortened sample (CPU test of the contract)
        ref_args.warmup, ref_args.steps = 0, 1
    plan = frame_plan(cfg, ref_args)
    clips = make_clips(cfg, plan["n_clip"], n_base=min(N_BASE, cores))
    t0 = time.time()
    scheds = oracle_tables_for_schedule(cfg, clips, plan)
    log("map schedule from the oracle's tables: %.1fs" % (time.time() - t0))
    vals = []
    for i in range(max(1, min(args.steps, 3))):      # each step = one bounded sample of the timed frames on every core
        fps, dt, _ = cpu_frontend_sample(cfg, clips, scheds, plan, cores)
        vals.append((fps, dt))
        log("reference sample %d: %.1f frames/s (%.2fs)" % (i, fps, dt))
    fps = float(np.mean([v[0] for v in vals]))
    sample = "%d streams (one per host thread), frames [%d, %d) of the %s clips timed, %d samples" % (
        cores, plan["timed_from"], plan["n_proc"] + plan["LA"], args.config, len(vals))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([v[1] for v in vals])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(cfg, args.gpus, frame_plan(cfg, args)),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                             "flags": "-O3 -march=x86-64-v3 -ffp-contract=off (oracle/Makefile; the reference builds -O3 -march=native)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/kernel_traffic_r2.json), None if absent."""
    p = os.path.join(ROOT, "profiles", "kernel_traffic_r2.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel)


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
def run_product(args, cfg):
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
    W, H, S = cfg["W"], cfg["H"], cfg["S"]
    plan = frame_plan(cfg, args)
    F, LA, n_steps = plan["F"], plan["LA"], plan["n_steps"]
    t0 = time.time()
    clips = make_clips(cfg, plan["n_clip"])
    log("[rank %d] generated %d clips x %d frames in %.1fs" % (rank, N_BASE, plan["n_clip"], time.time() - t0))
    cam = clips[0]["spec"].camera()
    MAPCAP = cfg["map_n"]
    fused = not os.environ.get("BENCH_GRID_MODE")     # default: slots resolved from the tile queues, no slot grid in HBM

    def new_context(n_streams=S, serial_raster=False, grid_mode=None):
        gm = (not fused) if grid_mode is None else grid_mode
        return lib.Context(n_streams, W, H, max_records_per_frame=cfg["max_records"], max_ref=cfg["max_ref"], window_frames=F,
                           max_tracks=cfg["max_tracks"], max_map_points=MAPCAP, has_grey=cfg["grey"], device=local,
                           serial_raster=serial_raster, output_grid=gm)

    def seed(ctx, n_streams):
        if cfg["dense"]:
            for s in range(n_streams):
                sd = synth.seed_tracks_lattice(clips[s % N_BASE]["spec"])
                ctx.set_tracks(s, sd, int(sd["track_id"].max()))

    def push_np(ctx, w):
        ctx.push_frames(w["n"], w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(),
                        None if w["grey"] is None else w["grey"].numpy())

    # ---- untimed pre-pass over the N_BASE distinct clips: the track tables at the window ends give every window's local map
    # (the mapping side is out of scope; its product is an input of the front-end, computed the same way in both arms) ------
    t0 = time.time()
    pre = new_context(n_streams=N_BASE)
    seed(pre, N_BASE)
    tabs = [dict() for _ in range(N_BASE)]
    w = pack_window(cfg, clips, N_BASE, 0, F + LA, pinned=False)
    push_np(pre, w)
    for k in range(n_steps + 1):
        if k > 0:
            f0 = F * k + LA
            push_np(pre, pack_window(cfg, clips, N_BASE, f0, f0 + F, pinned=False))
        pre.raster(F * k, F)
        pre.extract(F * k, F)
        for b in range(N_BASE):
            if k == 0:
                tabs[b][0] = pre.tracks(b, 0)
            tabs[b][F * (k + 1) - 1] = pre.tracks(b, F * (k + 1) - 1)
    pre.close()
    scheds = [schedule_from_tables(cfg, clips[b]["spec"], tabs[b], plan) for b in range(N_BASE)]
    log("[rank %d] map schedule from a GPU pre-pass: %.1fs, %d points per stream" % (rank, time.time() - t0, len(scheds[0][0][0])))

    def setup(ctx):
        """window 0 (untimed): seeds the tracks, installs the initial maps and poses"""
        ctx.set_camera(cam, T.pose_params(), 0.5)
        seed(ctx, S)
        push_np(ctx, pack_window(cfg, clips, S, 0, F + LA))
        ctx.raster(0, F)
        ctx.extract(0, F)
        pts, off, nkf = pack_maps([scheds[b][0] for b in range(N_BASE)], S)
        ctx.set_map_points_batch(pts, off, nkf, MAPCAP)
        for s in range(S):
            ctx.set_pose(s, synth.pose_struct(synth.pose_at(clips[s % N_BASE]["spec"], 0)))
        ctx.track_poses(0, F)
        ctx.synchronize()
        return torch.cuda.ExternalStream(ctx.stream_ptr, device=local)

    # ---- inputs of every step: host (pinned) and device-resident copies; one extra window feeds the pipelined push ----
    host, dev = [], []
    for k in range(n_steps + 1):
        f0 = F * (k + 1) + LA
        w = pack_window(cfg, clips, S, f0, f0 + F)
        if k < n_steps:
            pts, off, nkf = pack_maps([scheds[b][1][k][1:] for b in range(N_BASE)], S)
            w["map"] = {"pts": torch.from_numpy(pts.view(np.uint8).copy()).pin_memory(), "off": torch.from_numpy(off).pin_memory(),
                        "nkf": torch.from_numpy(nkf).pin_memory()}
        host.append(w)
        if k < n_steps:
            d = {kk: (v.cuda(non_blocking=True) if hasattr(v, "cuda") else v) for kk, v in w.items() if kk != "map"}
            d["map"] = {kk: v.cuda(non_blocking=True) for kk, v in w["map"].items()}
            dev.append(d)
    torch.cuda.synchronize()
    h0 = host[0]
    map_bytes = int(h0["map"]["pts"].numel() + h0["map"]["off"].numel() * 8 + h0["map"]["nkf"].numel() * 4)
    h2d = int(h0["n_records"] * (40 if os.environ.get("BENCH_E2E_RECORDS40") else 16) + h0["off"].numel() * 8 + h0["flags"].numel() + (h0["grey"].numel() if h0["grey"] is not None else 0) + map_bytes)
    d2h = int(np.zeros((S, F), T.POSE).nbytes + np.zeros((S, F), np.int32).nbytes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(ctx, ext, fn, label):
        """W warm-up steps then exactly K timed steps, CUDA events on the library's primary stream (the raster and pose
        streams are joined into it by a device-side fence before the closing event), max over ranks."""
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

    def make_step_device(ctx):
        def step_device(k):
            d = dev[k]
            first = F * (k + 1)
            ctx.push_frames_device(F, d["recs"].data_ptr(), d["off"].data_ptr(), d["n_records"], d["flags"].data_ptr(),
                                   None if d["grey"] is None else d["grey"].data_ptr())
            ctx.raster(first, F)
            ctx.extract(first, F)
            if not os.environ.get("BENCH_NO_POSE"):     # development only: how much the pose chain costs the other streams
                ctx.set_map_points_batch(d["map"]["pts"].data_ptr(), d["map"]["off"].data_ptr(), d["map"]["nkf"].data_ptr(), MAPCAP, on_device=True)
                ctx.track_poses(first, F)
        return step_device

    # ---- device-resident: inputs already in HBM when the timed region starts ------------------------------------------
    ctx = new_context()
    ext = setup(ctx)
    dev_ms, _, stage_ovl, launches, clocks = timed(ctx, ext, make_step_device(ctx), "device-resident")
    frames_total = world * S * F * args.steps
    value = frames_total / (dev_ms / 1e3)
    step_ms = dev_ms / args.steps
    if os.environ.get("BENCH_QUICK"):   # development sweeps: the device-resident headline region only
        if rank == 0:
            emit({"quick": True, "value": value, "ms_per_step": step_ms, "stage_ms_per_step": {k: v / args.steps for k, v in stage_ovl.items()}})
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- workload statistics + the GPU's tables and poses of the last window, for the byte counts and the parity check ----
    last_first = F * n_steps
    gpu_last = {}
    poses_last, ninl_last = ctx.poses(last_first, F)
    for b in range(N_BASE):
        gpu_last[b] = [ctx.tracks(b, f) for f in range(last_first, last_first + F)]
    T_mean = float(np.mean([len(t) for t in gpu_last[0]]))
    n_inl_mean = float(np.mean(ninl_last))
    med_inl = float(np.median(ninl_last))
    stats_ctx = ctx.workload_stats()        # correspondences per solve, candidates per track, hops per frame (device counters)
    ctx.close()

    # ---- the same steps once more, every stage alone on the GPU (MOVFE_CFG_SERIAL_RASTER: raster waits for propagation and
    # pose, so each stage's CUDA-event span times that stage alone, as under ncu). Feeds `roofline`; `value` does not use it.
    ctx = new_context(serial_raster=True)
    ext = setup(ctx)
    serial_ms, _, stage_ms, _, _ = timed(ctx, ext, make_step_device(ctx), "device-resident, stages serialised")
    ctx.close()
    grid_stage = None
    if fused:   # the slot-grid kernel as a sub-record: the same steps in grid-output (parity) mode
        ctx = new_context(serial_raster=True, grid_mode=True)
        ext = setup(ctx)
        _, _, grid_stage, _, _ = timed(ctx, ext, make_step_device(ctx), "device-resident, grid-output mode, stages serialised")
        ctx.close()
    del dev
    torch.cuda.empty_cache()

    peak, peak_src = peaks()
    n_rec = host[args.warmup]["n_records"]
    c_bar, hops_mean, P_corr, L_pts = stats_ctx["candidates_per_track"], stats_ctx["hops_per_frame"], stats_ctx["pose_correspondences"], float(MAPCAP)
    I_pose = stats_ctx["pose_passes_per_solve"]
    # ALGORITHMIC bytes of one step by SURVEY.md 8d's per-frame formulas:
    #   B_raster = 40 M + 12 Hops (+ 16 W H only when the slot grid is an OUTPUT: parity mode);
    #   B_prop = T (64 + 16 + 12 c + 64) + G T (1 + c) 272;  B_match = 36 L + 8 L + 4 T + 4 T;  B_pose = I 20 P + P/8 + 64, twice
    G = 1.0 if cfg["grey"] else 0.0
    grid_bytes = S * F * W * H * 16.0
    b_raster = 40.0 * n_rec + 12.0 * hops_mean * S * F + (0.0 if fused else grid_bytes)
    b_prop = S * F * (T_mean * (64 + 16 + 12 * c_bar + 64) + G * T_mean * (1 + c_bar) * 272)
    b_match = S * F * (44 * L_pts + 8 * T_mean)
    b_pose = S * F * 2 * (I_pose * 20 * P_corr + P_corr / 8 + 64)
    step_bytes = b_raster + b_prop + b_match + b_pose
    per = {k: v / args.steps for k, v in stage_ms.items()}
    stage_bytes = {"ingest": 40.0 * n_rec + (S * F * W * H if cfg["grey"] else 0), "hops": 40.0 * n_rec * (1 + LA / F) + 12.0 * hops_mean * S * F,
                   "grid": (0.0 if fused else grid_bytes), "extract": b_prop, "pose": b_match + b_pose}
    stages = {k: {"ms_per_step": per[k], "algorithmic_bytes": stage_bytes[k], "gbs": (stage_bytes[k] / 1e9 / (per[k] / 1e3)) if per[k] > 0 else None,
                  "frac": (stage_bytes[k] / 1e9 / (per[k] / 1e3) / peak) if per[k] > 0 else None} for k in per}
    # hop lists + slot resolution are ONE row of SURVEY 8d ("raster": 40 M + 12 Hops, + 16 W H when the grid is an output): they are
    # compared with the other stages as one stage, so that a fused-mode run (grid bytes not charged) never reports a stage without bytes
    per["raster"] = per["hops"] + per["grid"]
    stage_bytes["raster"] = b_raster
    stages["raster"] = {"ms_per_step": per["raster"], "algorithmic_bytes": b_raster, "gbs": b_raster / 1e9 / (per["raster"] / 1e3),
                        "frac": b_raster / 1e9 / (per["raster"] / 1e3) / peak, "what": "hops + grid stages together (SURVEY 8d's raster row)"}
    dom = max(("ingest", "raster", "extract", "pose"), key=lambda k: per[k])
    dom_kernel = {"extract": "track propagation chain (cand_lane_kernel + birth_lane_kernel + finalize_kernel per frame)",
                  "pose": "pose chain (tp_prep_kernel + track_poses2_kernel)",
                  "raster": "raster (count/emit/bbox hop-list kernels + grid_kernel: %s)" % ("per-tile cell tables, fused mode" if fused else "slot grid written"),
                  "ingest": "ingest kernels"}[dom]
    whole = step_bytes / 1e9 / (step_ms / 1e3)
    roofline = {"bound": "hbm", "kernel": dom_kernel, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac"],
                "frac_of_nominal_8000": stages[dom]["gbs"] / NOMINAL_HBM_GBS, "traffic": ncu_traffic(dom) if cfg["name"].startswith("C2") else None, "peak_source": peak_src,
                "algorithmic_bytes_per_step": stage_bytes[dom], "stage_ms_per_step": per[dom],
                "note": "dominant stage by time of the serialised region; its SURVEY 8d bytes are mostly L1/L2 hits (candidate patches overlap), "
                        "so this stage is bound by SM time, not by HBM: see DESIGN.md",
                "whole_step": {"algorithmic_bytes": step_bytes, "achieved": whole, "frac": whole / peak, "frac_of_nominal_8000": whole / NOMINAL_HBM_GBS,
                               "ms_per_step": step_ms, "roofline_frames_per_s": peak * 1e9 / (step_bytes / (S * F)),
                               "raster_term": "40 M + 12 Hops (fused mode: the slot grid is not an output)" if fused else "40 M + 16 W H + 12 Hops (grid output)",
                               "bytes": {"raster": b_raster, "propagation": b_prop, "match": b_match, "pose": b_pose},
                               "workload": {"tracks_per_frame": T_mean, "candidates_per_track": c_bar, "hops_per_frame": hops_mean,
                                            "pose_correspondences": P_corr, "pose_passes_per_solve": I_pose, "map_points": L_pts,
                                            "records_per_frame": n_rec / float(S * F), "mean_inliers": n_inl_mean}},
                "timed": "CUDA events around every stage in a second timed region of the same steps with the stages serialised (each alone on the "
                         "GPU, as under ncu); `overlapped` = the headline region, where raster, propagation and pose run on their own streams",
                "serial_ms_per_step": serial_ms / args.steps, "stages": stages,
                "overlapped": {"stage_ms_per_step": {k: v / args.steps for k, v in stage_ovl.items()}}}
    if grid_stage is not None:
        gms = grid_stage["grid"] / args.steps
        roofline["grid_kernel"] = {"note": "slot-grid kernel of the grid-output (parity) mode, alone on the GPU; not part of the headline step",
                                   "launch_ms": gms, "algorithmic_bytes_per_launch": grid_bytes, "achieved": grid_bytes / 1e9 / (gms / 1e3),
                                   "frac": grid_bytes / 1e9 / (gms / 1e3) / peak, "traffic": ncu_traffic("grid") if cfg["name"].startswith("C2") else None}

    # ---- end to end: same steps through the host-buffer API, H2D + D2H inside the timed region -------------------------
    # Software-pipelined as a streaming caller would: while window k is computed, window k+1 is pushed (its host->device
    # copy runs on the library's copy stream), then the poses of window k are read back. Every step does one push of one
    # step's inputs from pinned memory (records, grey planes, the window's local maps) and one device->host read of its result.
    ctx = new_context()
    ext = setup(ctx)
    last = {}

    packed_push = not os.environ.get("BENCH_E2E_RECORDS40")      # development: push the 40-byte records instead

    def push_host(k):
        w = host[k]
        if packed_push:
            ctx.push_frames_packed(w["n"], w["recs16"].numpy()[:w["n_records"] * 16].view(T.PACKED_RECORD), w["off"].numpy(), w["flags"].numpy(),
                                   None if w["grey"] is None else w["grey"].numpy())
        else:
            push_np(ctx, w)

    def map_host(k):
        m = host[k]["map"]
        ctx.set_map_points_batch(m["pts"].numpy().view(T.MAP_POINT), m["off"].numpy(), m["nkf"].numpy(), MAPCAP)

    push_host(0)
    map_host(0)

    copy_only = bool(os.environ.get("BENCH_E2E_COPY_ONLY"))    # development: the pushes alone (what the push path does without kernels beside it)

    def step_host(k):
        first = F * (k + 1)
        if copy_only:
            if k + 1 < n_steps:
                map_host(k + 1)
            push_host(k + 1)
            if k % 2:
                ctx.synchronize()
            return
        ctx.raster(first, F)
        ctx.extract(first, F)
        ctx.track_poses(first, F)
        # the next window's inputs: its local maps first (a small copy that must not queue behind the big one: the pose chain
        # of window k+1 waits for it), then records + grey planes; both are one step's inputs, counted in h2d_bytes_per_step.
        # (Pushing TWO windows ahead, so that the link never waits for the host to come back from the pose read: 7.12 ms per step
        # against 7.09 - the host is not what the copy waits for. The pushes alone, BENCH_E2E_COPY_ONLY=1, run at 6.48 ms per step
        # = 54.4 GB/s of the link's 55.5; beside the kernels the same copies reach 49.5-49.8 GB/s.)
        if k + 1 < n_steps:
            map_host(k + 1)
        push_host(k + 1)
        last["poses"], last["ninl"] = ctx.poses(first, F)      # device->host read of the step's result (synchronises)

    _, e2e_wall_ms, e2e_stage_ms, _, _ = timed(ctx, ext, step_host, "end-to-end")
    e2e_value = frames_total / (e2e_wall_ms / 1e3)
    if copy_only:
        log("[rank %d] BENCH_E2E_COPY_ONLY: %.3f ms per step = %.2f GB/s" % (rank, e2e_wall_ms / args.steps, h2d / 1e9 / (e2e_wall_ms / args.steps / 1e3)))
        last["ninl"] = ninl_last
    e2e_same = bool(np.array_equal(last["ninl"], ninl_last))
    ctx.close()

    if rank == 0:
        cores = os.cpu_count() or 1
        os.sched_setaffinity(0, all_cpus)        # the CPU baseline uses every host core again
        cpu_fps, cpu_dt, cpu_res = cpu_frontend_sample(cfg, clips, scheds, plan, max(cores, N_BASE), repeats=CPU_REPEATS)
        # ---- in-run parity: the GPU's last-window tables (bytes, via the checksum) and poses against the CPU front-end on the
        # SAME frames of the same streams -----------------------------------------------------------------------------------
        from oracle import pyoracle as orc
        tracks_ok, pose_rel, inl_diff, n_checked = True, 0.0, 0, 0
        for b in range(N_BASE):
            r = cpu_res[b]
            for i, f in enumerate(range(last_first, last_first + F)):
                if orc.table_checksum(gpu_last[b][i]) != int(r["track_hash"][f]) or len(gpu_last[b][i]) != int(r["n_tracks"][f]):
                    tracks_ok = False
                n_checked += 1
                inl_diff += int(ninl_last[b, i] != r["n_inliers"][f])
                for name in ("R", "t"):
                    ref = r["poses"][f][name]
                    pose_rel = max(pose_rel, float(np.max(np.abs(poses_last[b, i][name] - ref)) / max(1.0, float(np.max(np.abs(ref))))))
        parity = {"tracks": "bit-exact" if tracks_ok else "MISMATCH", "pose_rel": pose_rel, "pose_tolerance": 1e-5,
                  "inlier_count_mismatches": inl_diff, "tables_checked": n_checked,
                  "what": "last timed window (%d frames) of %d distinct streams: GPU track tables against the CPU front-end's checksums, "
                          "poses relative to the CPU front-end's (oracle; raster + propagation pinned to the reference's sources)" % (F, N_BASE),
                  "e2e_region_same_inliers": e2e_same}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": config_dict(cfg, world, plan),
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                           "samples": clocks["samples"]},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_wall_ms / args.steps, "median_inliers_last_step": med_inl,
                        "h2d_gbs": h2d / 1e9 / (e2e_wall_ms / args.steps / 1e3), "host_read_gbs_aggregate": world * h2d / 1e9 / (e2e_wall_ms / args.steps / 1e3),
                        "bound": "host->device copy of the step's inputs (16-byte packed records + full grey planes + local maps) over PCIe",
                        "records": "movfe_push_frames_packed: 16-byte records (movfe_pack_records, done by the decoder shim while it copies the side data)" if packed_push else "movfe_push_frames: 40-byte records",
                        "pipelining": "push of window k+1 overlaps compute of window k; poses of window k read back every step"},
                "gpu_launches": int(sum(launches.values())), "roofline": roofline, "parity_check": parity,
                "raster_mode": "fused (slots resolved from per-tile hop queues; no slot grid in HBM)" if fused else "grid output",
                "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d streams (one per host thread), frames [%d, %d) of the same clips timed, %.1fs" %
                                           (max(cores, N_BASE), plan["timed_from"], plan["n_proc"] + LA, cpu_dt),
                                 "flags": "-O3 -march=x86-64-v3 -ffp-contract=off (oracle/Makefile; the reference builds -O3 -march=native)"}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    claim_stdout()
    if args.config in ("C1", "C5"):
        import bench_extra
        bench_extra.bench._JSON_FD = _JSON_FD      # bench_extra imports this file as module `bench`: same stdout duplicate
        return bench_extra.run(args)
    cfg = dict(CONFIGS[args.config])
    if F_ENV:
        cfg["F"] = int(F_ENV)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_product(args, cfg)


if __name__ == "__main__":
    main()
