"""Quick device-timing probe of the raster stages at the C2 shape (dev tool; bench.py is the contract)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from movfe import lib, synth, types as T
from gpu_util import pack_streams

S = int(os.environ.get("S", 64)); F = int(os.environ.get("F", 16)); W, H = 640, 480; K = 3
NF = F * 3 + K + 1
t0 = time.time()
base = [synth.make_records(synth.Spec(W, H, n_frames=NF, refs=4, seed=0x5EED0100 + s, phase=0.1 * s)) for s in range(min(S, 8))]
streams = [base[s % len(base)] for s in range(S)]
print("gen %.1fs" % (time.time() - t0), flush=True)
ctx = lib.Context(S, W, H, max_records_per_frame=4800, max_ref=K, window_frames=F, has_grey=False)
ctx.profile_enable(True)
pushed = 0
for step in range(3):
    first = step * F
    want = min(NF, first + F + K + 1)
    r, o, fl = pack_streams(streams, NF, pushed, want)
    ctx.push_frames(want - pushed, r, o, fl); pushed = want
    ctx.raster(first, F)
    ms, ln = ctx.profile_read()
    nrec = len(r)
    gbytes = S * F * W * H * 16 / 1e9
    print(json.dumps(dict(step=step, ms=ms, launches=ln, grid_GBps=gbytes / (ms["grid"] / 1e3), frames=S * F,
                          frames_per_s=S * F / (sum(ms.values()) / 1e3))), flush=True)
