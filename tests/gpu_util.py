"""Helpers shared by the GPU parity tests: run the CUDA path through the C-ABI over a whole clip, window by
window, exactly as a streaming caller would."""
import numpy as np

from movfe import lib, types as T


def pack_streams(per_stream, n_frames, f0, f1):
    """per_stream: list of (recs, rec_off, flags) covering n_frames. -> packed arrays for frames [f0,f1)."""
    recs, off, flags = [], [0], []
    for r, o, fl in per_stream:
        for f in range(f0, f1):
            seg = r[o[f]:o[f + 1]]
            recs.append(seg)
            off.append(off[-1] + len(seg))
            flags.append(fl[f])
    recs = np.concatenate(recs) if recs else np.zeros(0, T.MV_RECORD)
    return recs, np.array(off, np.int64), np.array(flags, np.uint8)


def run_raster_clip(per_stream, W, H, n_frames, window, max_ref, max_records=4800, grey=None, collect=True, ctx=None,
                    packed=False, **ctx_kw):
    """Pushes and rasterises the clip window by window. Returns {(s,f): dict(grid,hops,kps,cov)} and the ctx."""
    S = len(per_stream)
    own = ctx is None
    if own:
        ctx = lib.Context(S, W, H, max_records_per_frame=max_records, max_ref=max_ref, window_frames=window,
                          has_grey=grey is not None, **ctx_kw)
    LA = max_ref + 1
    out = {}
    pushed = 0
    first = 0
    while first < n_frames:
        n_out = min(window, n_frames - first)
        want = min(n_frames, first + n_out + LA)
        if want > pushed:
            r, o, fl = pack_streams(per_stream, n_frames, pushed, want)
            g = None
            if grey is not None:
                g = np.stack([grey[s][pushed:want] for s in range(S)])
            if packed:      # 16-byte records, packed on the host (movfe_pack_records)
                ctx.push_frames_packed(want - pushed, lib.pack_records(r), o, fl, g)
            else:
                ctx.push_frames(want - pushed, r, o, fl, g)
            pushed = want
        ctx.raster(first, n_out)
        if collect:
            for s in range(S):
                for f in range(first, first + n_out):
                    nh, nk, cov = ctx.raster_counts(s, f)
                    out[(s, f)] = dict(grid=ctx.grid(s, f), hops=ctx.hops(s, f), kps=ctx.kps(s, f), cov=cov)
        yield_point = (first, n_out)
        first += n_out
    return out, ctx


def assert_raster_equal(orc_clip, got, s, f):
    g = got[(s, f)]
    eh, ek = orc_clip.hops(f), orc_clip.kps(f)
    assert len(g["hops"]) == len(eh), (s, f, len(g["hops"]), len(eh))
    assert len(g["kps"]) == len(ek), (s, f, len(g["kps"]), len(ek))
    assert g["hops"].tobytes() == eh.tobytes(), (s, f, "hops differ")
    assert g["kps"].tobytes() == ek.tobytes(), (s, f, "kps differ")
    assert g["cov"] == orc_clip.coverage(f), (s, f, g["cov"], orc_clip.coverage(f))
    eg = orc_clip.grid(f)
    if not np.array_equal(g["grid"], eg):
        bad = np.argwhere((g["grid"] != eg).any(-1))
        y, x = bad[0]
        raise AssertionError("grid differs at stream %d frame %d: %d px, first (y=%d,x=%d) got %s want %s" %
                             (s, f, len(bad), y, x, g["grid"][y, x], eg[y, x]))


def run_frontend_clip(per_stream, W, H, n_frames, window, max_ref, grey=None, seeds=None, max_records=4800,
                      max_tracks=4096, threshold=25, coverage_threshold=0.20, poses=False, ctx_hook=None, grey_stride=0):
    """Raster + extract (+ pose tracking) window by window. Returns ({(s,f): tracks}, extra, ctx)."""
    S = len(per_stream)
    ctx = lib.Context(S, W, H, max_records_per_frame=max_records, max_ref=max_ref, window_frames=window,
                      has_grey=grey is not None, max_tracks=max_tracks, express_threshold=threshold,
                      coverage_threshold=coverage_threshold)
    if seeds is not None:
        for s in range(S):
            ctx.set_tracks(s, seeds[s], int(seeds[s]["track_id"].max()) if len(seeds[s]) else 0)
    if ctx_hook:
        ctx_hook(ctx)
    LA = max_ref + 1
    tracks, extra = {}, {}
    pushed = first = 0
    while first < n_frames:
        n_out = min(window, n_frames - first)
        want = min(n_frames, first + n_out + LA)
        if want > pushed:
            r, o, fl = pack_streams(per_stream, n_frames, pushed, want)
            g = None if grey is None else np.stack([grey[s][pushed:want] for s in range(S)])
            if grey_stride:     # rows of grey_stride bytes (AVFrame::linesize), the padding filled with junk; 16-byte records
                gp = np.full(g.shape[:3] + (grey_stride,), 0xA5, np.uint8)
                gp[..., :W] = g
                ctx.push_frames_packed(want - pushed, lib.pack_records(r), o, fl, gp, grey_stride)
            else:
                ctx.push_frames(want - pushed, r, o, fl, g)
            pushed = want
        ctx.raster(first, n_out)
        ctx.extract(first, n_out)
        if poses:
            ctx.track_poses(first, n_out)
            P, ninl = ctx.poses(first, n_out)
            for s in range(S):
                for k in range(n_out):
                    extra[(s, first + k)] = (P[s, k].copy(), int(ninl[s, k]))
        for s in range(S):
            for f in range(first, first + n_out):
                tracks[(s, f)] = ctx.tracks(s, f)
        first += n_out
    return tracks, extra, ctx


def oracle_tracks(orc, stream, W, H, max_ref, grey=None, seeds=None, max_tracks=4096, threshold=25,
                  coverage_threshold=0.20):
    """Per-frame track tables from the oracle for one stream."""
    r, o, fl = stream
    clip = orc.Clip(W, H, r, o, fl, max_ref)
    prev = np.zeros(0, T.TRACK) if seeds is None else seeds
    cid = int(prev["track_id"].max()) if len(prev) else 0
    flat = np.full((H, W), 128, np.uint8)
    out = []
    for f in range(len(fl)):
        img = flat if grey is None else grey[f]
        t, _, cid, _ = orc.extract_frame(W, H, fl[f], img, clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f),
                                         prev, cid, threshold=threshold, coverage_threshold=coverage_threshold,
                                         max_tracks=max_tracks)
        out.append(t)
        prev = t
    return out


def assert_tracks_equal(got, want, tag):
    assert len(got) == len(want), (tag, "count", len(got), len(want))
    if got.tobytes() != want.tobytes():
        for name in got.dtype.names:
            a, b = got[name], want[name]
            if a.tobytes() != b.tobytes():
                bad = np.nonzero((a != b).reshape(len(a), -1).any(1))[0]
                raise AssertionError("%s: field %s differs at %d rows, first %d: got %s want %s" %
                                     (tag, name, len(bad), bad[0], a[bad[0]], b[bad[0]]))
        raise AssertionError("%s: tracks differ" % (tag,))
