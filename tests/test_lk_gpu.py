"""Host LK hand-over (src/MOVExtractor.cc:81-120, 161-243, 337-377): cv::calcOpticalFlowPyrLK stays on
the host, its results enter through movfe_set_lk_results / movfe_extract_frame and are merged on the device exactly where
the reference merges them. The oracle's LK-injected mode is pinned to the reference's own MOVExtractor.cc by
tests/test_ref_parity.py; here the CUDA path is compared with it bit for bit."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import assert_tracks_equal, pack_streams

pytestmark = pytest.mark.gpu


def _popc(tr):
    return np.array([sum(bin(int(w)).count("1") for w in t["desc"]) for t in tr], np.int64)


def lk_request(prev, is_p):
    """Indices of the previous table's features the reference hands to LK for the next frame, in its order (movfe.h)."""
    if not is_p:
        return list(range(len(prev)))
    pc = _popc(prev)
    order = sorted(range(len(prev)), key=lambda i: (-int(prev["age"][i]), -int(pc[i])))
    return [i for i in order if prev["flags"][i] & T.TRACK_COVERAGE]


def fake_lk(rng, prev, idx, W, H):
    n = len(idx)
    st = (rng.random(n) > 0.25).astype(np.uint8)
    out = np.stack([prev["pt_x"][idx], prev["pt_y"][idx]], 1).astype(np.float32) + rng.normal(0, 1.5, (n, 2)).astype(np.float32)
    far = rng.random(n) < 0.1
    out[far] += np.float32(max(W, H))
    return st, out.astype(np.float32)


@pytest.mark.parametrize("seed,cov_thr,iframe_at,strided", [(0, 0.95, None, False), (1, 0.95, 4, True), (2, 0.20, 3, False)])
def test_single_shot_extract_frame_with_lk(orc, seed, cov_thr, iframe_at, strided):
    W, H, NF, K = 320, 240, 8, 2
    sp = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0700 + seed, fx=160.0, fy=160.0)
    r, o, fl = synth.make_records(sp)
    fl = fl.copy()
    if iframe_at is not None:
        fl[iframe_at] &= ~np.uint8(T.FRAME_P)
    grey = synth.make_grey(sp)
    clip = orc.Clip(W, H, r, o, fl, K)
    ctx = lib.Context(1, W, H, max_records_per_frame=1200, max_ref=K, window_frames=1, max_tracks=4096, coverage_threshold=cov_thr)
    rng = np.random.Generator(np.random.PCG64(0x7700 + seed))
    prev_g, prev_w = np.zeros(0, T.TRACK), np.zeros(0, T.TRACK)
    cid_g = cid_w = 0
    carried = 0
    pad = np.zeros((H, W + 37), np.uint8)
    for f in range(NF):
        is_p = bool(fl[f] & T.FRAME_P)
        idx = lk_request(prev_w, is_p)
        lk = fake_lk(rng, prev_w, idx, W, H) if idx else None
        reloc = None
        if f == NF - 1:  # lost relocalisation seeds on the last frame
            n = 40
            reloc = np.zeros(n, T.RELOC_SEED)
            reloc["track_id"] = rng.integers(1, 500, n)
            reloc["q_indx"] = np.arange(n) * 2
            reloc["x"], reloc["y"] = rng.uniform(0, W, n), rng.uniform(0, H, n)
        args = (clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f))
        img = grey[f]
        if strided:       # a cv::Mat ROI / AVFrame plane: rows of W bytes inside rows of W + 37
            pad[:, :W] = grey[f]
            img = pad[:, :W]
        got, cid_g = ctx.extract_frame(fl[f], img, *args, prev_g, cid_g, lk_status=None if lk is None else lk[0],
                                       lk_pts=None if lk is None else lk[1], reloc=reloc)
        want, _, cid_w, _ = orc.extract_frame(W, H, fl[f], grey[f], *args, prev_w, cid_w, coverage_threshold=cov_thr, max_tracks=4096,
                                              lk_status=None if lk is None else lk[0], lk_pts=None if lk is None else lk[1], reloc=reloc)
        assert_tracks_equal(got, want, ("lk single", f))
        assert cid_g == cid_w
        carried += int(((want["q_indx"] >= 0) & (((want["flags"] & T.TRACK_COVERAGE) != 0) | (not is_p))).sum())
        prev_g, prev_w = got, want
    assert carried > 0
    assert ctx.dropped_lk_tracks() == 0
    ctx.close()


def test_batched_lk_one_frame_per_call_and_drop_counter(orc):
    """The batched path with host LK: one frame per movfe_extract, results installed per stream before it. A second context
    runs the same clip without results: the carried tracks are dropped (the oracle's lk_status == NULL mode) and counted."""
    W, H, NF, K, S = 160, 112, 9, 1, 3
    specs = [synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0710 + s, fx=80.0, fy=80.0, phase=0.3 * s) for s in range(S)]
    streams = [synth.make_records(sp) for sp in specs]
    for st in streams:
        st[2][5] &= ~np.uint8(T.FRAME_P)          # an intra picture in mid-stream
    grey = [synth.make_grey(sp) for sp in specs]
    clips = [orc.Clip(W, H, *st, K) for st in streams]
    rng = np.random.Generator(np.random.PCG64(0x7710))
    for with_lk in (True, False):
        ctx = lib.Context(S, W, H, max_records_per_frame=400, max_ref=K, window_frames=1, max_tracks=2048, coverage_threshold=0.95)
        prev = [np.zeros(0, T.TRACK) for _ in range(S)]
        cid = [0] * S
        pushed = 0
        dropped_want = 0
        for f in range(NF):
            want_pushed = min(NF, f + 1 + K + 1)
            if want_pushed > pushed:
                rr, oo, ff = pack_streams(streams, NF, pushed, want_pushed)
                ctx.push_frames(want_pushed - pushed, rr, oo, ff, np.stack([g[pushed:want_pushed] for g in grey]))
                pushed = want_pushed
            ctx.raster(f, 1)
            lks = []
            for s in range(S):
                is_p = bool(streams[s][2][f] & T.FRAME_P)
                idx = lk_request(prev[s], is_p)
                lk = fake_lk(rng, prev[s], idx, W, H) if idx else None
                if not with_lk:
                    dropped_want += len(idx)
                    lk = None
                elif lk is not None:
                    ctx.set_lk_results(s, lk[0], lk[1])
                lks.append(lk)
            ctx.extract(f, 1)
            for s in range(S):
                lk = lks[s]
                want, _, cid[s], _ = orc.extract_frame(W, H, streams[s][2][f], grey[s][f], clips[s].grid(f), clips[s].hops(f), clips[s].kps(f),
                                                       clips[s].coverage(f), prev[s], cid[s], coverage_threshold=0.95, max_tracks=2048,
                                                       lk_status=None if lk is None else lk[0], lk_pts=None if lk is None else lk[1])
                assert_tracks_equal(ctx.tracks(s, f), want, ("lk batched", with_lk, s, f))
                prev[s] = want
        assert ctx.dropped_lk_tracks() == dropped_want
        assert with_lk or dropped_want > 0
        ctx.close()


def test_extract_after_the_ring_moved_on_is_refused():
    """push(k), raster(k), push(k+1), push(k+2), extract(k): the grey planes of window k are gone."""
    W, H, F, K = 160, 112, 2, 1
    sp = synth.Spec(W, H, n_frames=16, refs=K + 1, seed=0x5EED0720, fx=80.0, fy=80.0)
    stream, grey = synth.make_records(sp), synth.make_grey(sp)
    ctx = lib.Context(1, W, H, max_records_per_frame=400, max_ref=K, window_frames=F, max_tracks=512)
    ring = 2 * F + K + 1

    def push(a, b):
        r, o, fl = pack_streams([stream], 16, a, b)
        ctx.push_frames(b - a, r, o, fl, grey[None, a:b])

    push(0, F + K + 1)
    ctx.raster(0, F)
    push(F + K + 1, F + K + 1 + ring - 1)        # the ring now holds nothing of frame 0
    with pytest.raises(lib.MovfeError, match="movfe error -4:.*left the ring"):
        ctx.extract(0, F)
    ctx.close()
