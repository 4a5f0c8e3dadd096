"""GPU parity: track propagation + EXPRESS (through the C-ABI) against the CPU oracle — bit-exact track tables."""
import numpy as np
import pytest

from movfe import synth, types as T

from gpu_util import assert_tracks_equal, oracle_tracks, run_frontend_clip

pytestmark = pytest.mark.gpu


def _run(orc, specs, window, max_ref, with_grey=True, seeds=None, max_tracks=4096, **kw):
    streams = [synth.make_records(sp) for sp in specs]
    grey = [synth.make_grey(sp) for sp in specs] if with_grey else None
    W, H, NF = specs[0].W, specs[0].H, specs[0].n_frames
    got, _, ctx = run_frontend_clip(streams, W, H, NF, window, max_ref, grey=grey, seeds=seeds, max_tracks=max_tracks, **kw)
    total = 0
    for s, sp in enumerate(specs):
        want = oracle_tracks(orc, streams[s], W, H, max_ref, grey=None if grey is None else grey[s],
                             seeds=None if seeds is None else seeds[s], max_tracks=max_tracks, **kw)
        for f in range(NF):
            assert_tracks_equal(got[(s, f)], want[f], (s, f))
            total += len(want[f])
    ctx.close()
    return total


def test_textured_scene_ref4(orc):
    specs = [synth.Spec(640, 480, n_frames=12, refs=4, seed=0x5EED0010 + s, phase=0.4 * s) for s in range(2)]
    assert _run(orc, specs, window=5, max_ref=3) > 20000


def test_textured_euroc_shape(orc):
    specs = [synth.Spec(752, 480, n_frames=7, refs=2, seed=0x5EED0011, fx=458.654, fy=457.296, cx=367.215, cy=248.375)]
    assert _run(orc, specs, window=3, max_ref=1) > 3000


def test_backfill_and_low_threshold(orc):
    # a high coverage threshold forces the coverage back-fill pass on every P frame
    specs = [synth.Spec(320, 240, n_frames=6, refs=2, seed=0x5EED0012)]
    assert _run(orc, specs, window=6, max_ref=1, coverage_threshold=2.0, threshold=10) > 500


def test_mv_only_seeded(orc):
    specs = [synth.Spec(640, 480, n_frames=9, refs=4, seed=0x5EED0013 + s, start_p=True) for s in range(2)]
    seeds = [synth.seed_tracks_lattice(sp) for sp in specs]
    assert _run(orc, specs, window=4, max_ref=3, with_grey=False, seeds=seeds) > 5000


def test_table_truncation(orc):
    specs = [synth.Spec(640, 480, n_frames=5, refs=2, seed=0x5EED0014)]
    _run(orc, specs, window=5, max_ref=1, max_tracks=700)


def test_dense4x4_mv_only(orc):
    specs = [synth.Spec(480, 272, n_frames=4, refs=1, seed=0x5EED0015, dense4x4=True, start_p=True)]
    seeds = [synth.seed_tracks_lattice(sp) for sp in specs]
    streams = [synth.make_records(sp) for sp in specs]
    got, _, ctx = run_frontend_clip(streams, 480, 272, 4, 2, 0, seeds=seeds, max_records=(480 // 4) * (272 // 4))
    want = oracle_tracks(orc, streams[0], 480, 272, 0, seeds=seeds[0])
    for f in range(4):
        assert_tracks_equal(got[(0, f)], want[f], (0, f))
    ctx.close()


def test_single_shot_extract_frame(orc):
    """movfe_extract_frame: MOVExtractor::operator() on host raster results (the shim's entry point), frame by frame."""
    from movfe import lib
    sp = synth.Spec(320, 240, n_frames=6, refs=3, seed=0x5EED0016)
    r, o, fl = synth.make_records(sp)
    grey = synth.make_grey(sp)
    clip = orc.Clip(sp.W, sp.H, r, o, fl, 2)
    ctx = lib.Context(1, sp.W, sp.H, max_records_per_frame=1200, max_ref=2, window_frames=1, max_tracks=2048)
    prev_g, prev_w = np.zeros(0, T.TRACK), np.zeros(0, T.TRACK)
    cid_g = cid_w = 0
    for f in range(sp.n_frames):
        args = (clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f))
        got, cid_g = ctx.extract_frame(fl[f], grey[f], *args, prev_g, cid_g)
        want, _, cid_w, _ = orc.extract_frame(sp.W, sp.H, fl[f], grey[f], *args, prev_w, cid_w, max_tracks=2048)
        assert_tracks_equal(got, want, ("single", f))
        assert cid_g == cid_w
        prev_g, prev_w = got, want
    ctx.close()


def test_decoder_context_feeding_extractor_context(orc):
    """The shims' arrangement: one context rasterises (whole ref range as look-ahead), another one extracts frame by frame
    from the downloaded results, a third context exists beside them."""
    from movfe import lib
    W, H, NF = 320, 240, 7
    sp = synth.Spec(W, H, n_frames=NF, refs=3, seed=0x5EED0020, fx=160.0, fy=160.0)
    r, o, fl = synth.make_records(sp)
    grey = synth.make_grey(sp)
    clip = orc.Clip(W, H, r, o, fl, 10)
    rctx = lib.Context(1, W, H, max_records_per_frame=1200, max_ref=10, window_frames=1, max_tracks=1, max_map_points=1, has_grey=False)
    ectx = lib.Context(1, W, H, max_records_per_frame=1200, max_ref=10, window_frames=1, max_tracks=8192, max_map_points=1)
    octx = lib.Context(1, 16, 16, max_records_per_frame=1, max_ref=0, window_frames=1, max_tracks=1, max_map_points=1, has_grey=False)
    for f in range(NF):
        rctx.push_frames(1, r[o[f]:o[f + 1]], np.array([0, o[f + 1] - o[f]], np.int64), fl[f:f + 1])
    prev_g, prev_w = np.zeros(0, T.TRACK), np.zeros(0, T.TRACK)
    cid_g = cid_w = 0
    for f in range(NF):
        rctx.raster(f, 1)
        nh, nk, cov = rctx.raster_counts(0, f)
        got, cid_g = ectx.extract_frame(fl[f], grey[f], rctx.grid(0, f), rctx.hops(0, f), rctx.kps(0, f), cov, prev_g, cid_g)
        want, _, cid_w, _ = orc.extract_frame(W, H, fl[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev_w,
                                              cid_w, max_tracks=8192)
        assert_tracks_equal(got, want, ("two-context", f))
        prev_g, prev_w = got, want
    for c in (rctx, ectx, octx):
        c.close()


@pytest.mark.parametrize("pad,env", [(48, {}), (0, {"MOVFE_GREY_DIRECT": "1"})])
def test_strided_luma_planes(orc, pad, env, monkeypatch):
    """Luma planes whose rows are grey_stride > width bytes apart (AVFrame::linesize, cv::Mat::step) go straight into the device's
    pitched ring (movfe_push_frames_packed); the ring wraps several times over the clip. Tables bit-exact. MOVFE_GREY_DIRECT=1 sends
    tightly packed planes the same way (the seeding push of other tests); strided ones always go row by row."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    W, H, NF, K = 320, 240, 26, 2
    specs = [synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED00D1 + s, phase=0.3 * s) for s in range(2)]
    streams = [synth.make_records(sp) for sp in specs]
    greys = [synth.make_grey(sp) for sp in specs]
    got, _, ctx = run_frontend_clip(streams, W, H, NF, 3, K, grey=greys, grey_stride=(W + pad) if pad else 0)
    ctx.close()
    for s in range(2):
        want = oracle_tracks(orc, streams[s], W, H, K, grey=greys[s])
        for f in range(NF):
            assert_tracks_equal(got[(s, f)], want[f], (s, f))
