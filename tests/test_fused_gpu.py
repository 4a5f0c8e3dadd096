"""Fused (grid-free) mode, MOVFE_CFG_NO_GRID: the per-pixel slot grid is never written; propagation resolves the four slots
of a pixel from its 32x32 tile's ordered hop queue (common.cuh: resolve_slots). Same track tables as the grid path and as
the oracle, bit for bit - including tiles whose queue overflows (dense 4x4 fields: the frame's hop list is scanned), the
lattice back-fill (which asks for uncovered pixels) and a local map that changes every window."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import assert_tracks_equal, oracle_tracks, pack_streams

pytestmark = pytest.mark.gpu


def _run(ctx, streams, grey, n_frames, F, LA, seeds=None):
    S = len(streams)
    if seeds is not None:
        for s in range(S):
            ctx.set_tracks(s, seeds[s], int(seeds[s]["track_id"].max()))
    pushed = first = 0
    out = {}
    while first < n_frames:
        n_out = min(F, n_frames - first)
        want = min(n_frames, first + n_out + LA)
        if want > pushed:
            r, o, fl = pack_streams(streams, n_frames, pushed, want)
            ctx.push_frames(want - pushed, r, o, fl, None if grey is None else np.stack([g[pushed:want] for g in grey]))
            pushed = want
        ctx.raster(first, n_out)
        ctx.extract(first, n_out)
        for s in range(S):
            for f in range(first, first + n_out):
                out[(s, f)] = ctx.tracks(s, f)
        first += n_out
    return out


@pytest.mark.parametrize("cov_thr", [0.20, 0.95])
def test_fused_equals_grid_mode_and_oracle(orc, cov_thr):
    W, H, NF, K, F = 640, 480, 11, 3, 4
    specs = [synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0800 + s, phase=0.2 * s) for s in range(2)]
    streams = [synth.make_records(sp) for sp in specs]
    grey = [synth.make_grey(sp) for sp in specs]
    res = {}
    for og in (True, False):
        ctx = lib.Context(2, W, H, max_records_per_frame=4800, max_ref=K, window_frames=F, max_tracks=8192, coverage_threshold=cov_thr,
                          output_grid=og)
        res[og] = _run(ctx, streams, grey, NF, F, K + 1)
        if not og:
            with pytest.raises(lib.MovfeError, match="MOVFE_CFG_NO_GRID"):
                ctx.grid(0, NF - 1)
        ctx.close()
    for s in range(2):
        want = oracle_tracks(orc, streams[s], W, H, K, grey=grey[s], max_tracks=8192, coverage_threshold=cov_thr)
        for f in range(NF):
            assert_tracks_equal(res[False][(s, f)], want[f], ("fused", s, f))
            assert_tracks_equal(res[True][(s, f)], want[f], ("grid", s, f))
    if cov_thr > 0.9:
        assert any(((res[False][(0, f)]["flags"] & T.TRACK_COVERAGE) != 0).any() for f in range(NF))   # the back-fill ran


def test_fused_dense4x4_overflowing_tiles(orc):
    """one 4x4 record per 4x4 cell: 64 blocks of 5x5 covered pixels meet every 32x32 tile, plus ref chains -> queues beyond
    MOVFE_TILE_Q entries fall back to the frame's hop list"""
    W, H, NF = 480, 272, 5
    sp = synth.Spec(W, H, n_frames=NF, refs=3, seed=0x5EED0810, dense4x4=True, start_p=True)
    seeds = synth.seed_tracks_lattice(sp)
    stream = synth.make_records(sp)
    ctx = lib.Context(1, W, H, max_records_per_frame=(W // 4) * (H // 4), max_ref=2, window_frames=2, max_tracks=2048, has_grey=False,
                      output_grid=False)
    got = _run(ctx, [stream], None, NF, 2, 3, seeds=[seeds])
    want = oracle_tracks(orc, stream, W, H, 2, seeds=seeds, max_tracks=2048)
    for f in range(NF):
        assert_tracks_equal(got[(0, f)], want[f], (0, f))
    st = ctx.workload_stats()
    assert st["frames_rastered"] == NF and st["hops_per_frame"] > 5000 and st["candidates_per_track"] > 1.0
    ctx.close()
