"""GPU parity: the CUDA raster (through the C-ABI) against the CPU oracle — bit-exact grid / hops / kps / coverage."""
import numpy as np
import pytest

from movfe import synth, types as T

from gpu_util import assert_raster_equal, run_raster_clip

pytestmark = pytest.mark.gpu


def _rec(dst, src, w=16, h=16, ref=0, source=-1):
    r = np.zeros((), T.MV_RECORD)
    r["source"], r["w"], r["h"], r["ref"] = source, w, h, ref
    r["dst_x"], r["dst_y"], r["src_x"], r["src_y"] = dst[0], dst[1], src[0], src[1]
    return r


def _clip(frames):
    recs, off, flags = [], [0], []
    for i, fr in enumerate(frames):
        fl = T.FRAME_P if i > 0 else 0
        if fr is not None:
            fl |= T.FRAME_MV
            recs += fr
        off.append(len(recs))
        flags.append(fl)
    arr = np.array(recs, T.MV_RECORD) if recs else np.zeros(0, T.MV_RECORD)
    return arr, np.array(off, np.int64), np.array(flags, np.uint8)


def _check(orc, streams, W, H, window, max_ref, max_records=4800):
    n_frames = len(streams[0][2])
    got, ctx = run_raster_clip(streams, W, H, n_frames, window, max_ref, max_records=max_records)
    for s, (r, o, fl) in enumerate(streams):
        clip = orc.Clip(W, H, r, o, fl, max_ref)
        for f in range(n_frames):
            assert_raster_equal(clip, got, s, f)
    rej = ctx.rejected_records()
    ctx.close()
    return rej


def test_kats_on_gpu(orc):
    streams = [
        _clip([[], [_rec((24, 24), (20, 22))]] + [[]] * 3),                                   # KAT-1
        _clip([[], [], [], [_rec((24, 24), (20, 22), ref=1)], []]),                           # KAT-2
        _clip([[], [_rec((24, 24), (24, 24)) for _ in range(5)], [], [], []]),                # KAT-3
        _clip([[], [_rec((56, 24), (50, 24))], [], [], []]),                                  # KAT-5
        _clip([[], [], [], [], [_rec((28, 20), (21, 16), w=8, h=8, ref=2)]]),                 # KAT-10
        _clip([[], None, [_rec((24, 24), (20, 22), ref=1)], None, []]),                       # mv-disabled frame
        _clip([[_rec((24, 24), (20, 22), ref=1)], [_rec((24, 24), (20, 22), ref=3)], [], [], []]),  # window-start drops
    ]
    _check(orc, streams, 64, 48, window=2, max_ref=3)


def test_b_frames_source0_negative_ref(orc):
    fr = [[], [_rec((24, 24), (20, 22), source=1), _rec((30, 20), (28, 20), source=0, ref=1),
               _rec((20, 30), (28, 20), ref=-1), _rec((20, 20), (21, 20), source=1, ref=2)], [], []]
    _check(orc, [_clip(fr)], 64, 48, window=4, max_ref=3)


def test_bad_ref_counted(orc):
    fr = [[], [_rec((24, 24), (20, 22), ref=4), _rec((24, 24), (20, 22), ref=1)], []]
    assert _check(orc, [_clip(fr)], 64, 48, window=3, max_ref=3) == 1


@pytest.mark.parametrize("W,H,window", [(640, 480, 4), (752, 480, 3), (100, 52, 5)])
def test_synthetic_scene(orc, W, H, window):
    streams = []
    for s in range(3):
        spec = synth.Spec(W, H, n_frames=11, refs=4, seed=0x5EED0002 + s, phase=0.3 * s)
        streams.append(synth.make_records(spec))
    _check(orc, streams, W, H, window, max_ref=3)


def test_stereo_packed(orc):
    spec = synth.Spec(752, 480, n_frames=10, refs=2, seed=0x5EED0003, stereo=True, fx=458.654, fy=457.296,
                      cx=367.215, cy=248.375)
    _check(orc, [synth.make_records(spec)], 752, 480, window=4, max_ref=1)


def test_dense_4x4(orc):
    spec = synth.Spec(480, 272, n_frames=4, refs=1, seed=0x5EED0004, dense4x4=True)
    _check(orc, [synth.make_records(spec)], 480, 272, window=2, max_ref=0, max_records=(480 // 4) * (272 // 4))


def test_random_unordered_records(orc):
    """Adversarial input: random positions (also outside the image), sizes 4/8/16, refs 0..3, no spatial order."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0005))
    W, H, NF = 200, 120, 9
    streams = []
    for s in range(2):
        frames = [[]]
        for f in range(1, NF):
            n = int(rng.integers(0, 900))
            r = np.zeros(n, T.MV_RECORD)
            r["source"] = rng.choice([-1, -1, -1, 0, 1], n)
            r["w"], r["h"] = rng.choice([4, 8, 16], n), rng.choice([4, 8, 16], n)
            r["dst_x"], r["dst_y"] = rng.integers(-20, W + 20, n), rng.integers(-20, H + 20, n)
            r["src_x"] = r["dst_x"] + rng.integers(-80, 80, n)
            r["src_y"] = r["dst_y"] + rng.integers(-80, 80, n)
            r["ref"] = rng.integers(0, min(3, f - 1) + 1, n)
            frames.append(list(r))
        streams.append(_clip(frames))
    _check(orc, streams, W, H, window=3, max_ref=3)


def test_deep_stacks_and_band_overflow(orc):
    """> 31 candidates in one tile (multi-chunk path), > 124 in one tile (row-streaming path from the staged list)
    and more hops in one band than the staged list holds (row-streaming path from global memory)."""
    many = [_rec((40 + (i % 7), 24 + (i % 5)), (38, 22)) for i in range(100)]
    more = [_rec((60 + (i % 9), 30 + (i % 11)), (58, 28)) for i in range(300)]
    huge = [_rec((100 + (i % 50), 40 + (i % 3)), (90, 40), w=8, h=8) for i in range(4800)]
    _check(orc, [_clip([[], many, more, huge, []])], 256, 64, window=5, max_ref=0)


def test_wide_blocks_span_many_tiles(orc):
    """Blocks wider than a 32-px tile (synthetic: H.264 stops at 16): a hop meets three or more tiles, which takes the
    slot-grid kernel's generic per-tile ballots instead of the match.any ranks. Mixed with ordinary blocks, several streams."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0007))
    W, H, NF = 328, 136, 6
    streams = []
    for s in range(3):
        frames = [[]]
        for f in range(1, NF):
            n = int(rng.integers(50, 400))
            r = np.zeros(n, T.MV_RECORD)
            r["source"] = -1
            r["w"] = rng.choice([8, 16, 40, 64, 100, 200], n, p=[.3, .3, .1, .1, .1, .1])
            r["h"] = rng.choice([8, 16, 48, 90], n, p=[.4, .4, .1, .1])
            r["dst_x"], r["dst_y"] = rng.integers(0, W, n), rng.integers(0, H, n)
            r["src_x"] = r["dst_x"] + rng.integers(-20, 20, n)
            r["src_y"] = r["dst_y"] + rng.integers(-20, 20, n)
            r["ref"] = rng.integers(0, min(2, f - 1) + 1, n)
            frames.append(list(r))
        streams.append(_clip(frames))
    _check(orc, streams, W, H, window=3, max_ref=2)


def test_more_chunks_in_a_band_than_the_chunk_list_holds(orc):
    """> 1024 32-hop chunks meet one 32-row band: the slot-grid kernel streams the whole hop list per row (its last resort)."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0008))
    W, H, n = 160, 40, 34000
    r = np.zeros(n, T.MV_RECORD)
    r["source"] = -1
    r["w"] = r["h"] = 4
    r["dst_x"], r["dst_y"] = rng.integers(4, W - 5, n), rng.integers(4, H - 5, n)
    r["src_x"] = r["dst_x"] + rng.integers(-2, 3, n)
    r["src_y"] = r["dst_y"] + rng.integers(-2, 3, n)
    _check(orc, [_clip([[], list(r), []])], W, H, window=3, max_ref=0, max_records=n)


def test_max_ref_10(orc):
    rng = np.random.Generator(np.random.PCG64(0x5EED0006))
    W, H, NF = 128, 96, 16
    frames = [[]]
    for f in range(1, NF):
        n = 60
        r = np.zeros(n, T.MV_RECORD)
        r["source"] = -1
        r["w"] = r["h"] = 16
        r["dst_x"], r["dst_y"] = rng.integers(8, W - 9, n), rng.integers(8, H - 9, n)
        r["src_x"] = r["dst_x"] + rng.integers(-30, 30, n)
        r["src_y"] = r["dst_y"] + rng.integers(-30, 30, n)
        r["ref"] = rng.integers(0, min(10, f - 1) + 1, n)
        frames.append(list(r))
    _check(orc, [_clip(frames)], W, H, window=2, max_ref=10)


def test_single_frame_pushes_full_ref_range(orc):
    """The decoder shim's usage: one frame per push, window of one frame, the reference's whole ref range as look-ahead."""
    from movfe import lib
    W, H, NF = 320, 240, 9
    sp = synth.Spec(W, H, n_frames=NF, refs=3, seed=0x5EED0031)
    r, o, fl = synth.make_records(sp)
    clip = orc.Clip(W, H, r, o, fl, 10)
    ctx = lib.Context(1, W, H, max_records_per_frame=1200, max_ref=10, window_frames=1, has_grey=False)
    for f in range(NF):
        ctx.push_frames(1, r[o[f]:o[f + 1]], np.array([0, o[f + 1] - o[f]], np.int64), fl[f:f + 1])
    got = {}
    for f in range(NF):
        ctx.raster(f, 1)
        got[(0, f)] = dict(grid=ctx.grid(0, f), hops=ctx.hops(0, f), kps=ctx.kps(0, f), cov=ctx.raster_counts(0, f)[2])
        assert_raster_equal(clip, got, 0, f)
    ctx.close()


def test_packed_push_equals_record_push(orc):
    """movfe_push_frames_packed (16-byte records packed by movfe_pack_records on the host) gives the raster results of the
    40-byte push, bit for bit: hop lists, kps, slot grids and coverage against the oracle."""
    W, H, NF, K = 320, 240, 9, 3
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED00C7)
    stream = synth.make_records(spec)
    got, ctx = run_raster_clip([stream, stream], W, H, NF, window=4, max_ref=K, packed=True)
    ctx.close()
    clip = orc.Clip(W, H, *stream, K)
    for s in range(2):
        for f in range(NF):
            assert_raster_equal(clip, got, s, f)
