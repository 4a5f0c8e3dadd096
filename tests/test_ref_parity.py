"""The oracle against THE REFERENCE'S OWN CODE.

oracle/_ref/libmovref{,_canon}.so are /root/reference's src/VideoDecoder.cc, src/MOVExtractor.cc, include/EXPRESS.h and
include/MOVMatcher.h (with include/Frame.h, MOVExtractor.h, VideoDecoder.h, VideoBase.h), compiled UNMODIFIED by
`make -C oracle ref` against the stand-in headers of oracle/ref_standin/ (OpenCV containers, a fake libav decoder that
hands out synthetic frames, injected cv::calcOpticalFlowPyrLK results). These tests pin every bit-exact claim of the
oracle - hop lists, kps, slot grids, coverage, descriptors, compute_express verdicts, track tables, match indices - to
that code on random inputs. The pose solver is not covered: the reference's PoseOptimization is cv::solvePnPRansac from
un-vendored OpenCV (parity unpinned).

Skipped when neither the prebuilt libraries nor /root/reference are present."""
import numpy as np
import pytest

from movfe import synth, types as T
from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def ref():
    pyref.lib("plain")
    pyref.lib("canon")
    return pyref


# ------------------------------------------------------------------------------------------------------ raster -----
def _random_clip(rng, W, H, NF, max_ref, sizes, qlen=12, n_max=70, mv_off=0.15):
    frames, flags = [], []
    for f in range(NF):
        mv_on = f > 0 and rng.random() > mv_off
        n = int(rng.integers(0, n_max)) if f > 0 else 0
        r = np.zeros(n, T.MV_RECORD)
        r["source"] = rng.choice([-1, -1, -1, -1, 0, 1], n)
        r["w"], r["h"] = rng.choice(sizes, n), rng.choice(sizes, n)
        r["dst_x"], r["dst_y"] = rng.integers(-6, W + 6, n), rng.integers(-6, H + 6, n)
        r["src_x"] = r["dst_x"] + rng.integers(-40, 41, n)
        r["src_y"] = r["dst_y"] + rng.integers(-40, 41, n)
        # the reference under-runs its deque (undefined behaviour) unless ref < frames queued: VideoDecoder.cc:247,322
        r["ref"] = rng.integers(0, max(min(max_ref, min(f, qlen - 1) - 1), 0) + 1, n)
        r["motion_scale"] = 4
        frames.append(r)
        flags.append((T.FRAME_P if f > 0 else 0) | (T.FRAME_MV if mv_on else 0))
    off = np.cumsum([0] + [len(fr) for fr in frames]).astype(np.int64)
    recs = np.concatenate(frames) if off[-1] else np.zeros(0, T.MV_RECORD)
    return recs, off, np.array(flags, np.uint8)


def _assert_raster_equal(want, got, NF, flags):
    total = 0
    for f in range(NF):
        assert want.frame_no(f) == f + 1 and want.is_p(f) == bool(flags[f] & T.FRAME_P)
        assert got.n_hops(f) == want.n_hops(f) and got.n_kps(f) == want.n_kps(f), (f, got.n_hops(f), want.n_hops(f))
        assert got.hops(f).tobytes() == want.hops(f).tobytes(), (f, "hops")
        assert got.kps(f).tobytes() == want.kps(f).tobytes(), (f, "kps")
        assert np.array_equal(got.grid(f), want.grid(f)), (f, "grid")
        if flags[f] & T.FRAME_MV and want.n_kps(f) + want.n_hops(f) > 0:
            # VideoImage::coverageArea is only assigned when side data was processed (VideoDecoder.cc:350)
            assert got.coverage(f) == want.coverage(f), (f, got.coverage(f), want.coverage(f))
        total += want.n_hops(f)
    return total


@pytest.mark.parametrize("seed,W,H,NF,max_ref,sizes", [(1, 64, 48, 8, 3, [4, 8, 16]), (2, 97, 61, 20, 2, [8, 16]), (3, 40, 40, 6, 0, [4, 8, 16]),
                                                       (4, 128, 72, 30, 10, [8, 16]), (5, 80, 64, 16, 5, [4, 8, 16])])
def test_raster_oracle_equals_reference_videodecoder(orc, ref, seed, W, H, NF, max_ref, sizes):
    """VideoDecoder::NextImage (12-deep deque, look-ahead back-fill of hops and kps, per-pixel slot writes) on random records:
    blocks hanging over every border, B-type and source-0 records, frames without side data, ref up to 10."""
    rng = np.random.Generator(np.random.PCG64(0x2EF0 + seed))
    recs, off, flags = _random_clip(rng, W, H, NF, max_ref, sizes)
    want = ref.Clip(W, H, recs, off, flags)
    got = orc.Clip(W, H, recs, off, flags, max_ref)
    assert got.bad_ref() == 0
    assert _assert_raster_equal(want, got, NF, flags) > 50


@pytest.mark.parametrize("kw,max_ref", [(dict(width=160, height=112, n_frames=18, refs=4, seed=0x5EED0501, fx=80.0, fy=80.0), 3),
                                         (dict(width=192, height=96, n_frames=16, refs=2, seed=0x5EED0502, fx=96.0, fy=96.0, stereo=True), 1),
                                         (dict(width=96, height=64, n_frames=5, refs=1, seed=0x5EED0503, fx=48.0, fy=48.0, dense4x4=True), 0)])
def test_raster_oracle_equals_reference_on_benchmark_shaped_clips(orc, ref, kw, max_ref):
    """The generator bench.py uses (ffmpeg-like macroblock partitions, ref chains, stereo frame packing, dense 4x4 fields)."""
    sp = synth.Spec(**kw)
    recs, off, flags = synth.make_records(sp)
    grey = synth.make_grey(sp)
    want = ref.Clip(sp.W, sp.H, recs, off, flags, grey=grey)
    got = orc.Clip(sp.W, sp.H, recs, off, flags, max_ref)
    assert _assert_raster_equal(want, got, sp.n_frames, flags) > 200
    for f in (0, sp.n_frames - 1):
        assert np.array_equal(want.grey(f), grey[f])   # the fake decoder delivered the luma plane untouched


# ----------------------------------------------------------------------------------------------------- EXPRESS -----
@pytest.mark.parametrize("shape", [(16, 16), (8, 8), (16, 8), (8, 16)])
def test_express_oracle_equals_reference_header(orc, ref, shape):
    rows, cols = shape
    rng = np.random.Generator(np.random.PCG64(0x2EE0 + rows * 3 + cols))
    spec = synth.Spec(160, 120, n_frames=2, refs=1, seed=0x5EED0061)
    textured = synth.make_grey(spec)[1]
    noisy = rng.integers(0, 256, (120, 160)).astype(np.uint8)
    steps = (np.add.outer(np.arange(120) // 9, np.arange(160) // 7) % 2 * 180 + 20).astype(np.uint8)
    smooth = (np.add.outer(np.arange(120), np.arange(160)) % 256).astype(np.uint8)   # exercises the uint8 wrap of the band
    n_true, descs = 0, []
    for img in (textured, noisy, steps, smooth):
        for _ in range(150):
            x0, y0 = int(rng.integers(0, 160 - cols - 1)), int(rng.integers(0, 120 - rows))
            thr = int(rng.choice([5, 20, 25, 40, 60, 140]))
            assert orc.express_center(img, x0, y0, cols, rows) == ref.express_center(img, x0, y0, cols, rows)
            d = orc.express_descriptor(img, x0, y0, cols, rows, thr)
            assert np.array_equal(d, ref.express_descriptor(img, x0, y0, cols, rows, thr)), (x0, y0, thr)
            e = orc.express_test(img, x0, y0, cols, rows, thr)
            assert e == ref.express_test(img, x0, y0, cols, rows, thr), (x0, y0, thr)
            n_true += e
            descs.append(d)
    assert 0 < n_true < 600
    for a, b in zip(descs[::7], descs[3::7]):
        assert orc.express_distance(a, b) == ref.express_distance(a, b)


# ------------------------------------------------------------------------------------------------ MOVExtractor -----
def _popc(tr):
    return np.array([sum(bin(int(w)).count("1") for w in t["desc"]) for t in tr], np.int64)


def _sorted_order(tr):
    """stable (age desc, popcount desc) order = the canonicalised MOVExtractor.cc:249-252"""
    return sorted(range(len(tr)), key=lambda i: (-int(tr["age"][i]), -int(_popc(tr[i:i + 1])[0])))


def _lk_for(rng, pts_xy, W, H):
    """synthetic LK outcome for the given points: most carried with a small shift, some lost, some pushed outside"""
    n = len(pts_xy)
    st = (rng.random(n) > 0.25).astype(np.uint8)
    out = np.asarray(pts_xy, np.float32).reshape(n, 2) + rng.normal(0, 1.5, (n, 2)).astype(np.float32)
    far = rng.random(n) < 0.1
    out[far] += np.float32(max(W, H))
    return st, out.astype(np.float32)


def _assert_tables_equal(got, want, tag):
    assert len(got) == len(want), (tag, len(got), len(want))
    for name in got.dtype.names:
        assert got[name].tobytes() == want[name].tobytes(), (tag, name)


@pytest.mark.parametrize("seed,cov_thr,iframe_at,W,H,NF,K", [(0, 0.20, None, 160, 112, 8, 2), (1, 0.95, None, 160, 112, 8, 2), (2, 0.95, 4, 160, 112, 8, 2),
                                                                (3, 0.20, 3, 160, 112, 8, 2), (4, 0.20, None, 640, 480, 10, 3)])
def test_extractor_oracle_equals_reference_over_clips(orc, ref, seed, cov_thr, iframe_at, W, H, NF, K):
    """MOVExtractor::operator() frame after frame, each side fed its own previous table: I-frame seeding, propagation with
    candidate choice, claims, descriptor gate, births, lattice back-fill (cov_thr 0.95), and the LK carry-over of coverage
    features and of a mid-stream I frame with the SAME injected cv::calcOpticalFlowPyrLK results on both sides.
    Reference build: 'canon' (stable order of ties, see oracle/ref_standin/sort_canon.h)."""
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0570 + seed, fx=W / 2.0, fy=W / 2.0)   # (4): BASELINE config C2's frame
    recs, off, flags = synth.make_records(spec)
    flags = flags.copy()
    if iframe_at is not None:
        flags[iframe_at] &= ~np.uint8(T.FRAME_P)    # an intra picture in mid-stream: every track is carried by LK (:81-120)
    grey = synth.make_grey(spec)
    rclip = ref.Clip(W, H, recs, off, flags, grey=grey)
    oclip = orc.Clip(W, H, recs, off, flags, K)
    rng = np.random.Generator(np.random.PCG64(0x2ED0 + seed))
    prev_o, cid_o = np.zeros(0, T.TRACK), 0
    prev_r, cid_r = np.zeros(0, T.TRACK), 0
    n_cov_carried = n_iframe_carried = n_backfill = n_short_desc = 0
    for f in range(NF):
        is_p = bool(flags[f] & T.FRAME_P)
        lk = None
        if len(prev_o):
            if is_p:
                order = _sorted_order(prev_o)
                cov = [i for i in order if prev_o["flags"][i] & T.TRACK_COVERAGE]
                if cov:
                    lk = _lk_for(rng, np.stack([prev_o["pt_x"][cov], prev_o["pt_y"][cov]], 1), W, H)
            else:
                lk = _lk_for(rng, np.stack([prev_o["pt_x"], prev_o["pt_y"]], 1), W, H)
        got, _, cid_o, _ = orc.extract_frame(W, H, flags[f], grey[f], oclip.grid(f), oclip.hops(f), oclip.kps(f), oclip.coverage(f),
                                             prev_o, cid_o, threshold=25, coverage_threshold=cov_thr, max_tracks=8192,
                                             lk_status=None if lk is None else lk[0], lk_pts=None if lk is None else lk[1])
        r = ref.extract_frame(W, H, flags[f], grey[f], rclip.grid(f), rclip.hops(f), rclip.kps(f), rclip.coverage(f), prev_r, cid_r,
                              threshold=25, coverage_threshold=cov_thr, lk_calls=[] if lk is None else [lk], has_prev=f > 0,
                              variant="canon")
        assert r["consistent"], f
        assert r["lk_calls"] == (0 if lk is None else 1), (f, r["lk_calls"])
        if lk is not None:   # the reference asked LK for exactly the points the hand-over convention lists, in that order
            assert np.array_equal(r["lk_last_points"], (np.stack([prev_o["pt_x"][cov], prev_o["pt_y"][cov]], 1) if is_p else
                                                        np.stack([prev_o["pt_x"], prev_o["pt_y"]], 1)).astype(np.float32)), f
        assert cid_o == r["current_id"], (f, cid_o, r["current_id"])
        _assert_tables_equal(got, r["tracks"], f)
        n_backfill += int((((got["flags"] & T.TRACK_COVERAGE) != 0) & (got["q_indx"] < 0)).sum())
        n_cov_carried += int((((got["flags"] & T.TRACK_COVERAGE) != 0) & (got["q_indx"] >= 0)).sum())
        n_iframe_carried += 0 if is_p or f == 0 else len(got)
        # the reference pushes no descriptor for back-fill features (the inner `descriptors` shadows the argument, :421)
        assert r["n_descriptors"] == len(got) - int((((got["flags"] & T.TRACK_COVERAGE) != 0) & (got["q_indx"] < 0)).sum()), f
        n_short_desc += r["n_descriptors"] != len(got)
        prev_o, prev_r, cid_r = got, r["tracks"], r["current_id"]
    assert len(prev_o) > 20
    if cov_thr > 0.9:
        assert n_backfill > 0 and n_cov_carried > 0 and n_short_desc > 0
    if iframe_at is not None:
        assert n_iframe_carried > 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_extractor_oracle_equals_unmodified_reference_on_tie_free_tables(orc, ref, seed):
    """The 'plain' build keeps the reference's std::sort. On a previous table whose (age, popcount) pairs are all distinct
    the comparator of MOVExtractor.cc:249-252 is a strict total order, so every sort gives the same permutation: the
    oracle must equal the reference exactly as written, including the in-place order it leaves in prev->mvVF."""
    W, H, NF, K = 160, 112, 6, 2
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0580 + seed, fx=80.0, fy=80.0)
    recs, off, flags = synth.make_records(spec)
    grey = synth.make_grey(spec)
    rclip = ref.Clip(W, H, recs, off, flags, grey=grey)
    oclip = orc.Clip(W, H, recs, off, flags, K)
    rng = np.random.Generator(np.random.PCG64(0x2EC0 + seed))
    prev, cid = np.zeros(0, T.TRACK), 0
    checked = 0
    for f in range(NF):
        if f > 0:
            # tie-free variant of the table: random distinct ages, shuffled order
            tf = prev.copy()
            tf["age"] = rng.permutation(len(tf)) + 1
            tf = tf[rng.permutation(len(tf))]
            got, sp_o, cid_o, _ = orc.extract_frame(W, H, flags[f], grey[f], oclip.grid(f), oclip.hops(f), oclip.kps(f), oclip.coverage(f),
                                                    tf, cid, threshold=25, coverage_threshold=0.2, max_tracks=8192)
            r = ref.extract_frame(W, H, flags[f], grey[f], rclip.grid(f), rclip.hops(f), rclip.kps(f), rclip.coverage(f), tf, cid,
                                  threshold=25, coverage_threshold=0.2, variant="plain")
            assert r["consistent"] and cid_o == r["current_id"]
            _assert_tables_equal(got, r["tracks"], f)
            _assert_tables_equal(sp_o, r["sorted_prev"], (f, "sorted prev"))
            checked += len(got)
        prev, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], oclip.grid(f), oclip.hops(f), oclip.kps(f), oclip.coverage(f),
                                            prev, cid, threshold=25, coverage_threshold=0.2, max_tracks=8192)
    assert checked > 100


def reloc_seeds_host_side(pts_kf, track_ids, status, pts_out, W, H, reloc_distance):
    """The host half of the lost-relocalisation branch (src/MOVExtractor.cc:199-215): status, image bounds and the distance
    test on the LK output, in the reference's types (cv::norm in double). -> RELOC_SEED array for the device half."""
    th = reloc_distance * np.sqrt(float(H * H + W * W))
    seeds = []
    for i in range(len(pts_kf)):
        x, y = np.float32(pts_out[i][0]), np.float32(pts_out[i][1])
        if status[i] == 0 or x < 0 or y < 0 or x >= W or y >= H:
            continue
        dx, dy = np.float32(x - np.float32(pts_kf[i][0])), np.float32(y - np.float32(pts_kf[i][1]))
        if np.sqrt(float(dx) * float(dx) + float(dy) * float(dy)) < th:
            seeds.append((int(track_ids[i]), i, x, y))
    return np.array(seeds, T.RELOC_SEED) if seeds else np.zeros(0, T.RELOC_SEED)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_lost_relocalisation_oracle_equals_reference(orc, ref, seed):
    """prev->mLost (MOVExtractor.cc:161-243): keyframe map points carried by LK seed tracks ahead of the propagated ones.
    The reference runs the whole branch; the oracle gets the seeds that passed the host-side tests."""
    W, H, NF, K = 160, 112, 4, 2
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0590 + seed, fx=80.0, fy=80.0)
    recs, off, flags = synth.make_records(spec)
    grey = synth.make_grey(spec)
    rclip = ref.Clip(W, H, recs, off, flags, grey=grey)
    oclip = orc.Clip(W, H, recs, off, flags, K)
    rng = np.random.Generator(np.random.PCG64(0x2E90 + seed))
    prev, cid = np.zeros(0, T.TRACK), 0
    for f in range(NF - 1):
        prev, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], oclip.grid(f), oclip.hops(f), oclip.kps(f), oclip.coverage(f),
                                            prev, cid, threshold=25, coverage_threshold=0.95, max_tracks=8192)
    f = NF - 1
    n_kf = 120
    in_view = (rng.random(n_kf) < 0.8).astype(np.uint8)
    proj = np.stack([rng.uniform(0, W, n_kf), rng.uniform(0, H, n_kf)], 1).astype(np.float32)
    ids = rng.integers(1, 400, n_kf).astype(np.int32)
    sel = np.nonzero(in_view)[0]                       # the points the reference hands to LK, in order (:171-192)
    st = (rng.random(len(sel)) > 0.2).astype(np.uint8)
    out = proj[sel] + rng.normal(0, 12.0, (len(sel), 2)).astype(np.float32)
    order = _sorted_order(prev)
    cov = [i for i in order if prev["flags"][i] & T.TRACK_COVERAGE]
    lk_cov = _lk_for(rng, np.stack([prev["pt_x"][cov], prev["pt_y"][cov]], 1), W, H) if cov else None
    r = ref.extract_frame(W, H, flags[f], grey[f], rclip.grid(f), rclip.hops(f), rclip.kps(f), rclip.coverage(f), prev, cid,
                          threshold=25, coverage_threshold=0.95, relocalization_distance=0.1, lost=True, kf_points=(in_view, proj, ids),
                          lk_calls=[(st, out)] + ([lk_cov] if cov else []), variant="canon")
    assert r["consistent"] and r["lk_calls"] == (2 if cov else 1)
    seeds = reloc_seeds_host_side(proj[sel], ids[sel], st, out, W, H, 0.1)
    assert 0 < len(seeds) < len(sel)                   # the distance test rejected some
    got, _, cid_o, _ = orc.extract_frame(W, H, flags[f], grey[f], oclip.grid(f), oclip.hops(f), oclip.kps(f), oclip.coverage(f), prev, cid,
                                         threshold=25, coverage_threshold=0.95, max_tracks=8192, lk_status=lk_cov[0] if cov else None,
                                         lk_pts=lk_cov[1] if cov else None, reloc=seeds)
    assert cid_o == r["current_id"]
    _assert_tables_equal(got, r["tracks"], "lost")
    n_seeded = int(((got["age"] == 0) & (got["q_indx"] >= 0)).sum())
    assert 0 < n_seeded <= len(seeds)


def test_plain_and_canon_reference_builds_differ_only_by_tie_order(ref):
    """Documents what the canonicalisation changes: on a table WITH ties both builds keep the same multiset of tracks in
    sorted_prev; the order of ties is std::sort's business in the plain build."""
    W, H = 96, 64
    rng = np.random.Generator(np.random.PCG64(0x2EB0))
    tr = np.zeros(200, T.TRACK)
    tr["age"] = rng.integers(0, 3, len(tr))
    tr["track_id"] = np.arange(1, len(tr) + 1)
    tr["pt_x"], tr["pt_y"] = rng.uniform(8, W - 9, len(tr)), rng.uniform(8, H - 9, len(tr))
    tr["mb"]["w"], tr["mb"]["h"] = 16, 16
    tr["desc"][:, 0] = rng.integers(0, 16, len(tr))    # popcounts 0..4: many ties
    grid = np.full((H, W, 4), -1, np.int32)
    flat = np.full((H, W), 128, np.uint8)
    out = {}
    for v in ("plain", "canon"):
        out[v] = ref.extract_frame(W, H, T.FRAME_P | T.FRAME_MV, flat, grid, np.zeros(0, T.HOP), np.zeros(0, T.RECT), 1.0, tr, 200, variant=v)
    key = lambda t: sorted(map(int, t["track_id"]))
    assert key(out["plain"]["sorted_prev"]) == key(out["canon"]["sorted_prev"]) == list(range(1, 201))
    canon = out["canon"]["sorted_prev"]
    assert list(canon["track_id"]) == [int(tr["track_id"][i]) for i in _sorted_order(tr)]
    for v in out.values():       # both are sorted by the comparator
        sp = v["sorted_prev"]
        k = list(zip(-sp["age"].astype(int), -_popc(sp)))
        assert k == sorted(k)


# -------------------------------------------------------------------------------------------------- MOVMatcher -----
def test_matcher_oracle_equals_reference_header(orc, ref):
    rng = np.random.Generator(np.random.PCG64(0x2EA0))
    for trial in range(20):
        nt, nm = int(rng.integers(1, 300)), int(rng.integers(1, 400))
        tr = np.zeros(nt, T.TRACK)
        tr["track_id"] = rng.integers(1, 200, nt)          # duplicates: first index wins in mvVFMap
        tr["pt_x"], tr["pt_y"] = rng.uniform(0, 640, nt), rng.uniform(0, 480, nt)
        tr["mb"]["w"] = 16
        mp = np.zeros(nm, T.MAP_POINT)
        mp["track_id"] = rng.integers(1, 260, nm)          # duplicates: last map point wins
        mp["flags"] = np.where(rng.random(nm) < 0.1, T.MP_BAD, 0)
        proj = np.zeros(nm, T.PROJECTION)
        proj["in_view"] = rng.random(nm) < 0.7
        proj["depth"] = rng.uniform(1, 30, nm)
        init = np.where(rng.random(nt) < 0.2, rng.integers(0, nm, nt), -1).astype(np.int32)
        for far, th in ((False, 0.0), (True, 15.0)):
            wn, wm = ref.search_by_video_feature(tr, mp, proj, init, far, th)
            gn, gm = orc.search_by_video_feature(tr, mp, proj, init, far, th)
            assert gn == wn and np.array_equal(gm, wm), (trial, far)
        kf = mp.copy()
        kf["flags"] = np.where(rng.random(nm) < 0.15, T.MP_NULL, kf["flags"])
        wn, wm = ref.search_by_keyframe(tr, kf)
        gn, gm = orc.search_by_keyframe(tr, kf)
        assert gn == wn and np.array_equal(gm, wm), trial
        f2 = np.zeros(int(rng.integers(1, 300)), T.TRACK)
        f2["track_id"] = rng.integers(1, 200, len(f2))
        f2["pt_x"], f2["pt_y"] = rng.uniform(0, 640, len(f2)), rng.uniform(0, 480, len(f2))
        pm = rng.uniform(0, 640, (nt, 2)).astype(np.float32)
        wn, wm, wp = ref.search_for_initialization(tr, f2, pm)
        gn, gm, gp = orc.search_for_initialization(tr, f2, pm)
        assert gn == wn and np.array_equal(gm, wm) and np.array_equal(gp, wp), trial
