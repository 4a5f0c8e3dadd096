"""GPU parity of the pyramidal Lucas-Kanade tracker (movfe_lk, csrc/lk.cu) - SURVEY.md 8f item 3: cv::calcOpticalFlowPyrLK as the
reference calls it (src/MOVExtractor.cc:91-92,196-197,347-348; src/Frame.cc:305). Against OpenCV's own outputs
(tests/golden/lk_golden.npz, made with cv2 by tests/golden/make_lk_golden.py): status flags identical, positions within 5e-3 px,
min-eigenvalues within 1e-6; against the numpy oracle (oracle/lk.py) on fresh image pairs: status identical, 5e-3 px. Tolerances
are written here because the window sums are formed in a different order than OpenCV's (float32, last bits)."""
import os

import numpy as np
import pytest

from movfe import lib

from oracle import lk as olk

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lk_golden.npz"))
POS_TOL = 5e-3


def test_against_opencv_golden():
    for k in range(3):
        prev, nxt, pts, win = G["prev%d" % k], G["next%d" % k], G["pts%d" % k], int(G["win%d" % k])
        H, W = prev.shape
        ctx = lib.Context(1, W, H, has_grey=False)
        out, st, err = ctx.lk(prev, nxt, pts, [0, len(pts)], win=win)
        ctx.close()
        assert np.array_equal(st, G["status%d" % k]), (k, np.nonzero(st != G["status%d" % k])[0])
        ok = st == 1
        assert np.max(np.linalg.norm(out[ok] - G["out%d" % k][ok], axis=1)) <= POS_TOL, k
        assert np.max(np.abs(err - G["err%d" % k])) <= 1e-6, k


def test_batched_pairs_against_oracle():
    """several image pairs and point lists of different length in one call (an I-frame carry-over hands in every track of a stream)"""
    rng = np.random.default_rng(21)
    W, H, NP = 256, 192, 3
    prevs, nxts, pts, off = [], [], [], [0]
    for p in range(NP):
        img = rng.random((H + 16, W + 16))
        for _ in range(3):          # smooth texture: box blur passes
            img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5.0
        img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
        dx, dy = int(rng.integers(-4, 5)), int(rng.integers(-3, 4))
        prevs.append(img[8:8 + H, 8:8 + W])
        nxts.append(img[8 + dy:8 + dy + H, 8 + dx:8 + dx + W])
        n = [40, 1, 25][p]
        pts.append(np.stack([rng.uniform(0, W - 1, n), rng.uniform(0, H - 1, n)], 1).astype(np.float32))
        off.append(off[-1] + n)
    ctx = lib.Context(1, W, H, has_grey=False)
    out, st, err = ctx.lk(np.stack(prevs), np.stack(nxts), np.concatenate(pts), off)
    ctx.close()
    for p in range(NP):
        want, wst, werr = olk.track(prevs[p], nxts[p], pts[p])
        sl = slice(off[p], off[p + 1])
        assert np.array_equal(st[sl], wst), p
        ok = wst == 1
        if ok.any():
            assert np.max(np.linalg.norm(out[sl][ok] - want[ok], axis=1)) <= POS_TOL, p
        assert np.max(np.abs(err[sl] - werr)) <= 1e-6, p
    assert st.sum() > 40


def test_device_resident_carry_over(orc):
    """movfe_lk_carry: the batched path follows coverage tracks (src/MOVExtractor.cc:337-377) and carries every track over an intra
    picture in mid-stream (:81-120) without the host - LK on the ring's grey planes, results installed where movfe_set_lk_results
    would put them. Against the oracle chain fed with the numpy LK oracle's results: same tracks, ids, ages, order, blocks and
    descriptors in every frame up to and including the intra picture; positions of LK-carried tracks within 5e-3 px."""
    from movfe import synth, types as T
    from gpu_util import pack_streams
    W, H, NF, K, thr, cov_thr, iframe_at = 320, 240, 5, 2, 25, 0.95, 4
    specs = [synth.Spec(W, H, n_frames=NF + K + 1, refs=K + 1, seed=0x5EED0B30 + s, fx=160.0, fy=160.0, phase=0.2 * s) for s in range(2)]
    streams, greys, want = [], [], []
    for sp in specs:
        recs, off, flags = synth.make_records(sp)
        flags = flags.copy()
        flags[iframe_at] &= ~np.uint8(T.FRAME_P)
        grey = synth.make_grey(sp)
        clip = orc.Clip(W, H, recs, off, flags, K)
        prev, cid, tabs = np.zeros(0, T.TRACK), 0, []
        for f in range(NF):
            lkr = None
            if f > 0 and len(prev):
                if not (flags[f] & T.FRAME_P):
                    pts = np.stack([prev["pt_x"], prev["pt_y"]], 1)
                else:       # coverage tracks in the sorted order of the previous table (age desc, descriptor popcount desc, stable)
                    pc = np.array([sum(bin(int(w)).count("1") for w in t["desc"]) for t in prev], np.int64)
                    order = sorted(range(len(prev)), key=lambda i: (-int(prev["age"][i]), -int(pc[i])))
                    cov = [i for i in order if prev["flags"][i] & T.TRACK_COVERAGE]
                    pts = np.stack([prev["pt_x"][cov], prev["pt_y"][cov]], 1) if cov else np.zeros((0, 2), np.float32)
                if len(pts):
                    out, st, _ = olk.track(grey[f - 1], grey[f], pts)
                    lkr = (st, out)
                else:
                    lkr = (np.zeros(0, np.uint8), np.zeros((0, 2), np.float32))
            t, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev, cid, threshold=thr,
                                             coverage_threshold=cov_thr, max_tracks=4096, lk_status=None if lkr is None else lkr[0],
                                             lk_pts=None if lkr is None else lkr[1])
            tabs.append(t)
            prev = t
        streams.append((recs, off, flags))
        greys.append(grey)
        want.append(tabs)
    n_cov = sum(int(((t["flags"] & T.TRACK_COVERAGE) != 0).sum()) for t in want[0][:iframe_at])
    assert n_cov > 0 and len(want[0][iframe_at]) > 50      # coverage tracks were followed, and the intra picture carried tracks
    ctx = lib.Context(2, W, H, max_records_per_frame=4800, max_ref=K, window_frames=NF, max_tracks=4096, has_grey=True, express_threshold=thr,
                      coverage_threshold=cov_thr)
    n_all = NF + K + 1
    r, o, fl = pack_streams(streams, n_all, 0, n_all)
    ctx.push_frames(n_all, r, o, fl, np.stack([g[:n_all] for g in greys]))
    ctx.raster(0, NF)
    for f in range(NF):
        if f > 0:
            ctx.lk_carry(f)
        ctx.extract(f, 1)
    for s in range(2):
        for f in range(NF):
            got, w = ctx.tracks(s, f), want[s][f]
            assert len(got) == len(w), (s, f, len(got), len(w))
            for name in ("track_id", "age", "q_indx", "flags", "mb", "desc"):
                assert got[name].tobytes() == w[name].tobytes(), (s, f, name)
            assert np.max(np.abs(got["pt_x"] - w["pt_x"]), initial=0) <= POS_TOL and np.max(np.abs(got["pt_y"] - w["pt_y"]), initial=0) <= POS_TOL, (s, f)
    assert ctx.dropped_lk_tracks() == 0
    ctx.close()
