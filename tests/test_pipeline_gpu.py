"""GPU parity of the PIPELINED front-end: windows are enqueued back to back with no synchronising download in between, so
the raster of window k+1 (raster stream) really runs beside the propagation of window k (primary stream) and the pose
chain of window k (pose stream), the push of window k+2 reuses ring slots and staging buffers, and the raster results
alternate between the two buffers. Only the LAST window is read back: every frame's table depends on all earlier frames,
so a race anywhere in the chain shows up there. Bars: track tables bit-exact, poses within 1e-5 relative."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import assert_tracks_equal, oracle_tracks, pack_streams

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("serial_raster,env,output_grid", [(False, {}, True), (True, {}, True), (False, {"MOVFE_POSE_SPLIT": "1"}, True),
                                                           (False, {"MOVFE_PDL": "1"}, True), (False, {"MOVFE_EXTRACT_GROUPS": "3"}, True), (False, {"MOVFE_EXTRACT_GROUPS": "1"}, False),
                                                           (False, {}, False), (True, {}, False), (False, {"MOVFE_PDL": "1"}, False),
                                                           (False, {"MOVFE_CAND_LANE": "0"}, False), (False, {"MOVFE_POSE_V1": "1"}, False), (False, {"MOVFE_CAND_LANE": "0", "MOVFE_CAND_PIPE": "0"}, True),
                                                           (False, {"MOVFE_HOPS_PRIO": "0"}, False), (False, {"MOVFE_INGEST_STREAM": "1"}, False),
                                                           (False, {"MOVFE_INGEST_STREAM": "2", "MOVFE_RING_EXTRA": "1"}, True)])
def test_pipelined_windows_match_oracle(orc, serial_raster, env, output_grid, monkeypatch):
    """output_grid False = MOVFE_CFG_NO_GRID, the fused mode bench.py's headline uses: slots resolved from the per-tile hop
    queues, same tables bit for bit."""
    for k, v in env.items():      # development switches of the library, read at movfe_create: every path stays parity-tested
        monkeypatch.setenv(k, v)
    W, H, F, K, NW, S, NB = 640, 480, 8, 3, 6, 12, 3
    LA = K + 1
    n_frames = F * NW + LA
    specs = [synth.Spec(W, H, n_frames=n_frames, refs=K + 1, seed=0x5EED0040 + 31 * b, phase=0.41 * b) for b in range(NB)]
    clips = [synth.make_records(sp) for sp in specs]
    greys = [synth.make_grey(sp) for sp in specs]
    per_stream = [clips[s % NB] for s in range(S)]
    grey = [greys[s % NB] for s in range(S)]
    want = [oracle_tracks(orc, clips[b], W, H, K, grey=greys[b], max_tracks=8192) for b in range(NB)]
    maps = [synth.map_from_tracks(specs[b], want[b][0], synth.pose_at(specs[b], 0)) for b in range(NB)]
    cam, pp = specs[0].camera(), T.pose_params()

    ctx = lib.Context(S, W, H, max_records_per_frame=4800, max_ref=K, window_frames=F, max_tracks=8192, max_map_points=2048,
                      has_grey=True, serial_raster=serial_raster, output_grid=output_grid)
    ctx.set_camera(cam, pp, 0.5)
    off = np.cumsum([0] + [len(maps[s % NB]) for s in range(S)]).astype(np.int64)   # all streams' maps in one call
    ctx.set_map_points_batch(np.concatenate([maps[s % NB] for s in range(S)]), off, [len(maps[s % NB]) // 2 for s in range(S)], 2048)
    for s in range(S):
        ctx.set_pose(s, synth.pose_struct(synth.pose_at(specs[s % NB], 0)))

    def push(f0, f1):
        r, o, fl = pack_streams(per_stream, n_frames, f0, f1)
        ctx.push_frames(f1 - f0, r, o, fl, np.stack([grey[s][f0:f1] for s in range(S)]))

    push(0, F + LA)
    for k in range(NW):                      # nothing in this loop waits for the GPU
        first = F * k
        ctx.raster(first, F)
        ctx.extract(first, F)
        ctx.track_poses(first, F)
        if k + 1 < NW:
            push(F * (k + 1) + LA, F * (k + 2) + LA)
    last = F * (NW - 1)
    P, ninl = ctx.poses(last, F)
    for s in range(S):
        b = s % NB
        for f in range(last, last + F):
            assert_tracks_equal(ctx.tracks(s, f), want[b][f], (s, f))
    worst = 0.0
    for b in range(NB):
        r, o, fl = clips[b]
        # the whole clip, look-ahead frames included: they back-fill hops into the last window (VideoDecoder.cc:315-323)
        ref = orc.frontend_run(W, H, r, o, fl, greys[b], None, maps[b],
                               synth.pose_struct(synth.pose_at(specs[b], 0)), cam, pp, max_ref=K, max_tracks=8192,
                               n_kf_points=len(maps[b]) // 2)
        for s in range(b, S, NB):
            for k in range(F):
                assert ninl[s, k] == ref["n_inliers"][last + k], (s, k, ninl[s, k], ref["n_inliers"][last + k])
                for name in ("R", "t"):
                    a, w = P[s, k][name], ref["poses"][last + k][name]
                    worst = max(worst, float(np.max(np.abs(a - w)) / max(1.0, float(np.max(np.abs(w))))))
    assert worst <= 1e-5, worst
    ctx.close()


def test_feature_grid_of_resident_tables(orc):
    """Frame::AssignFeaturesToGrid (Frame.cc:356-388) on the device-resident track table of a frame == the oracle's grid of
    the same keypoints; a radius query through it returns the oracle's list."""
    from gpu_util import run_frontend_clip
    W, H, NF, K = 640, 480, 7, 2
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0051)
    stream, grey = synth.make_records(spec), synth.make_grey(spec)
    tracks, _, ctx = run_frontend_clip([stream], W, H, NF, 4, K, grey=[grey])
    for f in (NF - 1, NF - 3):
        tr = tracks[(0, f)]
        assert len(tr) > 300
        start, items = ctx.track_feature_grid(0, f)
        ws, wi = orc.assign_features_to_grid(tr, W, H)
        assert np.array_equal(start, ws)
        assert np.array_equal(items[:ws[-1]], wi[:ws[-1]]) and (items[ws[-1]:] == -1).all()
        pts = np.stack([tr["pt_x"], tr["pt_y"]], 1)
        q = np.array([(0, 320.0, 240.0, 50.0), (0, 10.0, 470.0, 30.0)], T.AREA_QUERY)
        out, cnt = ctx.features_in_area(pts, [0, len(tr)], start[None], items, q, 2048)
        for k in range(2):
            w = orc.get_features_in_area(tr, W, H, ws, wi, q[k]["x"], q[k]["y"], q[k]["r"])
            assert cnt[k] == len(w) and np.array_equal(out[k, :len(w)], w)
    ctx.close()


def test_device_resident_push_and_two_contexts(orc):
    """movfe_push_frames_device (inputs already in HBM, what bench.py's `value` region uses) gives the same tables as the
    host push, and two contexts alive on one GPU (the bench creates them back to back) do not disturb each other."""
    import torch
    W, H, F, K, NW, S = 320, 240, 4, 2, 3, 2
    LA = K + 1
    n_frames = F * NW + LA
    specs = [synth.Spec(W, H, n_frames=n_frames, refs=K + 1, seed=0x5EED0055 + s, fx=160.0, fy=160.0, phase=0.3 * s) for s in range(S)]
    clips = [synth.make_records(sp) for sp in specs]
    greys = [synth.make_grey(sp) for sp in specs]
    want = [oracle_tracks(orc, clips[s], W, H, K, grey=greys[s], max_tracks=2048) for s in range(S)]
    ctx_a = lib.Context(S, W, H, max_records_per_frame=1300, max_ref=K, window_frames=F, max_tracks=2048, has_grey=True)
    ctx_b = lib.Context(S, W, H, max_records_per_frame=1300, max_ref=K, window_frames=F, max_tracks=2048, has_grey=True)
    keep = []

    def push(ctx, f0, f1, device):
        r, o, fl = pack_streams(clips, n_frames, f0, f1)
        g = np.stack([greys[s][f0:f1] for s in range(S)])
        if not device:
            ctx.push_frames(f1 - f0, r, o, fl, g)
            return
        r = np.ascontiguousarray(r, T.MV_RECORD)              # np.concatenate may hand back a packed (36-byte) dtype
        pad = np.zeros(len(r) * 40 + 16, np.uint8)
        pad[:len(r) * 40] = r.view(np.uint8)
        d = [torch.from_numpy(a).cuda() for a in (pad, o, fl, g)]
        torch.cuda.synchronize()
        keep.append(d)                                         # device inputs stay alive until the run is read back
        ctx.push_frames_device(f1 - f0, d[0].data_ptr(), d[1].data_ptr(), len(r), d[2].data_ptr(), d[3].data_ptr())

    push(ctx_a, 0, F + LA, True)
    push(ctx_b, 0, F + LA, False)
    for k in range(NW):
        for ctx, device in ((ctx_a, True), (ctx_b, False)):
            ctx.raster(F * k, F)
            ctx.extract(F * k, F)
            if k + 1 < NW:
                push(ctx, F * (k + 1) + LA, F * (k + 2) + LA, device)
    for ctx in (ctx_a, ctx_b):
        for s in range(S):
            for f in range(F * (NW - 1), F * NW):
                assert_tracks_equal(ctx.tracks(s, f), want[s][f], (s, f))
        assert ctx.rejected_records() == 0
        ctx.close()
