"""Error behaviour of the C ABI (include/movfe.h): every call returns a negative MOVFE_E_* code with a message instead of
crashing or silently corrupting state, and the context stays usable afterwards - the same contract as the reference's
interfaces, which return -1 / 0 / nullptr and never throw."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import pack_streams

pytestmark = pytest.mark.gpu


def _err(fn, code):
    with pytest.raises(lib.MovfeError, match=r"movfe error %d:" % code):
        fn()


def test_call_order_and_capacity_errors_leave_the_context_usable(orc):
    W, H, F, K = 160, 112, 3, 1
    spec = synth.Spec(W, H, n_frames=12, refs=K + 1, seed=0x5EED00E0, fx=80.0, fy=80.0)
    stream, grey = synth.make_records(spec), synth.make_grey(spec)
    ctx = lib.Context(1, W, H, max_records_per_frame=400, max_ref=K, window_frames=F, max_tracks=64, max_map_points=8)
    ring = 2 * F + K + 1
    E_INVALID, E_CAPACITY, E_STATE = -1, -3, -4

    _err(lambda: ctx.raster(0, F), E_STATE)                                     # nothing pushed yet
    _err(lambda: ctx.extract(0, 1), E_STATE)                                    # no raster window
    _err(lambda: ctx.track_poses(0, 1), E_STATE)                                # nothing extracted
    r, o, fl = pack_streams([stream], 12, 0, ring + 1)
    _err(lambda: ctx.push_frames(ring + 1, r, o, fl, grey[None, :ring + 1]), E_CAPACITY)   # more frames than the ring holds
    assert ctx.frames_pushed() == 0

    r, o, fl = pack_streams([stream], 12, 0, F + K + 1)
    ctx.push_frames(F + K + 1, r, o, fl, grey[None, :F + K + 1])
    _err(lambda: ctx.raster(0, F + 1), E_CAPACITY)                              # longer than window_frames
    _err(lambda: ctx.raster(4, F), E_STATE)                                     # frames 5, 6 were not pushed
    ctx.raster(0, F)
    _err(lambda: ctx.grid(0, F), E_STATE)                                       # outside the raster window
    _err(lambda: ctx.grid(1, 0), E_INVALID)                                     # no such stream
    _err(lambda: ctx.extract(1, 1), E_STATE)                                    # frames are consumed in order
    _err(lambda: ctx.extract(0, F + 1), E_STATE)                                # beyond the raster window
    _err(lambda: ctx.set_tracks(0, np.zeros(65, T.TRACK), 0), E_CAPACITY)
    _err(lambda: ctx.set_map_points(0, np.zeros(9, T.MAP_POINT), 0), E_CAPACITY)
    _err(lambda: ctx.set_map_points(1, np.zeros(1, T.MAP_POINT), 0), E_INVALID)
    ctx.extract(0, F)
    _err(lambda: ctx.tracks(0, F), E_STATE)                                     # not extracted yet
    _err(lambda: ctx.track_poses(0, F + 1), E_STATE)
    q = np.array([(3, 1.0, 1.0, 1.0)], T.AREA_QUERY)
    _err(lambda: ctx.features_in_area(np.zeros((2, 2), np.float32), [0, 2], np.zeros((1, 64 * 48 + 1), np.int32),
                                      np.zeros(2, np.int32), q, 4), E_INVALID)   # query names a set that does not exist

    # after all of that the context still produces the oracle's tables (capacity 64: both sides stop at the cap)
    clip = orc.Clip(W, H, *stream, K)
    prev, cid = np.zeros(0, T.TRACK), 0
    for f in range(F):
        want, _, cid, _ = orc.extract_frame(W, H, stream[2][f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f),
                                            clip.coverage(f), prev, cid, max_tracks=64)
        assert ctx.tracks(0, f).tobytes() == want.tobytes(), f
        prev = want
    ctx.close()


def test_create_rejects_bad_configurations():
    for kw in (dict(n_streams=0), dict(width=8), dict(max_ref=11), dict(window_frames=0), dict(max_tracks=70000)):
        args = dict(n_streams=1, width=64, height=48, max_ref=1, window_frames=2, max_tracks=64)
        args.update(kw)
        with pytest.raises(lib.MovfeError, match="movfe_create failed"):
            lib.Context(args.pop("n_streams"), args.pop("width"), args.pop("height"), **args)
