"""Host side of the batched front-end, CPU only: the decoder pool (mov-slam_b200/shim/decoder_pool.cc: one libav
context per stream on a pool of host threads, side data packed to 16-byte records while it is copied out of the AVFrame, windows in
the layout of movfe_push_frames_packed) against a serial packing of the same clips, and the trajectory writers against the formats of
src/System.cc:363-423 (TUM) and :778-838 (this fork's KITTI form)."""
import os
import subprocess

import numpy as np
import pytest

from movfe import synth, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "mov-slam_b200", "shim")


@pytest.mark.parametrize("streams,frames,threads", [(5, 4, 3), (8, 7, 8), (3, 16, 1)])
def test_pool_windows_equal_serial_packing(tmp_path, streams, frames, threads):
    subprocess.run(["make", "-C", SHIM, "test_pool"], check=True, capture_output=True)
    W, H, NF = 160, 128, 30
    sp = synth.Spec(W, H, n_frames=NF, refs=3, seed=0x5EED0F40)
    recs, off, flags = synth.make_records(sp)
    grey = synth.make_grey(sp)
    d = str(tmp_path)
    np.ascontiguousarray(recs, T.MV_RECORD).tofile(d + "/recs.bin")
    off.tofile(d + "/off.bin")
    flags.tofile(d + "/flags.bin")
    grey.tofile(d + "/grey.bin")
    open(d + "/meta.txt", "w").write("%d %d %d\n" % (W, H, NF))
    r = subprocess.run([os.path.join(SHIM, "test_pool"), d, str(streams), str(frames), str(threads)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    # the shortest stream (the last one, one frame shorter) ends the run for all
    assert int(r.stdout.split()[3]) == NF - streams
    # trajectory files: identity; a quarter turn about z with t = (1, 2, 3); the third frame is lost and not written
    tum = [l.split() for l in open(d + "/traj_tum.txt").read().splitlines()]
    assert len(tum) == 2 and tum[0][0] == "0.000000" and tum[1][0] == "0.033333"
    assert [float(x) for x in tum[0][1:]] == [0, 0, 0, 0, 0, 0, 1]
    # Twc = (R^T, -R^T t): R^T rotates by -90 degrees about z -> q = (0, 0, -sin 45, cos 45), t = -(R^T (1, 2, 3)) = (-2, 1, -3)
    v = [float(x) for x in tum[1][1:]]
    assert np.allclose(v[:3], [-2, 1, -3], atol=1e-6) and np.allclose(v[3:], [0, 0, -np.sqrt(0.5), np.sqrt(0.5)], atol=1e-6)
    assert all(len(x.split(".")[1]) == 9 for x in tum[1][1:])          # setprecision(9), fixed
    kitti = [l.split() for l in open(d + "/traj_kitti.txt").read().splitlines()]
    assert len(kitti) == 2 and len(kitti[1]) == 13 and kitti[1][0] == "1"
    M = np.array([float(x) for x in kitti[1][1:]]).reshape(3, 4)
    assert np.allclose(M[:, :3], [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], atol=1e-6) and np.allclose(M[:, 3], [-2, 1, -3], atol=1e-6)
