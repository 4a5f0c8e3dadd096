"""Multi-GPU host logic on CPU: stream -> rank partition and the pose gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

from movfe import shard, types as T


def test_partition_is_a_bijection():
    for n, world in [(64, 1), (64, 2), (64, 8), (7, 4), (3, 8), (1024, 8)]:
        seen = []
        for r in range(world):
            ids = shard.streams_of_rank(n, r, world)
            assert all(shard.owner(s, world) == r for s in ids)
            assert [shard.local_index(s, world) for s in ids] == list(range(len(ids)))
            seen += ids
        assert sorted(seen) == list(range(n))
        sizes = [len(shard.streams_of_rank(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _fake_result(stream, F):
    """Deterministic per-stream result standing for what the rank's context would return."""
    p = np.zeros(F, T.POSE)
    for f in range(F):
        p[f]["R"] = np.arange(9) * 0.5 + stream + 0.01 * f
        p[f]["t"] = np.array([stream, f, stream * f], np.float64)
    return p, (np.arange(F) + 10 * stream).astype(np.int32)


def _worker(rank, world, port, n_streams, F, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard.streams_of_rank(n_streams, rank, world)
    lp = np.zeros((len(ids), F), T.POSE)
    li = np.zeros((len(ids), F), np.int32)
    for k, s in enumerate(ids):
        lp[k], li[k] = _fake_result(s, F)
    poses, inl = shard.gather_poses(lp, li, n_streams, dist)
    np.save(os.path.join(out_dir, "poses_%d.npy" % rank), poses)
    np.save(os.path.join(out_dir, "inl_%d.npy" % rank), inl)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [8, 5])
def test_gather_poses_world2_gloo(tmp_path, n_streams):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    F, world = 3, 2
    mp.spawn(_worker, args=(world, port, n_streams, F, str(tmp_path)), nprocs=world, join=True)
    want_p = np.zeros((n_streams, F), T.POSE)
    want_i = np.zeros((n_streams, F), np.int32)
    for s_ in range(n_streams):
        want_p[s_], want_i[s_] = _fake_result(s_, F)
    for r in range(world):
        got_p = np.load(tmp_path / ("poses_%d.npy" % r))
        got_i = np.load(tmp_path / ("inl_%d.npy" % r))
        assert got_p.tobytes() == want_p.tobytes()
        assert np.array_equal(got_i, want_i)
