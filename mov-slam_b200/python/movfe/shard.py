"""Stream sharding over the GPUs of one box: streams are the independent unit, stream s belongs to
rank s mod G, and there is no collective on the data path. The only cross-GPU step is the gather of per-stream poses
(and inlier counts) to the caller, done on the host side with torch.distributed (NCCL on the GPU box, gloo in the CPU
tests)."""
import numpy as np


def streams_of_rank(n_streams, rank, world):
    """Global stream ids owned by `rank` (static partition s -> s mod world), ascending."""
    return list(range(rank, n_streams, world))


def local_index(stream, world):
    """Index of a global stream inside its owner's context."""
    return stream // world


def owner(stream, world):
    return stream % world


def gather_poses(local_poses, local_inliers, n_streams, dist=None):
    """All ranks contribute poses [n_local, F] (dtype types.POSE) and inliers [n_local, F]; every rank gets the arrays of
    all n_streams streams in global stream order. `dist` is torch.distributed (initialised) or None for one rank."""
    from . import types as T
    local_poses = np.ascontiguousarray(local_poses, T.POSE)
    local_inliers = np.ascontiguousarray(local_inliers, np.int32)
    if dist is None or dist.get_world_size() == 1:
        return local_poses.copy(), local_inliers.copy()
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    F = local_poses.shape[1]
    per = (n_streams + world - 1) // world            # ranks own per or per-1 streams; pad to `per`
    flat = np.zeros((per, F, 13), np.float64)         # 9 R + 3 t + inlier count
    n_loc = len(streams_of_rank(n_streams, rank, world))
    assert local_poses.shape[0] == n_loc
    flat[:n_loc, :, :9] = local_poses["R"]
    flat[:n_loc, :, 9:12] = local_poses["t"]
    flat[:n_loc, :, 12] = local_inliers
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(flat).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    poses = np.zeros((n_streams, F), T.POSE)
    inl = np.zeros((n_streams, F), np.int32)
    for r in range(world):
        a = parts[r].cpu().numpy()
        ids = streams_of_rank(n_streams, r, world)
        poses["R"][ids] = a[:len(ids), :, :9]
        poses["t"][ids] = a[:len(ids), :, 9:12]
        inl[ids] = a[:len(ids), :, 12].astype(np.int32)
    return poses, inl
