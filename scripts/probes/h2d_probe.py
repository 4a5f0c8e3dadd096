#!/usr/bin/env python
"""Plain pinned host->device copy bandwidth of this box with NO kernels running, one process per GPU (torchrun), so that the
end-to-end scaling of bench.py can be separated into "what the box can feed N GPUs" and "what our pipeline does with it".
Every rank copies its own pinned buffer (the size of one bench step's inputs, ~400 MB) to its GPU back to back for a few
seconds; ranks start together (barrier) and the aggregate is the sum of the per-rank rates over the common interval.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/probes/h2d_probe.py
Prints one JSON line (rank 0)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = int(os.environ.get("PROBE_MB", 400))
SECONDS = float(os.environ.get("PROBE_SECONDS", 4))
src = torch.empty(MB << 20, dtype=torch.uint8, pin_memory=True)
src.random_(0, 255)
dst = torch.empty_like(src, device="cuda")
for _ in range(3):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < SECONDS:
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
        n += 1
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
gbs = n * src.numel() / 1e9 / (e0.elapsed_time(e1) / 1e3)
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    per = [float(x[0]) for x in allr]
    numa = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    print(json.dumps({"probe": "pinned host->device copy, no kernels", "n_gpus": world, "buffer_mb": MB, "per_gpu_gbs": [round(p, 1) for p in per],
                      "aggregate_gbs": round(sum(per), 1), "host_cpus": os.cpu_count(), "numa_nodes": len(numa)}), flush=True)
if world > 1:
    dist.destroy_process_group()
