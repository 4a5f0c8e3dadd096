#!/usr/bin/env python
"""Host->device copy bandwidth from plain pinned memory against WRITE-COMBINED pinned memory (cudaHostAllocWriteCombined), and with
the copy split over two streams, no kernels running: is there anything left on the PCIe link that the end-to-end path (51 of
55 GB/s) does not use?  python scripts/probes/h2d_wc_probe.py   (one JSON line)"""
import ctypes as C
import json
import os

rt = C.CDLL("libcudart.so.12")
MB = int(os.environ.get("PROBE_MB", 352))
N = MB << 20


def ck(rc, what):
    if rc:
        raise RuntimeError("%s: cudaError %d" % (what, rc))


ck(rt.cudaSetDevice(0), "set device")
dev = C.c_void_p()
ck(rt.cudaMalloc(C.byref(dev), C.c_size_t(N)), "malloc")
streams = [C.c_void_p(), C.c_void_p()]
for s in streams:
    ck(rt.cudaStreamCreate(C.byref(s)), "stream")
e0, e1 = C.c_void_p(), C.c_void_p()
rt.cudaEventCreate(C.byref(e0))
rt.cudaEventCreate(C.byref(e1))


def measure(flags, split):
    host = C.c_void_p()
    ck(rt.cudaHostAlloc(C.byref(host), C.c_size_t(N), C.c_uint(flags)), "host alloc")
    C.memset(host, 0x5a, N)
    best = 0.0
    for rep in range(6):
        ck(rt.cudaDeviceSynchronize(), "sync")
        rt.cudaEventRecord(e0, streams[0])
        reps = 8
        for _ in range(reps):
            if split:
                h = N // 2
                rt.cudaMemcpyAsync(dev, host, C.c_size_t(h), 1, streams[0])
                rt.cudaMemcpyAsync(C.c_void_p(dev.value + h), C.c_void_p(host.value + h), C.c_size_t(N - h), 1, streams[1])
            else:
                rt.cudaMemcpyAsync(dev, host, C.c_size_t(N), 1, streams[0])
        ck(rt.cudaStreamSynchronize(streams[1]), "sync 1")
        rt.cudaEventRecord(e1, streams[0])
        ck(rt.cudaEventSynchronize(e1), "sync e1")
        ms = C.c_float()
        rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
        if rep:
            best = max(best, reps * N / 1e9 / (ms.value / 1e3))
    rt.cudaFreeHost(host)
    return round(best, 2)


def measure_cycling(n_buf, parts):
    """n_buf distinct pinned buffers copied in turn (bench.py's end-to-end region reads a different host window every step), each as
    `parts` copies (records / meta / grey planes are separate buffers there)."""
    hosts = []
    for _ in range(n_buf):
        h = C.c_void_p()
        ck(rt.cudaHostAlloc(C.byref(h), C.c_size_t(N), C.c_uint(0)), "host alloc")
        C.memset(h, 0x3c, N)
        hosts.append(h)
    best = 0.0
    for rep in range(4):
        ck(rt.cudaDeviceSynchronize(), "sync")
        rt.cudaEventRecord(e0, streams[0])
        for h in hosts:
            step = N // parts
            for q in range(parts):
                rt.cudaMemcpyAsync(C.c_void_p(dev.value + q * step), C.c_void_p(h.value + q * step), C.c_size_t(step), 1, streams[0])
        rt.cudaEventRecord(e1, streams[0])
        ck(rt.cudaEventSynchronize(e1), "sync e1")
        ms = C.c_float()
        rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
        if rep:
            best = max(best, n_buf * (N // parts) * parts / 1e9 / (ms.value / 1e3))
    for h in hosts:
        rt.cudaFreeHost(h)
    return round(best, 2)


out = {"probe": "host->device copy, %d MB, best of 5 x 8 copies" % MB,
       "pinned_gbs": measure(0, False), "write_combined_gbs": measure(4, False),
       "pinned_two_streams_gbs": measure(0, True), "write_combined_two_streams_gbs": measure(4, True)}
NB = int(os.environ.get("PROBE_BUFFERS", 12))
out["pinned_%d_distinct_buffers_gbs" % NB] = measure_cycling(NB, 1)
out["pinned_%d_distinct_buffers_3_copies_each_gbs" % NB] = measure_cycling(NB, 3)
print(json.dumps(out), flush=True)
