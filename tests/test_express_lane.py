"""The thread-level EXPRESS of the propagation kernels (mov-slam_b200/csrc/express_lane.cuh: one lane per block half / per block,
four pixels per 32-bit operation) is plain integer code compiled for host AND device. Here the host compilation is checked
against the oracle's restatement of include/EXPRESS.h:79-192 - centre, candidate descriptor (both halves), compute_express
verdict and birth descriptor - on random blocks of all four H.264 shapes, thresholds 0..127 and images whose band limits wrap
around uint8, bit-exact. The GPU parity tests then cover the same code through the kernels."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_thread_level_express_matches_oracle(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    exe = str(tmp_path / "xlcheck")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cxx", "express_lane_check.cc"),
                    "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    out = subprocess.run([exe, "240000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, cases, passes, alls = out.stdout.split()
    assert tag == "OK" and int(cases) == 240000
    assert int(passes) > 10000 and int(alls) > 10000      # feature blocks and wrapped bands were both exercised
