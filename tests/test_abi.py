"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/movfe.h declares,
the record layouts match, and — with no GPU — it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from movfe import lib, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "movfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(movfe_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmovfe.so does not export %s" % n
    assert sorted(lib.EXPORTS) == names


def test_record_layouts_match_c():
    import subprocess
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "movfe.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(movfe_mv_record), offsetof(movfe_mv_record, ref),
  offsetof(movfe_mv_record, dst_x), sizeof(movfe_hop), sizeof(movfe_rect), sizeof(movfe_track), sizeof(movfe_map_point),
  sizeof(movfe_projection), sizeof(movfe_camera), sizeof(movfe_pose));
 printf("%zu %zu %zu\n", sizeof(movfe_pose_params), sizeof(movfe_config), offsetof(movfe_config, coverage_threshold));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    v = [int(x) for x in out]
    assert v[:10] == [T.MV_RECORD.itemsize, T.MV_RECORD.fields["ref"][1], T.MV_RECORD.fields["dst_x"][1], T.HOP.itemsize,
                      T.RECT.itemsize, T.TRACK.itemsize, T.MAP_POINT.itemsize, T.PROJECTION.itemsize,
                      T.CAMERA.itemsize, T.POSE.itemsize]
    assert v[10] == T.POSE_PARAMS.itemsize and v[11] == C.sizeof(lib.Config) and v[12] == lib.Config.coverage_threshold.offset


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.MovfeError, match="no CUDA device"):
        lib.Context(1, 64, 48)


def test_create_rejects_bad_config():
    L = lib.load()
    cfg = lib.Config(0, 0, 64, 48, 100, 3, 4, 128, 0, 25, 0.2, 0, 0)
    h = C.c_void_p()
    assert L.movfe_create(C.byref(cfg), C.byref(h)) == -1 and not h.value
    assert b"out of range" in L.movfe_last_error(None)


def test_product_never_touches_the_oracle():
    """The product tree (mov-slam_b200/, include/) must not reference oracle/ in any way."""
    for base in ("mov-slam_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dp or os.sep + "lib" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc", ".log")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "liboracle" not in txt and "pyoracle" not in txt and "orc_" not in txt, os.path.join(dp, fn)


def test_pack_records_is_the_seven_fields_the_path_reads():
    """movfe_pack_records (host code, no GPU): the 16-byte record carries source sign, block size, both centres and ref."""
    rng = np.random.default_rng(5)
    n = 1000
    r = np.zeros(n, T.MV_RECORD)
    r["source"] = rng.integers(-3, 4, n)
    r["w"] = rng.choice([4, 8, 16], n)
    r["h"] = rng.choice([4, 8, 16], n)
    for k in ("src_x", "src_y", "dst_x", "dst_y"):
        r[k] = rng.integers(-300, 2000, n)
    r["ref"] = rng.integers(-1, 12, n)
    r["flags"] = rng.integers(0, 1 << 40, n)
    r["motion_x"] = rng.integers(-500, 500, n)
    p = lib.pack_records(r)
    assert p.dtype.itemsize == 16
    for k in ("src_x", "src_y", "dst_x", "dst_y", "w", "h", "ref"):
        assert np.array_equal(p[k], r[k]), k
    assert np.array_equal(p["source_sign"], np.sign(r["source"]))
    assert not p["reserved"].any()
