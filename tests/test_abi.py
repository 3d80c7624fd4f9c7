"""CPU: the C-ABI library builds, loads, exports every symbol include/orbx.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from monocular_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "orbx.h")).read()
    return re.findall(r"ORBX_API\s+[\w\s\*]+?\b((?:orbx|hamx|fmx|trx|bowx|jpgx)_\w+)\s*\(", src)


def test_library_builds_and_exports_header():
    _lib.build()
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 30
    assert sorted(declared) == sorted(_lib.SYMBOLS), "include/orbx.h and _lib.SYMBOLS disagree"
    for s in declared:
        assert hasattr(L, s), "liborbx.so does not export %s" % s
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    names = {ln.split()[-1] for ln in exported.splitlines() if " T " in ln}
    assert names == set(declared), "exported symbols differ from the header: %s" % (names ^ set(declared))


def test_header_is_plain_c_and_matches_integration_md(tmp_path):
    """include/orbx.h compiled as C99 (the boundary is a C ABI), together with the call sequences INTEGRATION.md documents."""
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "cpp", "abi_c_check.c"), "-o", str(tmp_path / "abi_c_check.o")])


def test_pod_layouts():
    assert _lib.KEYPOINT_DTYPE.itemsize == 28 and _lib.DMATCH_DTYPE.itemsize == 16 and _lib.TOP2_DTYPE.itemsize == 16
    assert C.sizeof(_lib.Params) == 36
    p = _lib.Params()
    _lib.lib().orbx_default_params(C.byref(p))
    assert (p.nfeatures, round(p.scale_factor, 4), p.nlevels, p.edge_threshold, p.first_level, p.wta_k, p.score_type,
            p.patch_size, p.fast_threshold) == (500, 1.2, 8, 31, 0, 2, 0, 31, 20)   # reference defaults, FeatureExtractor.h:23-24


def test_sass_is_blackwell_native():
    """The matcher must be built for sm_100a, stage its train tiles with the TMA bulk-copy engine and use the 5th-generation tensor cores."""
    sass = subprocess.check_output(["cuobjdump", "-sass", _lib.LIB_PATH], text=True)
    assert "sm_100a" in sass or "SM100" in sass.upper()
    assert "UBLKCP" in sass and "POPC" in sass and "SYNCS" in sass
    assert "UTCIMMA" in sass and "LDTM" in sass          # tcgen05.mma (kind::i8) and tcgen05.ld: the tensor-core matcher


@pytest.mark.skipif(_lib.lib().orbx_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    from monocular_slam_b200 import ORB, BFMatcher, FundamentalFilter, OrbxError
    with pytest.raises(OrbxError) as e:
        ORB(nfeatures=500)
    assert e.value.status == _lib.E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(OrbxError):
        BFMatcher()
    with pytest.raises(OrbxError):
        FundamentalFilter()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "monocular_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "orb_oracle" not in txt, f
    for f in ("orbx.h", "orbx_shim.hpp"):
        p = os.path.join(ROOT, "include", f)
        if os.path.exists(p):
            assert "oracle" not in open(p).read()
