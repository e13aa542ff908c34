"""CPU: the C-ABI library builds, loads, exports every symbol declared in include/ssdr_b200.h, and fails loudly
(never falls back) when no GPU is present.  No compute call succeeds here by design."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ssdr_b200.h")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib():
    from ssdr_al_b200.build import build
    build()
    from ssdr_al_b200 import _lib
    return _lib


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssdr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert len(declared) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (ssdr_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/ssdr_b200.h but not exported: %s" % missing


def test_ctypes_signatures_cover_the_header(lib):
    declared = set(_declared_symbols())
    assert declared == set(lib.SIGNATURES.keys())
    L = lib.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "ssdr_b200.h"\nint main(void){ssdr_knn_stats s; (void)s; return SSDR_OK;}\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_sass_is_sm100a(lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", lib.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_version_and_error_plumbing(lib):
    L = lib.lib()
    assert L.ssdr_version() >= 100
    rc = L.ssdr_device_count(None)
    assert rc == 1 and b"NULL" in L.ssdr_last_error()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    import ssdr_al_b200 as S
    p = np.random.default_rng(0).random((50, 3)).astype(np.float32)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.nearest_neighbors.knn(p, p, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.grid_subsampling.compute(p, sampleDl=0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.selection.fps(p, 5, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        S.kCenterGreedy(p).select_batch_([0], 3)


def test_wrapper_argument_checks_happen_before_the_device(lib):
    """wrapper.cpp:78-188 messages, raised by the shim without touching the GPU."""
    import ssdr_al_b200 as S
    G = S.grid_subsampling
    pts = np.zeros((10, 3), np.float32)
    with pytest.raises(RuntimeError, match=r"points.shape is not \(N, 3\)"):
        G.compute(np.zeros((10, 4)), sampleDl=0.1)
    with pytest.raises(RuntimeError, match=r"features.shape is not \(N, d\)"):
        G.compute(pts, features=np.zeros(10), sampleDl=0.1)
    with pytest.raises(RuntimeError, match=r"classes.shape is not \(N,\) or \(N, d\)"):
        G.compute(pts, classes=np.zeros((10, 1, 1), np.int32), sampleDl=0.1)
    with pytest.raises(RuntimeError, match="Error parsing method"):
        G.compute(pts, method="median")
    with pytest.raises(RuntimeError, match="converting input points"):
        G.compute([["a", "b", "c"]], sampleDl=0.1)
    with pytest.raises(RuntimeError, match="dim == 3"):
        S.nearest_neighbors.knn(np.zeros((5, 4), np.float32), np.zeros((5, 4), np.float32), 2)


def test_product_never_imports_the_oracle():
    """The shipped package must not reference oracle/ (a product path through the oracle would void parity)."""
    pkg = os.path.join(ROOT, "ssdr_al_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in ("import oracle", "from oracle", "liboracle", "oracle/", "oracle.", "_ref/"):
                    assert pat not in text, "%s mentions %r" % (os.path.join(dirpath, f), pat)


def test_batch_view_picks_the_upload_route_without_a_device():
    """knn_batch uploads slices of a larger float32 array as they are (item stride passed to the C ABI) and packs
    everything else on the host like knn.pyx:95-96 does."""
    from ssdr_al_b200.nearest_neighbors import _batch_view
    base = np.arange(4 * 100 * 3, dtype=np.float32).reshape(4, 100, 3)
    a, st = _batch_view(base)
    assert a is base and st == 300
    v = base[:, :25, :]
    a, st = _batch_view(v)
    assert a is v and st == 300 and not v.flags["C_CONTIGUOUS"]
    a, st = _batch_view(base[1:3, 10:20])            # offset slices keep the parent's item stride
    assert st == 300 and a.shape == (2, 10, 3)
    for odd in (base[:, ::2, :], base[:, :, ::-1], base.astype(np.float64), base.tolist(), base[::-1]):
        a, st = _batch_view(odd)
        assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float32 and st == a.shape[1] * 3
        assert np.array_equal(a, np.asarray(odd, dtype=np.float32))
