"""CPU: the drop-in modules resolve at the reference's own import paths, with the UNMODIFIED helper_tool.py on top
(needs /root/reference; skipped on the GPU box where it does not exist)."""
import importlib
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/SSDR_AL_s3dis"


def test_compat_modules_import_standalone():
    sys.path.insert(0, os.path.join(ROOT, "compat", "utils"))
    try:
        nn = importlib.import_module("nearest_neighbors.lib.python.nearest_neighbors")
        gs = importlib.import_module("cpp_wrappers.cpp_subsampling.grid_subsampling")
        assert callable(nn.knn) and callable(nn.knn_batch) and callable(gs.compute)
        import ssdr_al_b200
        assert nn.knn_batch is ssdr_al_b200.nearest_neighbors.knn_batch
        assert gs.compute is ssdr_al_b200.grid_subsampling.compute
    finally:
        sys.path.remove(os.path.join(ROOT, "compat", "utils"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_unmodified_helper_tool_dispatches_into_the_drop_in(monkeypatch):
    # test-only stub for `from open3d import linux as open3d` (helper_tool.py:1); open3d is not installed here
    o3d = types.ModuleType("open3d")
    o3d.linux = types.ModuleType("open3d.linux")
    monkeypatch.setitem(sys.modules, "open3d", o3d)
    monkeypatch.setitem(sys.modules, "open3d.linux", o3d.linux)
    for m in [k for k in sys.modules if k.split(".")[0] in ("nearest_neighbors", "cpp_wrappers", "helper_tool")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat", "utils"))
    ht = importlib.import_module("helper_tool")
    import ssdr_al_b200
    assert ht.cpp_subsampling.compute is ssdr_al_b200.grid_subsampling.compute
    assert ht.nearest_neighbors.knn_batch is ssdr_al_b200.nearest_neighbors.knn_batch
    # DataProcessing.grid_sub_sampling forwards with the reference's keyword usage (helper_tool.py:226-235)
    seen = {}
    monkeypatch.setattr(ht.cpp_subsampling, "compute", lambda p, **kw: seen.update(kw) or "ok")
    assert ht.DataProcessing.grid_sub_sampling("pts", features="f", labels="l", grid_size=0.04) == "ok"
    assert seen == {"features": "f", "classes": "l", "sampleDl": 0.04, "verbose": 0}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_selection_patch_rebinds_reference_modules(monkeypatch):
    monkeypatch.syspath_prepend(REF)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    for m in ("fps_gcn_cpu", "kcenterGreedy", "ssdr_b200_patch"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    patch = importlib.import_module("ssdr_b200_patch")
    kc = importlib.import_module("kcenterGreedy")  # compat/ is first on the path -> the drop-in module
    import ssdr_al_b200
    assert kc.kCenterGreedy is ssdr_al_b200.selection.kCenterGreedy
    fps_mod = importlib.import_module("fps_gcn_cpu")  # the reference's own file
    assert fps_mod.__file__.startswith(REF)
    done = patch.install(modules=("fps_gcn_cpu",))
    assert "fps_gcn_cpu.farthest_features_sample" in done
    assert fps_mod.farthest_features_sample is ssdr_al_b200.selection.farthest_features_sample
    assert "fps_gcn_cpu.create_cd" in done and fps_mod.create_cd is ssdr_al_b200.chamfer.create_cd
    for m in ("fps_gcn_cpu", "kcenterGreedy", "ssdr_b200_patch"):
        sys.modules.pop(m, None)
