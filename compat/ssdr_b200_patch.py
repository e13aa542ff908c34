"""One-line activation of the GPU selection loop inside an unmodified SSDR-AL checkout:

    import ssdr_b200_patch; ssdr_b200_patch.install()

farthest_features_sample is defined inside fps_gcn_cpu.py / fps_gcn_cuda.py themselves (fps_gcn_cpu.py:119,
fps_gcn_cuda.py:123) and looked up as a module global by GCN_FPS_sampling (fps_gcn_cpu.py:170), so rebinding the
module attribute is enough; kCenterGreedy is rebound in gcn.py's namespace (gcn.py:10 star-import); create_cd
(fps_gcn_cpu.py:25, called as a module global by fps_adj_all at :93) and sampler2.farthest_superpoint_sample
(sampler2.py:49, called at :577) are rebound the same way, and so are fps_adj_all / GCN_FPS_sampling
(fps_gcn_cpu.py:40, :150; sampler2.py:10 imports the latter by name, so sampler2's binding is replaced as well)."""
import importlib
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


def install(modules=("fps_gcn_cpu", "fps_gcn_cuda", "gcn", "kcenterGreedy", "sampler2")):
    from ssdr_al_b200.selection import farthest_features_sample, kCenterGreedy
    from ssdr_al_b200.chamfer import create_cd, farthest_superpoint_sample
    from ssdr_al_b200.fps_gcn import GCN_FPS_sampling, fps_adj_all
    patched = []
    for name in modules:
        try:
            mod = sys.modules.get(name) or importlib.import_module(name)
        except Exception:
            continue
        if hasattr(mod, "farthest_features_sample"):
            mod.farthest_features_sample = farthest_features_sample
            patched.append(name + ".farthest_features_sample")
        if hasattr(mod, "kCenterGreedy"):
            mod.kCenterGreedy = kCenterGreedy
            patched.append(name + ".kCenterGreedy")
        if name == "fps_gcn_cpu" and hasattr(mod, "create_cd"):  # (fps_gcn_cuda's create_cd is a torch/chamfer3D op)
            mod.create_cd = create_cd
            patched.append(name + ".create_cd")
        if name == "fps_gcn_cpu" and hasattr(mod, "fps_adj_all"):  # fps_gcn_cpu.py:40 / :150 (numpy variant only)
            mod.fps_adj_all = fps_adj_all
            mod.GCN_FPS_sampling = GCN_FPS_sampling
            patched += [name + ".fps_adj_all", name + ".GCN_FPS_sampling"]
        if name == "sampler2" and getattr(mod, "GCN_FPS_sampling", None) is not None and \
                getattr(mod.GCN_FPS_sampling, "__module__", "") == "fps_gcn_cpu":  # sampler2.py:10 `from fps_gcn_cpu import`
            mod.GCN_FPS_sampling = GCN_FPS_sampling
            patched.append(name + ".GCN_FPS_sampling")
        if name == "sampler2" and hasattr(mod, "farthest_superpoint_sample"):  # sampler2.py:49, called at :577
            mod.farthest_superpoint_sample = farthest_superpoint_sample
            patched.append(name + ".farthest_superpoint_sample")
    return patched
