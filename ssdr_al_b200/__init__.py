"""ssdr_al_b200 -- B200-native (sm_100a) drop-in for SSDR-AL's data-parallel point-cloud hot path.

Three parts, each behind the reference's own Python interface (file:line relative to SSDR_AL_s3dis/):
  * nearest_neighbors : knn, knn_batch                      (utils/nearest_neighbors/knn.pyx:33-109)
  * grid_subsampling  : compute(points, features=, classes=, sampleDl=, ...)
                                                            (utils/cpp_wrappers/cpp_subsampling/wrapper.cpp:58-286)
  * selection         : farthest_features_sample, kCenterGreedy
                                                            (fps_gcn_cpu.py:119-147, kcenterGreedy.py:48-128)
  * chamfer           : create_cd, the superpoint adjacency that feeds the FPS loop  (fps_gcn_cpu.py:12-38)
  * fps_gcn           : fps_adj_all, GCN_FPS_sampling (normalised adjacency, feature propagation)
                                                            (fps_gcn_cpu.py:40-117, :150-178)
Python only marshals numpy arrays into the C ABI of include/ssdr_b200.h (ctypes -> libssdr_b200.so -> CUDA).
There is no CPU fallback: without the built library or without a GPU every call raises.
"""
from . import _lib  # noqa: F401
from . import nearest_neighbors, grid_subsampling, selection, projection, chamfer, fps_gcn  # noqa: F401
from .selection import farthest_features_sample, kCenterGreedy  # noqa: F401

__version__ = "0.1.0"
