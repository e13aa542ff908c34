"""Drop-in for the reference's compiled Cython module `nearest_neighbors` (built by compile_op.sh into
utils/nearest_neighbors/lib/python/).  helper_tool.py:15 imports it as
`nearest_neighbors.lib.python.nearest_neighbors`; put <repo>/compat/utils ahead of the reference's utils/ on sys.path
(or overlay this tree onto a reference checkout) and that import resolves here, unchanged."""
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..", ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from ssdr_al_b200.nearest_neighbors import knn, knn_batch, knn_batch_distance_pick  # noqa: E402,F401
