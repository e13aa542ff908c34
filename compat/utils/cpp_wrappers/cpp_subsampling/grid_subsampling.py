"""Drop-in for the reference's compiled CPython module `grid_subsampling` (built in place by compile_wrappers.sh).
helper_tool.py:14 imports it as `cpp_wrappers.cpp_subsampling.grid_subsampling`; with <repo>/compat/utils ahead of
the reference's utils/ on sys.path that import resolves here, unchanged."""
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from ssdr_al_b200.grid_subsampling import compute  # noqa: E402,F401
