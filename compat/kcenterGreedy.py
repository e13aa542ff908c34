"""Drop-in for SSDR_AL_s3dis/kcenterGreedy.py (imported by gcn.py:10 and fps_gcn_cpu.py:9 with `from kcenterGreedy
import *`): same public names, the greedy loop runs on the GPU."""
import os
import sys

_REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import abc  # noqa: E402,F401
import numpy as np  # noqa: E402,F401

from ssdr_al_b200.selection import kCenterGreedy  # noqa: E402,F401


class SamplingMethod(object):
    """Name kept for `from kcenterGreedy import *` users (kcenterGreedy.py:14-44)."""
    __metaclass__ = abc.ABCMeta
