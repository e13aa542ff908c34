"""Diversity selection on the GPU behind the reference's own interfaces.

  farthest_features_sample(feature_list, sample_number)  -- fps_gcn_cpu.py:119-147 (== fps_gcn_cuda.py:123-151)
  kCenterGreedy(X).select_batch_(already_selected, N)    -- kcenterGreedy.py:48-128

Both keep the reference's argument meaning, return types and random-number use (the first FPS index comes from
np.random.randint exactly like fps_gcn_cpu.py:133, so seeding numpy reproduces the reference's stream).
"""
import numpy as np

from . import _lib


def _as_features(x):
    a = np.array(x) if not isinstance(x, np.ndarray) else x
    if a.dtype == np.float64:
        return np.ascontiguousarray(a), "f64"
    if a.dtype == np.float32:
        return np.ascontiguousarray(a), "f32"
    # ints / float16 / object lists: numpy would compute in the promoted dtype; float64 covers ints exactly
    return np.ascontiguousarray(a, dtype=np.float64), "f64"


def fps(features, sample_number, first):
    """FPS with the first index given: returns (sample_number,) int32, picks[0] == first."""
    F, tag = _as_features(features)
    if F.ndim != 2:
        F = F.reshape(len(F), -1)
    out = np.zeros([sample_number], dtype=np.int32)
    if sample_number == 0:
        return out
    fn = _lib.lib().ssdr_fps_f32 if tag == "f32" else _lib.lib().ssdr_fps_f64
    _lib.check(fn(_lib.ptr(F), F.shape[0], F.shape[1], int(first), int(sample_number), _lib.ptr(out)))
    return out


def farthest_features_sample(feature_list, sample_number):
    """Drop-in for fps_gcn_cpu.farthest_features_sample (fps_gcn_cpu.py:119-147)."""
    list_num = len(feature_list)
    first = np.random.randint(0, list_num)  # fps_gcn_cpu.py:133 -- same RNG call, same stream
    return fps(feature_list, sample_number, first)


def kcenter(X, already_selected, n_pick):
    """k-center greedy picks as an int64 array (no Python-object boxing)."""
    F, tag = _as_features(X)
    sel = np.ascontiguousarray(np.asarray(already_selected, dtype=np.int64).reshape(-1))
    out = np.zeros(int(n_pick), dtype=np.int64)
    fn = _lib.lib().ssdr_kcenter_f32 if tag == "f32" else _lib.lib().ssdr_kcenter_f64
    _lib.check(fn(_lib.ptr(F), F.shape[0], F.shape[1], _lib.ptr(sel) if len(sel) else None, len(sel), int(n_pick),
                  _lib.ptr(out)))
    return out


class kCenterGreedy(object):
    """Drop-in for kcenterGreedy.kCenterGreedy (kcenterGreedy.py:48-128): same constructor and select_batch_.

    The per-pick `pairwise_distances` + `np.minimum` + `np.argmax` loop runs inside one persistent CUDA kernel.
    """

    def __init__(self, X, metric='euclidean'):
        if metric != 'euclidean':
            raise RuntimeError("ssdr_al_b200.kCenterGreedy supports metric='euclidean' only (the reference default)")
        self.X = X
        shape = X.shape
        self.flat_X = X if len(shape) <= 2 else np.reshape(X, (shape[0], int(np.prod(shape[1:]))))
        self.name = 'kcenter'
        self.features = self.flat_X
        self.metric = metric
        self.min_distances = None
        self.max_distances = None
        self.n_obs = self.X.shape[0]
        self.already_selected = []

    def select_batch_(self, already_selected, N, **kwargs):
        already = np.asarray(already_selected, dtype=np.int64).reshape(-1)
        if len(self.already_selected) and not np.isin(np.asarray(self.already_selected, dtype=np.int64), already).all():
            # Every call rebuilds min_distances from its own `already_selected` (kcenterGreedy.py:104, reset_dist=True),
            # so a later call is a fresh computation -- unless a centre of the EARLIER call is missing from the new
            # list: the reference then refuses to apply it as a new centre (only_new filter, :74-75) and repeats the
            # same pick.  That corner is not reproduced.
            raise RuntimeError("ssdr_al_b200.kCenterGreedy: a later select_batch_ must keep the earlier call's "
                               "already_selected entries (gcn.py builds a fresh object per call)")
        picks = kcenter(self.features, already, N)
        clash = np.intersect1d(picks, already)
        assert clash.size == 0  # kcenterGreedy.py:118
        self.already_selected = already_selected
        return [np.int64(i) for i in picks]
