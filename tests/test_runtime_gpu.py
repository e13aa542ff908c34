"""GPU: runtime behaviour of the C-ABI library -- re-entrancy per calling thread, fork detection, workspace reuse
across differently sized calls, error reporting."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import ssdr_al_b200 as S
    return S


def test_two_threads_call_concurrently(S, oracle):
    rng = np.random.default_rng(0)
    clouds = [rng.random((20000 + 1000 * i, 3), dtype=np.float32) for i in range(2)]
    feats = [rng.standard_normal((30000, 32)).astype(np.float32) for _ in range(2)]
    want_knn = [oracle.knn(c, c, 16, threads=4) for c in clouds]
    want_fps = [oracle.fps(f, 50, 3) for f in feats]
    errors = []

    def work(i):
        try:
            for _ in range(5):  # ctypes releases the GIL: the two threads really overlap inside the library
                assert np.array_equal(S.nearest_neighbors.knn(clouds[i], clouds[i], 16), want_knn[i])
                assert np.array_equal(S.selection.fps(feats[i], 50, 3), want_fps[i])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors


def test_workspaces_survive_growing_and_shrinking_calls(S, oracle):
    rng = np.random.default_rng(1)
    for n in (500, 80_000, 3_000, 150_000, 17):
        p = rng.random((n, 3), dtype=np.float32)
        k = min(16, n)
        assert np.array_equal(S.nearest_neighbors.knn(p, p, k), oracle.knn(p, p, k, threads=4))
        got = S.grid_subsampling.compute(p, sampleDl=0.05)
        assert got.tobytes() == oracle.grid_subsample(p, None, None, 0.05, order="key")[0].tobytes()


def test_forked_child_gets_a_clear_error_not_a_hang(S):
    p = np.random.default_rng(2).random((1000, 3), dtype=np.float32)
    S.nearest_neighbors.knn(p, p, 4)  # initialise CUDA in the parent, like TF does before the DataLoader forks
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:  # child
        os.close(r)
        msg = b"no error"
        try:
            S.nearest_neighbors.knn(p, p, 4)
        except RuntimeError as e:
            msg = str(e).encode()
        os.write(w, msg[:400])
        os._exit(0)
    os.close(w)
    out = os.read(r, 1000).decode()
    os.waitpid(pid, 0)
    assert "fork" in out and "spawn" in out, out


def test_errors_are_runtime_errors_with_messages(S):
    p = np.zeros((10, 3), np.float32)
    with pytest.raises(RuntimeError, match="K=100 > 64"):
        S.nearest_neighbors.knn(p, p, 100)
    with pytest.raises(RuntimeError, match="sampleDl must be positive"):
        S.grid_subsampling.compute(p, sampleDl=-1.0)
    with pytest.raises(RuntimeError, match="first index"):
        S.selection.fps(np.zeros((5, 4), np.float32), 3, 7)
    with pytest.raises(RuntimeError, match="selected"):
        S.selection.kcenter(np.zeros((5, 4), np.float32), [9], 2)
