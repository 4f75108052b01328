"""k-means landmarks with the distance work on the device.

The reference computes its landmarks with ``sklearn.cluster.k_means(x, k, n_init=1, random_state=seed)[0]``
(``mellon/parameters.py:243-291``).  This module runs THAT algorithm — same centring, same k-means++ seeding driven
by the same ``RandomState`` draws, same Lloyd iterations and stopping rules
(``sklearn/cluster/_kmeans.py``: ``KMeans.fit``, ``_kmeans_plusplus``, ``_kmeans_single_lloyd``) — with the two
O(N k D) pieces on the GPU:

* seeding: the distances of every cell to the ``2 + log k`` candidate rows of a step, the minimum with the closest
  distance so far and the candidates' potentials (``mb_sqdist_min``); the cumulative sum + ``searchsorted`` that turn
  the uniform draws into candidate rows stay NumPy's, on the downloaded closest-distance vector;
* Lloyd: the assignment step is the exact nearest-neighbour search ``mb_nn_distances`` of the cells against the
  centres (K1's distance tile with a running-minimum epilogue); the centroid averages are NumPy ``bincount`` sums.

The selected seed rows and the assignments are scikit-learn's own (they differ only if two squared distances agree
to rounding, which the tests would show); the centroids then agree to ~1e-13.  An empty cluster — scikit-learn
relocates it with a rule of its own — hands the whole computation to scikit-learn.
"""

from __future__ import annotations

import logging

import numpy as np

logger = logging.getLogger("mellon")

MAX_ITER = 300   # sklearn.cluster.k_means defaults
TOL = 1e-4


def _fallback(x, n_clusters, random_state, why):
    from sklearn.cluster import k_means

    logger.info("k-means on the host (scikit-learn): %s.", why)
    return k_means(x, n_clusters, n_init=1, random_state=random_state)[0]


def kmeans_plusplus(be, xd, xn_d, X, n_clusters, rs):
    """``_kmeans_plusplus`` with unit sample weights: returns the indices of the seed rows."""
    n = X.shape[0]
    n_local_trials = 2 + int(np.log(n_clusters))
    weight = np.ones(n)
    center_id = rs.choice(n, p=weight / weight.sum())
    indices = np.full(n_clusters, -1, dtype=int)
    indices[0] = center_id
    out, pot = be.sqdist_min(xd, xn_d, X[center_id][None, :], None)
    closest_d = out
    closest = be.download(out).ravel()
    current_pot = float(pot[0])
    for c in range(1, n_clusters):
        rand_vals = rs.uniform(size=n_local_trials) * current_pot
        candidate_ids = np.searchsorted(np.cumsum(closest), rand_vals)
        np.clip(candidate_ids, None, n - 1, out=candidate_ids)
        out, pot = be.sqdist_min(xd, xn_d, X[candidate_ids], closest_d)
        best = int(np.argmin(pot))
        current_pot = float(pot[best])
        closest_d = be.copy_rows(out, best, 1)
        closest = be.download(closest_d).ravel()
        indices[c] = candidate_ids[best]
    return indices


def k_means(x, n_clusters, random_state=None, backend=None):
    """Centroids of ``sklearn.cluster.k_means(x, n_clusters, n_init=1, random_state=random_state)``."""
    from sklearn.utils import check_random_state

    from .backend import get_backend

    be = backend or get_backend()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, d = x.shape
    n_local_trials = 2 + int(np.log(n_clusters))
    if n_local_trials > 16 or d > 512:
        return _fallback(x, n_clusters, random_state, "shape outside the device kernel's limits")
    rs = check_random_state(random_state)
    tol = float(np.mean(np.var(x, axis=0)) * TOL)
    x_mean = x.mean(axis=0)
    X = x - x_mean                                   # KMeans.fit centres the data for more accurate distances
    with be.replicated():                            # every rank computes the same landmarks on its own GPU
        xd = be.upload(X)
        xn_d = be.upload(np.einsum("ij,ij->i", X, X))   # sklearn.utils.extmath.row_norms(X, squared=True)
        indices = kmeans_plusplus(be, xd, xn_d, X, n_clusters, rs)
        centers = X[indices].copy()
        labels_old = np.full(n, -1, dtype=np.int64)
        for _ in range(MAX_ITER):
            labels = be.nearest_rows(xd, centers)
            counts = np.bincount(labels, minlength=n_clusters)
            if np.any(counts == 0):
                return _fallback(x, n_clusters, random_state, "a cluster went empty (scikit-learn relocates it)")
            centers_new = np.stack([np.bincount(labels, weights=X[:, j], minlength=n_clusters) for j in range(d)], axis=1)
            centers_new /= counts[:, None]
            center_shift = np.sqrt(np.sum((centers_new - centers) ** 2, axis=1))
            centers = centers_new
            if np.array_equal(labels, labels_old):
                break
            if float((center_shift ** 2).sum()) <= tol:
                break
            labels_old = labels
    return centers + x_mean
