"""`pynndescent.NNDescent` stand-in: EXACT k-nearest-neighbour search (the real one is approximate)."""

import numpy as _np
from sklearn.neighbors import NearestNeighbors as _NN


class NNDescent:
    def __init__(self, data, n_neighbors=30, random_state=None, **kwargs):
        data = _np.asarray(data, dtype=float)
        dist, idx = _NN(n_neighbors=n_neighbors).fit(data).kneighbors(data)
        self.neighbor_graph = (idx, dist)
