"""`jax.example_libraries.optimizers.adam` (the published algorithm), NumPy."""

import numpy as _np


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
    step = step_size if callable(step_size) else (lambda i: step_size)

    def init(x0):
        x0 = _np.asarray(x0, dtype=float)
        return x0, _np.zeros_like(x0), _np.zeros_like(x0)

    def update(i, g, state):
        x, m, v = state
        g = _np.asarray(g)
        m = (1 - b1) * g + b1 * m
        v = (1 - b2) * _np.square(g) + b2 * v
        mhat = m / (1 - b1 ** (i + 1))
        vhat = v / (1 - b2 ** (i + 1))
        x = x - step(i) * mhat / (_np.sqrt(vhat) + eps)
        return x, m, v

    def get_params(state):
        return state[0]

    return init, update, get_params
