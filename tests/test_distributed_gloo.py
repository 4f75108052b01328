"""The N > 1 path on CPU: world_size-2 (and 3) runs over gloo of the host-side sharding logic
(contiguous row blocks, all-reduce call sites for Gram / (loss, grad) / Hessian diagonal, row
gathers), compared with the single-process oracle."""

import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import mellon_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_row_block_partition():
    from mellon_b200.backend import row_block

    for n in (0, 1, 7, 100, 1501):
        for world in (1, 2, 3, 8):
            blocks = [row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]                       # contiguous, order preserving
            assert all(hi - lo <= per for lo, hi, per in blocks)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fit_matches_single_process_oracle(tmp_path, world):
    out = str(tmp_path / "res")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "_dist_worker.py"), out]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    res = [json.load(open(f"{out}.{r}")) for r in range(world)]

    rng = np.random.default_rng(0)
    X = rng.random((1501, 6))
    lm = X[:90].copy()
    nn = O.compute_nn_distances(X)
    ref = O.fit_density(X, landmarks=lm, nn_distances=nn)
    Y = rng.random((203, 6))
    ref_pred = O.predict_density(ref, X, Y)
    ref_nys = O.fit_density(X, landmarks=lm, nn_distances=nn, rank=30)
    yv = np.stack([np.sin(3 * X[:, 0]) + X[:, 1], np.cos(2 * X[:, 2])], axis=1)
    fe = O.function_fit(X, yv, landmarks=lm, mu=0.0, cov_func=O.Matern52(0.8), sigma=np.array([0.2, 0.5]),
                        obs_variance=True)
    for r in res:
        np.testing.assert_allclose(r["fe_pred"], O.conditional_mean(Y, lm, fe.weights, 0.0, fe.cov_func), rtol=1e-8)
        np.testing.assert_allclose(r["fe_lev"], O.function_leverage(fe, X), rtol=1e-8)
        np.testing.assert_allclose(r["fe_obsvar"], O.function_obs_variance(fe, Y), rtol=1e-7, atol=1e-10)
    for r in res:
        # every rank ends with the full, identical result
        np.testing.assert_allclose(r["nn"], nn, rtol=1e-12)
        np.testing.assert_allclose(r["dens"], ref.log_density_x, rtol=1e-5)
        np.testing.assert_allclose(r["pred"], ref_pred, rtol=1e-5)
        np.testing.assert_allclose(r["dens_nys"], ref_nys.log_density_x, rtol=1e-5)
        np.testing.assert_allclose(r["pred_nys"], O.predict_density(ref_nys, X, Y), rtol=1e-4)
        assert r["L_full_shape"] == [1501, 90]
        assert {"mb_gram", "mb_loss_grad", "mb_ridge_init"} <= set(r["calls"])
    assert res[0]["dens"] == res[1]["dens"]  # bit-identical across ranks
    std_ref = O.laplace_std_from_diag(O.hessian_diag(ref.L, nn, ref.d, ref.mu, ref.pre_transformation))
    np.testing.assert_allclose(res[0]["std"], std_ref, rtol=1e-3)
