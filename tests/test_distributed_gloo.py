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
            cr = max(1, -(-n // 32))                      # whole chunks of the reduction tree
            assert all(lo % cr == 0 and (hi % cr == 0 or hi == n) for lo, hi, _ in blocks)


def run_world(tmp_path, world):
    out = str(tmp_path / f"res{world}")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "_dist_worker.py"), out]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    return [json.load(open(f"{out}.{r}")) for r in range(world)]


def test_results_do_not_depend_on_the_number_of_ranks(tmp_path):
    """Every sum over cells goes through the fixed chunk tree, and rank boundaries are chunk boundaries: on fixed
    sharded operands the reductions are bit-identical for 1, 2 and 3 ranks.  The whole fits agree to 1e-9 only,
    because the test double's row-wise products go through NumPy's BLAS, whose per-row rounding depends on the
    shape of the block it is handed; the CUDA library's row-wise kernels do not, and it is held to bit-identical
    fits on hardware (tools/check_multi_gpu.py, bench.py's `parity.vs_one_gpu_sha256` at every N)."""
    one = run_world(tmp_path, 1)[0]
    for world in (2, 3):
        for r in run_world(tmp_path, world):
            for key in ("gram_fix", "loss_fix", "grad_fix", "hess_fix", "gemv_fix", "z0_fix"):
                assert r[key] == one[key], f"{key} differs between 1 and {world} ranks (rank {r['rank']})"
            for key in ("dens", "pred", "std", "dens_nys", "pred_nys", "fe_pred", "fe_lev", "fe_obsvar", "dens_full",
                        "dens_fnys", "std_full"):
                np.testing.assert_allclose(r[key], one[key], rtol=1e-9, atol=1e-12, err_msg=key)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fit_matches_single_process_oracle(tmp_path, world):
    res = run_world(tmp_path, world)

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
    # replicated factors (FULL / FULL_NYSTROEM) are not summed over the ranks
    Xs = X[:260]
    nns = O.compute_nn_distances(Xs)
    ref_full = O.fit_density(Xs, nn_distances=nns, n_landmarks=0)
    ref_fnys = O.fit_density(Xs, nn_distances=nns, n_landmarks=0, rank=0.9)
    for r in res:
        assert not r["full_sharded"]
        np.testing.assert_allclose(r["dens_full"], ref_full.log_density_x, rtol=1e-5)
        np.testing.assert_allclose(r["pred_full"], O.predict_density(ref_full, Xs, Y), rtol=1e-5)
        np.testing.assert_allclose(r["dens_fnys"], ref_fnys.log_density_x, rtol=1e-5)
        assert r["rank_full"] == res[0]["rank_full"] and r["rank_sparse"] == res[0]["rank_sparse"]
    std_full_ref = O.laplace_std_from_diag(O.hessian_diag(ref_full.L, nns, ref_full.d, ref_full.mu,
                                                           ref_full.pre_transformation))
    np.testing.assert_allclose(res[0]["std_full"], std_full_ref, rtol=1e-3)
    std_ref = O.laplace_std_from_diag(O.hessian_diag(ref.L, nn, ref.d, ref.mu, ref.pre_transformation))
    np.testing.assert_allclose(res[0]["std"], std_ref, rtol=1e-3)
