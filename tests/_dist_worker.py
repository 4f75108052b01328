"""Worker of tests/test_distributed_gloo.py: one rank of a world_size-N run of the HOST side of the
multi-GPU path (row sharding, all-reduce call sites, gathers) over gloo, with the NumPy test double
of the C ABI standing in for the CUDA library.  Launched by torch.distributed.run."""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch.distributed as dist  # noqa: E402

import mellon_b200 as mb  # noqa: E402
from fake_lib import FakeBackend  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo")
    mb.setup_logging().setLevel("WARNING")
    be = FakeBackend(rank, world)
    mb.set_backend(be)
    rng = np.random.default_rng(0)
    n, d, m = 1501, 6, 90  # ragged: 1501 rows do not divide by the world size
    X = rng.random((n, d))
    lm = X[:m].copy()
    nn = be.nn_distances(X)
    est = mb.DensityEstimator(landmarks=lm, nn_distances=nn, predictor_with_uncertainty=True)
    dens = est.fit_predict(X)
    Ld = est.L
    assert Ld.sharded and Ld.local_shape[0] == mb.backend.row_block(n, rank, world)[1] - mb.backend.row_block(n, rank, world)[0]
    Y = rng.random((203, d))
    pred = est.predict(Y)
    nys = mb.DensityEstimator(landmarks=lm, nn_distances=nn, rank=30)
    dens_nys = nys.fit_predict(X)
    pred_nys = nys.predict(Y)
    # regression on the same sharded cells (FunctionEstimator, sparse conditional): two outputs, per-feature noise
    yv = np.stack([np.sin(3 * X[:, 0]) + X[:, 1], np.cos(2 * X[:, 2])], axis=1)
    fe = mb.FunctionEstimator(landmarks=lm, ls=0.8, sigma=np.array([0.2, 0.5]), obs_variance=True).fit(X, yv)
    res_fe = {"fe_pred": fe.predict(Y).tolist(), "fe_lev": fe.leverage().tolist(),
              "fe_obsvar": fe.get_obs_variance(Y).tolist()}
    # replicated (non-sharded) factors: FULL and FULL_NYSTROEM keep an N x N / N x p factor on every rank, which
    # must NOT be summed over the ranks; util.test_rank on a replicated and on the sharded factor
    Xs = np.ascontiguousarray(X[:260])
    nns = be.nn_distances(Xs)
    full = mb.DensityEstimator(n_landmarks=0, nn_distances=nns, predictor_with_uncertainty=True)
    dens_full = full.fit_predict(Xs)
    fnys = mb.DensityEstimator(n_landmarks=0, rank=0.9, nn_distances=nns)
    dens_fnys = fnys.fit_predict(Xs)
    res_rep = {"dens_full": dens_full.tolist(), "pred_full": full.predict(Y).tolist(),
               "std_full": full.pre_transformation_std.tolist(), "dens_fnys": dens_fnys.tolist(),
               "rank_full": int(mb.util.test_rank(full.L)), "rank_sparse": int(mb.util.test_rank(est.L)),
               "full_sharded": bool(full.L.sharded) if hasattr(full.L, "sharded") else False}
    # the reductions themselves on FIXED sharded operands: what must be bit-identical for any number of ranks
    Lfix = rng.standard_normal((n, 40))
    Ldev = be.upload(Lfix, sharded=True)
    st = be.objective(Ldev, np.linspace(-1.0, 1.0, n), 0.25, -3.0, 40.0)
    loss_fix, grad_fix = be.loss_grad(st, np.linspace(-0.1, 0.1, 40))
    res_fix = {"gram_fix": be.gram(Ldev).numpy().tolist(), "loss_fix": loss_fix, "grad_fix": grad_fix.tolist(),
               "hess_fix": be.hess_diag(st, np.linspace(-0.1, 0.1, 40)).tolist(),
               "gemv_fix": be.gemv_t(Ldev, np.cos(np.arange(n))).tolist(),
               "z0_fix": be.ridge_init(Ldev, np.sin(np.arange(n))).tolist()}
    res = {
        **res_fe, **res_rep, **res_fix,
        "rank": rank, "dens": dens.tolist(), "pred": pred.tolist(), "nn": nn.tolist(),
        "std": est.pre_transformation_std.tolist(), "L_full_shape": list(np.asarray(Ld).shape),
        "dens_nys": dens_nys.tolist(), "pred_nys": pred_nys.tolist(),
        "calls": sorted(set(be.lib.calls)),
    }
    with open(f"{out_path}.{rank}", "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
