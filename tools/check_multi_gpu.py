"""Multi-GPU parity check, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py

Every rank fits the same (ragged) cell matrix sharded over the ranks through the real CUDA library with
NCCL all-reduces; rank 0 compares with the CPU oracle (log density 1e-5 relative, north_star) and with the
Nystroem and Laplace variants; all ranks must hold identical bits."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mellon_b200 as mb
from mellon_b200 import distributed as dist
from oracle import mellon_oracle as O


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    mb.setup_logging().setLevel("WARNING")
    be = mb.get_backend()
    rng = np.random.default_rng(0)
    n, d, m = 20011, 20, 600  # ragged: 20011 rows do not divide by the world size
    X = rng.random((n, d))
    lm = X[np.sort(rng.choice(n, m, replace=False))].copy()
    nn = be.nn_distances(X)
    est = mb.DensityEstimator(landmarks=lm, nn_distances=nn, check_rank=False, predictor_with_uncertainty=True)
    dens = est.fit_predict(X)
    Y = rng.random((5003, d))
    pred = est.predict(Y)
    nys = mb.DensityEstimator(landmarks=lm, nn_distances=nn, rank=200, check_rank=False)
    dens_nys = nys.fit_predict(X)
    # regression on the same sharded cells (FunctionEstimator, sparse conditional, per-feature noise, observation variance)
    yv = np.stack([np.sin(3 * X[:, 0]) + X[:, 1], np.cos(2 * X[:, 2])], axis=1)
    fe = mb.FunctionEstimator(landmarks=lm, ls=0.8, sigma=np.array([0.2, 0.5]), obs_variance=True).fit(X, yv)
    fe_pred, fe_lev, fe_var = fe.predict(Y), fe.leverage(), fe.get_obs_variance(Y)
    digest = hashlib.sha256(dens.tobytes() + pred.tobytes() + dens_nys.tobytes() + fe_pred.tobytes() + fe_lev.tobytes()
                            + fe_var.tobytes()).hexdigest()
    import torch
    import torch.distributed as td

    same = True
    if world > 1:
        gathered = [None] * world
        td.all_gather_object(gathered, digest)
        same = len(set(gathered)) == 1
    if rank == 0:
        nn_ref = O.compute_nn_distances(X)
        ref = O.fit_density(X, landmarks=lm, nn_distances=nn_ref)
        ref_pred = O.predict_density(ref, X, Y)
        ref_nys = O.fit_density(X, landmarks=lm, nn_distances=nn_ref, rank=200)
        fo = O.function_fit(X, yv, landmarks=lm, mu=0.0, cov_func=O.Matern52(0.8), sigma=np.array([0.2, 0.5]), obs_variance=True)
        amax = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
        errs = {
            "function_predict": amax(fe_pred, O.conditional_mean(Y, lm, fo.weights, 0.0, fo.cov_func)),
            "function_leverage": amax(fe_lev, O.function_leverage(fo, X)),
            "function_obs_variance": amax(fe_var, O.function_obs_variance(fo, Y)),
            "nn_distances": rel(nn, nn_ref),
            "log_density": rel(dens, ref.log_density_x),
            "predict": rel(pred, ref_pred),
            "nystroem_log_density": rel(dens_nys, ref_nys.log_density_x),
            "laplace_std": rel(est.pre_transformation_std,
                               O.laplace_std_from_diag(O.hessian_diag(ref.L, nn_ref, ref.d, ref.mu, est.pre_transformation))),
        }
        print(f"world={world} identical_bits_on_all_ranks={same} nfev={est.opt_state.num_fun_eval} errors={errs}")
        ok = same and errs["nn_distances"] < 1e-12 and errs["log_density"] < 1e-5 and errs["predict"] < 1e-5 \
            and errs["nystroem_log_density"] < 1e-5 and errs["function_predict"] < 1e-6 and errs["function_leverage"] < 1e-6 \
            and errs["function_obs_variance"] < 1e-6
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    dist.barrier()


if __name__ == "__main__":
    main()
