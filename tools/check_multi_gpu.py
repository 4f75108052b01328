"""Multi-GPU parity check, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multi_gpu.py

Every rank fits the same (ragged) cell matrix sharded over the ranks through the real CUDA library (sums over
cells through the fixed 32-leaf tree, NCCL between ranks); rank 0 compares with the CPU oracle (log density 1e-5
relative, north_star) and with the Nystroem and Laplace variants; all ranks must hold identical bits, AND those
bits must equal what ONE GPU computes alone (rank 0 repeats everything unsharded): results do not depend on the
number of GPUs.  FULL / FULL_NYSTROEM fits (replicated factors) run too: they must not be summed over the ranks."""
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
    Xs = np.ascontiguousarray(X[:700])
    nns = be.nn_distances(Xs)
    full = mb.DensityEstimator(n_landmarks=0, nn_distances=nns, predictor_with_uncertainty=True)
    dens_full = full.fit_predict(Xs)
    dens_fnys = mb.DensityEstimator(n_landmarks=0, rank=0.9, nn_distances=nns).fit_predict(Xs)

    def sha(*arrays):
        return hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)).hexdigest()

    parts = {"log_density": dens, "predict": pred, "nystroem": dens_nys, "laplace_std": est.pre_transformation_std,
             "function_predict": fe_pred, "function_leverage": fe_lev, "function_obs_variance": fe_var,
             "full": dens_full, "full_nystroem": dens_fnys}
    digest = sha(*parts.values())
    one_gpu = None
    if world > 1 and rank == 0:
        # the same work on ONE GPU (nothing sharded), inside this job: bits must match the sharded results
        with be.replicated():
            e1 = mb.DensityEstimator(landmarks=lm, nn_distances=nn, check_rank=False, predictor_with_uncertainty=True)
            d1 = e1.fit_predict(X)
            n1 = mb.DensityEstimator(landmarks=lm, nn_distances=nn, rank=200, check_rank=False).fit_predict(X)
            f1 = mb.FunctionEstimator(landmarks=lm, ls=0.8, sigma=np.array([0.2, 0.5]), obs_variance=True).fit(X, yv)
            alone = {"log_density": d1, "predict": e1.predict(Y), "nystroem": n1, "laplace_std": e1.pre_transformation_std,
                     "function_predict": f1.predict(Y), "function_leverage": f1.leverage(),
                     "function_obs_variance": f1.get_obs_variance(Y)}
        one_gpu = {k: (sha(v) == sha(parts[k]), float(np.max(np.abs(np.asarray(v) - np.asarray(parts[k]))))) for k, v in alone.items()}
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
        ref_full = O.fit_density(Xs, nn_distances=O.compute_nn_distances(Xs), n_landmarks=0)
        ref_fnys = O.fit_density(Xs, nn_distances=O.compute_nn_distances(Xs), n_landmarks=0, rank=0.9)
        errs["full_log_density"] = rel(dens_full, ref_full.log_density_x)
        errs["full_nystroem_log_density"] = rel(dens_fnys, ref_fnys.log_density_x)
        print(f"world={world} identical_bits_on_all_ranks={same} nfev={est.opt_state.num_fun_eval} errors={errs}")
        print(f"world={world} same_bits_as_one_gpu (identical, max |diff|): {one_gpu}")
        inv = one_gpu is None or all(v[0] for v in one_gpu.values())
        ok = same and inv and errs["full_log_density"] < 1e-5 and errs["full_nystroem_log_density"] < 1e-5 and errs["nn_distances"] < 1e-12 and errs["log_density"] < 1e-5 and errs["predict"] < 1e-5 \
            and errs["nystroem_log_density"] < 1e-5 and errs["function_predict"] < 1e-6 and errs["function_leverage"] < 1e-6 \
            and errs["function_obs_variance"] < 1e-6
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    dist.barrier()


if __name__ == "__main__":
    main()
