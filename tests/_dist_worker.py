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
    res = {
        **res_fe,
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
