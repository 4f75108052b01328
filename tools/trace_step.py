"""Host-clock stage breakdown of one fit_predict step (device synchronised between stages).

    python tools/trace_step.py [N] [M] [D]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mellon_b200 as mb
from mellon_b200 import cov as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 50
mb.setup_logging().setLevel("WARNING")
be = mb.get_backend()
x = np.random.default_rng(0).random((N, D))
idx = np.sort(np.random.default_rng(1).choice(N, size=M, replace=False))
lm = np.ascontiguousarray(x[idx])
nn = be.nn_distances(x)
xd = be.upload(x, sharded=True)

for rep in range(2):
    est = mb.DensityEstimator(cov_func_curry=C.Matern52, landmarks=lm, nn_distances=nn, check_rank=False)
    be.sync()
    t = [time.perf_counter()]
    names = []

    def mark(name):
        be.sync()
        t.append(time.perf_counter())
        names.append(name)

    est.set_x(xd)
    for a in ("n_landmarks", "rank", "gp_type"):
        est._prepare_attribute(a)
    est.validate_parameter()
    for a in ("nn_distances", "d", "mu", "ls", "cov_func", "landmarks"):
        est._prepare_attribute(a)
    mark("params(mu,ls,..)")
    est._prepare_attribute("Lp"); mark("Lp = chol(K_MM)")
    est._prepare_attribute("L"); mark("L = K_NM Lp^-T")
    est._prepare_attribute("initial_value"); mark("ridge init")
    est._prepare_attribute("transform"); est._prepare_attribute("loss_func"); mark("loss_func setup")
    est.run_inference(); mark("L-BFGS-B (%d evals)" % est.opt_state.num_fun_eval)
    est.process_inference(build_predict=False); mark("log_density = Lz+mu")
    del est
    mark("free")
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"--- rep {rep}: total {1e3 * (t[-1] - t[0]):.1f} ms (world {os.environ.get('WORLD_SIZE', '1')})")
        for nm, a, b in zip(names, t[:-1], t[1:]):
            print(f"  {nm:28s} {1e3 * (b - a):9.1f} ms")
