"""Drop-in check: run the REFERENCE's own pytest files for this path against mellon_b200 (imported as `mellon`).

    python tools/run_reference_tests.py [--cuda] [pytest args ...]

The reference's tests import `mellon` and `jax`; here `mellon` resolves to this package and `jax` / `jaxopt` /
`pynndescent` to the NumPy stand-ins of oracle/refshim (test infrastructure).  Without --cuda the NumPy test double of
the C ABI (tests/fake_lib.py) stands in for the GPU, so this runs in the CPU container; reads /root/reference, so it
is a local tool, not part of the test suite that travels to the GPU box."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MELLON_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.dont_write_bytecode = True

import mellon_b200  # noqa: E402

use_cuda = "--cuda" in sys.argv
args = [a for a in sys.argv[1:] if a != "--cuda"]
if not use_cuda:
    from fake_lib import FakeBackend  # noqa: E402

    mellon_b200.set_backend(FakeBackend(0, 1))

sys.modules["mellon"] = mellon_b200
for name in ("cov", "base_cov", "util", "parameters", "inference", "conditional", "base_predictor", "base_model",
             "density_estimator", "function_estimator", "time_sensitive_density_estimator", "decomposition", "validation",
             "parameter_validation", "model", "compute_ls_time"):
    try:
        sys.modules[f"mellon.{name}"] = importlib.import_module(f"mellon_b200.{name}")
    except ImportError:
        pass

import pytest  # noqa: E402

DEFAULT = ["test_density_estimator.py", "test_cov.py", "test_base_cov.py", "test_parameters.py", "test_inference.py",
           "test_laplace.py", "test_time_sensitive_density_estimator.py", "test_util.py", "test_validation.py",
           # FunctionEstimator; test_sigma_to_y_cov_factor.py imports a private helper that materialises
           # eye(n) * sigma, which this package never forms
           "test_function_estimator.py", "test_reference_results.py", "test_leverage.py", "test_pergene_sigma.py",
           "test_perobservation_sigma.py"]
files = [a for a in args if a.endswith(".py")] or DEFAULT
rest = [a for a in args if not a.endswith(".py")]
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", *rest, *[os.path.join(REF, "tests", f) for f in files]]))
