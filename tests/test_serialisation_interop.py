"""SURVEY.md §8f.4 — predictor serialisation is interchangeable with the reference's.

tests/golden/reference_predictors.json holds predictors fitted and written by the UNMODIFIED reference
(`Predictor.to_json`, base_predictor.py:541-734; minted by oracle/make_golden.py --predictors-only) together with the
reference's own predictions.  Loading the JSON text with this package must reproduce them, and what this package
writes must have the reference's layout (same keys, same array tags), so either side loads the other's files.
The reverse direction (the reference loading a file written here) needs /root/reference and lives in
tools/check_json_interop.py; its output is profiles/json_interop_r01.txt."""

import json
import os

import numpy as np
import pytest

import mellon_b200 as mb

PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_predictors.json")
with open(PATH) as _f:
    CASES = json.load(_f)


def _args(case):
    return [np.asarray(a) for a in case["pred_args"]]


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_written_predictor_loads_and_predicts(be, name):
    case = CASES[name]
    pred = mb.Predictor.from_json_str(case["json"])
    assert type(pred).__name__ == case["classname"]
    Y = np.asarray(case["Y"])
    np.testing.assert_allclose(pred(Y, *_args(case)), case["mean"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(pred(Y, *_args(case), normalize=True), case["mean_normalized"], rtol=1e-9, atol=1e-9)
    if "covariance" in case:
        np.testing.assert_allclose(pred.covariance(Y), case["covariance"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(pred.mean_covariance(Y), case["mean_covariance"], rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(pred.uncertainty(Y), case["uncertainty"], rtol=1e-5, atol=1e-9)
    if "mean_multi_time" in case:
        np.testing.assert_allclose(pred(Y, multi_time=[0.0, 1.5]), case["mean_multi_time"], rtol=1e-9, atol=1e-9)


def _strip_volatile(node):
    if isinstance(node, dict):
        return {k: _strip_volatile(v) for k, v in node.items()
                if k not in ("serialization_date", "module_version", "python_version")}
    return node


def _layout(node):
    """The shape of a serialised document: keys and array tags, not the numbers."""
    if isinstance(node, dict):
        if set(node) >= {"type", "data"} and isinstance(node["data"], (list, str, int, float)):
            return {"type": node["type"], **{k: "…" for k in node if k != "type"}}
        return {k: _layout(v) for k, v in sorted(node.items())}
    if isinstance(node, list):
        return "list"
    return type(node).__name__


@pytest.mark.parametrize("name", sorted(CASES))
def test_rewritten_document_has_the_reference_layout(be, name):
    case = CASES[name]
    ref_doc = json.loads(case["json"])
    ours = json.loads(mb.Predictor.from_json_str(case["json"]).to_json())
    lo, lr = _layout(ours), _layout(ref_doc)
    assert lo == lr
    # class and module names are the reference's, so stock Mellon resolves them; version and date differ
    assert _strip_volatile(ours["metadata"]) == _strip_volatile(ref_doc["metadata"])
    assert _strip_volatile(ours["cov_func"]) == _strip_volatile(ref_doc["cov_func"])
    assert ours["metadata"]["module_version"].startswith("1.7.1")
    # numbers survive the round trip bit for bit
    def arrays(node, path=""):
        if isinstance(node, dict):
            if node.get("type") == "jax.numpy":
                yield path, np.asarray(node["data"], dtype=float)
            else:
                for k, v in node.items():
                    yield from arrays(v, path + "/" + k)
    a, b = dict(arrays(ours["data"])), dict(arrays(ref_doc["data"]))
    assert a.keys() == b.keys() and len(a) >= 2
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def test_covariance_json_matches_reference_layout():
    for case in CASES.values():
        ref_doc = json.loads(case["cov_func_json"])
        ours = json.loads(mb.cov.Covariance.from_json(case["cov_func_json"]).to_json())
        assert _strip_volatile(ours) == _strip_volatile(ref_doc)
