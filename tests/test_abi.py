"""The C-ABI shared library loads without a GPU and exports exactly what include/mellon_b200.h
declares; the ctypes table mirrors the header; device calls fail loudly without a device."""

import ctypes
import os
import re

import pytest

from mellon_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mellon_b200.h")
DECL = re.compile(r"^(?:const\s+char\s*\*|int64_t|int)\s+(mb_[a-z0-9_]+)\s*\(", re.M)


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(DECL.findall(text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert len(syms) >= 50
    for must in ("mb_cov_build", "mb_cov_chol", "mb_lowrank_standard", "mb_gram", "mb_ridge_init", "mb_loss_grad",
                 "mb_hess_diag", "mb_transform", "mb_predict_mean", "mb_comm_init", "mb_nn_distances"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(nat.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(nat.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_mirrors_header():
    assert sorted(nat.SIGNATURES) == header_symbols()
    lib = nat.load_library()
    assert lib.mb_version() >= 100


def test_header_cites_reference_call_sites():
    text = open(HEADER).read()
    for cite in ("util.py:351-366", "decomposition.py:79-123", "decomposition.py:174-210", "parameters.py:877-896",
                 "inference.py:35-48", "conditional.py"):
        assert cite in text


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product backend refuses to start (no silent CPU path)."""
    if nat.device_count() > 0:
        pytest.skip("a GPU is visible")
    from mellon_b200.backend import CudaBackend

    with pytest.raises(nat.DeviceError, match="no CPU fallback"):
        CudaBackend()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(nat.NativeLibraryError):
        nat.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mellon_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in re.sub(r"#.*", "", src).replace("an oracle", ""), name


def test_library_is_built_from_the_sources_in_the_tree():
    """The prebuilt library reports the hash of the sources it was compiled from; it must be this tree's (a stale
    .so travelling with the snapshot would otherwise pass for the current code), and csrc/Makefile must list the same
    translation units in the same order as _native.SOURCES (both hash them)."""
    import re

    from mellon_b200 import _native as nat

    lib = nat.load_library()
    assert lib.mb_source_hash().decode() == nat.source_hash()
    mk = open(os.path.join(os.path.dirname(nat.__file__), "csrc", "Makefile")).read()
    assert re.search(r"^SRCS := (.*)$", mk, re.M).group(1).split() == nat.SOURCES


def test_product_kernels_contain_the_blackwell_instructions():
    """The library that ships (not a prototype) runs its contractions on tcgen05: the int8 digit-slice GEMMs and K1 hold
    UTCIMMA (tcgen05.mma kind::i8), LDTM (tcgen05.ld) and UBLKCP (bulk TMA copies); K5 streams L with UBLKCP; the FP64
    remainder is DMMA.  Counted from the SASS of the built library (tools/sass_summary.py; no GPU needed)."""
    import shutil
    import sys

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump is not installed")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import sass_summary
    finally:
        sys.path.pop(0)
    per, total = sass_summary.summarise(os.path.join(ROOT, "mellon_b200", "libmellon_b200.so"))
    for kernel in ("gram_i8_kernel<4>", "gemm_nt_i8_kernel<4>", "cov_i8_kernel<1, 0>", "cov_i8_kernel<6, 1>"):
        c = per[kernel]
        assert c["UTCIMMA"] == 28 and c["LDTM"] >= 7 and c["UBLKCP"] >= 1, (kernel, dict(c))
    assert per["gram_i8_kernel<0>"]["UTCCP"] == 7                      # the A-in-TMEM variant stages 7 slices by tcgen05.cp
    k5 = [n for n in per if n.startswith("stream_rows_kernel")]
    assert k5 and all(per[n]["UBLKCP"] >= 1 for n in k5)
    assert any(per[n]["DMMA"] > 0 for n in per if n.startswith("gemm_dmma_kernel"))
    assert total["UTCIMMA"] >= 13 * 28
