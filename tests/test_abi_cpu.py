"""CPU checks of the boundary: libgpx.so loads, exports every symbol include/gpx.h declares, and the
product path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os

import numpy as np
import pytest

from gaussian_process_b200 import _lib


def test_header_parses_and_library_exports_every_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 35
    for must in ("gpx_create", "gpx_cov_build", "gpx_potrf", "gpx_trsv", "gpx_trsm", "gpx_trtri", "gpx_lauum", "gpx_gemm",
                 "gpx_lml", "gpx_lml_grad", "gpx_gp_fit", "gpx_gp_fit_grad", "gpx_host_lml", "gpx_logistic_terms",
                 "gpx_build_B", "gpx_softmax_classes", "gpx_mg_fit_grad", "gpx_nccl_init", "gpx_gp_small_posterior_host",
                 "gpx_gp_small_fit_host", "gpx_gp_small_sample_host", "gpx_gp_small_lml_grad_host", "gpx_gp_small_ascent_host", "gpx_gp_small_prior_factor_host"):
        assert must in protos, must
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libgpx.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), "libgpx.so does not export %s" % name


def test_version_and_padding_without_gpu():
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libgpx.so not built")
    lib = _lib.load()
    assert lib.gpx_version() == 200
    assert lib.gpx_padded_dim(1) == 128 and lib.gpx_padded_dim(128) == 128 and lib.gpx_padded_dim(129) == 256
    assert lib.gpx_small_max() == 128


def test_host_pointer_entry_points_reject_bad_arguments_without_gpu():
    """Argument validation of the host-pointer calls happens before any CUDA work: NULL handle -> status -1."""
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("libgpx.so not built")
    lib = _lib.load()
    out = (ctypes.c_double * 8)()
    x = (ctypes.c_double * 4)(0.0, 1.0, 2.0, 3.0)
    th = (ctypes.c_double * 2)(1.0, 1.0)
    assert lib.gpx_gp_small_lml_grad_host(None, 0, x, 4, 1, x, th, 2, 5e-4, out, None) == -1
    assert lib.gpx_gp_small_ascent_host(None, x, 4, 1, x, 1.0, 1.0, 5e-4, 0.01, 1e-3, 10, out) == -1
    assert lib.gpx_gp_small_sample_host(None, 4, 1, x, out) == -1
    assert b"bad argument" in lib.gpx_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gaussian_process_b200 import GpxError, GP_regression as G
    with pytest.raises(GpxError):
        G.RBF_kernel(np.zeros((3, 1)), np.zeros((3, 1)), 1, 1)
    if os.path.isfile(_lib.LIB_PATH):
        lib = _lib.load()
        h = ctypes.c_void_p()
        assert lib.gpx_create(0, ctypes.byref(h)) < 0
        assert b"no CPU fallback" in lib.gpx_last_error() or b"CUDA" in lib.gpx_last_error()


def test_status_mapping():
    with pytest.raises(np.linalg.LinAlgError):
        _lib.check(7, "gpx_potrf")
    _lib.check(0)


def test_product_package_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "gaussian_process_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
