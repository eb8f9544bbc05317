"""User-defined cluster types (device functors through the C-ABI): registration and run-time compilation need no
GPU (NVRTC cross-compiles for sm_100a)."""
import pytest

import pmdi_b200  # noqa: F401
from pmdi_b200 import capi

import user_types as ut


def test_user_types_compile_for_sm100a():
    for name, src, kind in (("MyGaussian", ut.USER_GAUSSIAN, capi.F64), ("MyPoisson", ut.USER_POISSON, capi.I64)):
        tag = capi.register_cluster_type(name, src, name, kind)
        assert tag >= 16
        assert capi.register_cluster_type(name, src, name, kind) == tag   # same name: same tag
        capi.cluster_type_check(tag)


def test_compile_error_is_reported_with_the_compiler_log():
    tag = capi.register_cluster_type("Broken", ut.BROKEN, "Broken", capi.F64)
    with pytest.raises(capi.PmdiError) as e:
        capi.cluster_type_check(tag)
    assert "undefined_symbol" in str(e.value)
