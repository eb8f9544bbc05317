"""The C-ABI library (no GPU needed): it builds for sm_100a, loads, exports every symbol that
include/pmdi_cuda.h declares, and fails LOUDLY - never falls back - when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pmdi_b200  # noqa: F401
from pmdi_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pmdi_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmdi_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    capi.build()
    L = capi.lib()
    names = _declared()
    assert len(names) >= 17
    for s in names:
        assert hasattr(L, s), f"libpmdi_cuda.so does not export {s}"
    assert sorted(capi.EXPORTS) == names  # the ctypes binding covers the whole header
    assert L.pmdi_version() >= 100


def test_struct_layouts_match_header():
    """Field order of the ctypes mirrors = field order in the header."""
    src = open(os.path.join(ROOT, "include", "pmdi_cuda.h")).read()
    for cname, cls in (("pmdi_sweep_args", capi.SweepArgs), ("pmdi_sweep_out", capi.SweepOut)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = [re.sub(r"\[.*\]", "", d.split()[-1].lstrip("*")) for d in body.split(";") if d.strip()]
        assert fields == [f[0] for f in cls._fields_], cname


def test_no_silent_fallback_without_gpu():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = capi.lib().pmdi_ctx_create(C.byref(h), 1, 10, 3, 4, 0)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in capi.lib().pmdi_last_error()
    with pytest.raises(capi.PmdiError):
        capi.Context([np.zeros((10, 2))], [capi.GAUSSIAN], 3, 4)


@pytest.mark.parametrize("K,n,N,P", [(0, 10, 3, 4), (9, 10, 3, 4), (1, 10, 1, 4), (1, 10, 11, 4),
                                     (1, 10, 3, 1), (1, 1000, 300, 4)])
def test_ctx_create_preconditions(K, n, N, P):
    """The asserts of pmdi() (src/pmdi.jl:50-55) are checked before any device is touched."""
    h = C.c_void_p()
    rc = capi.lib().pmdi_ctx_create(C.byref(h), K, n, N, P, 0)
    assert rc == 1 and not h.value
    assert capi.lib().pmdi_last_error()


def test_uniform_is_the_oracles_philox():
    from oracle import oracle as orc
    for args in [(1, 0, 0, 0, 0, 1), (2 ** 40 + 5, 7, 2, 300, 0, 99), (42, 3, 4, 0, 2, 1999)]:
        assert capi.uniform(*args) == orc.uniform(*args)


def test_sweep_kernel_resources_stay_within_the_measured_build():
    """The persistent sweep kernel is bound by per-step latency; a change in code it never executes moved every step
    by 13 % once, through the whole-kernel register allocation (stack frame 352 -> 848 bytes,
    profiles/r02_spec_kernel.md).  Guard: the production instantiation keeps 128 registers (one 512-thread CTA per SM)
    and a stack frame no larger than the measured build's."""
    import re
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "--dump-resource-usage", capi.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(r"Function k_sweep_spec:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "k_sweep_spec not found in the library"
    regs, stack = int(m.group(1)), int(m.group(2))
    assert regs <= 128
    assert stack <= 384, f"k_sweep_spec stack frame {stack} B (measured build: 352 B): time it A/B before shipping (scripts/gpu_ab.sh)"
