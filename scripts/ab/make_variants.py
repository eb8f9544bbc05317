#!/usr/bin/env python
"""Layout experiment: the device functions of a kernel are placed in the order of their mangled names, and the
per-step time of the persistent sweep kernel depends on which functions share instruction-cache sets.
Builds variants of the library with hot functions renamed (10-character identifiers sort first) into
scripts/ab/lib_<variant>.so;  scripts/gpu_ab.sh times them on one box."""
import os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "particlemdi.jl_b200", "csrc")
P = ["spec_gsync", "spec_propose", "spec_commit", "spec_wait_decision", "pm_uniform", "pm_exp", "pm_log", "pm_div"]
E = ["spec_fix", "spec_eval", "pool_issue_obs", "gauss_block", "nb_block", "cat_block", "pm_lfact_stirling"]
def names(order, base=0):
    return {f: "h%02d_%s" % (base + i, (re.sub("[^a-z]", "", f) + "xxxxxx")[:6]) for i, f in enumerate(order)}
VARIANTS = {
    "head": {},
    "pullend": {"spec_row_pull": "zrow_pull"},
    "pfirst": {**names(P + E), "spec_row_pull": "zrow_pull"},
    "efirst": {**names(E + P), "spec_row_pull": "zrow_pull"},
    "ponly": {**names(P), "spec_row_pull": "zrow_pull"},
}
for v, ren in VARIANTS.items():
    d = os.path.join(ROOT, "gpurun_out", "ab_build", v)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(os.path.join(d, "pkg", "csrc")); os.makedirs(os.path.join(d, "include"))
    shutil.copy(os.path.join(ROOT, "include", "pmdi_cuda.h"), os.path.join(d, "include"))
    for f in os.listdir(SRC):
        if not f.endswith((".cu", ".cuh", ".h", ".inc")):
            continue
        s = open(os.path.join(SRC, f)).read()
        if not f.endswith(".inc"):
            for a, b in ren.items():
                assert len(b) in (9, 10), b
                s = re.sub(r"\b%s\b" % a, b, s)
        open(os.path.join(d, "pkg", "csrc", f), "w").write(s)
    out = os.path.join(ROOT, "scripts", "ab", "lib_%s.so" % v)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                           "-fPIC", "-shared", "-w", "-o", out, "pmdi_cuda.cu"], cwd=os.path.join(d, "pkg", "csrc"))
    print("built", out, flush=True)
