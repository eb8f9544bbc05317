#!/usr/bin/env python
"""Which of the last changes slowed the single-GPU sweep by 13 %?  Variants of HEAD with one change taken back each."""
import os, re, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "particlemdi.jl_b200", "csrc")
OLD = os.path.join(ROOT, "gpurun_out", "old_src", "csrc")
def block(s, a, b):
    i0 = s.index(a); i1 = s.index(b, i0); return i0, i1
new_spec = open(os.path.join(SRC, "spec_kernel.cuh")).read()
old_spec = open(os.path.join(OLD, "spec_kernel.cuh")).read()
A_PULL, B_PULL = "// A live row of another rank's pool", "// ------------------------------------------------------------------------------------------------\n// Resampling after step `st`"
A_PULL_OLD = "// One warp copies a live row of another rank's pool"
A_LOOP, B_LOOP = "    for (long long item = gw; item < njobs * SPEC_PULL_PARTS; item += GWp) {", "    if (gt == 0) sp.counters[3] += njobs;"
A_LOOP_OLD = "    for (long long job = gw; job < njobs; job += GWp) {"
A_DEC, B_DEC = "    if (sp.R > 1) {\n      // One (max, sum w, sum w^2)", "    if (lane == 0) {\n      const int res"
A_DEC_OLD = "    if (sp.R > 1) {  // one (max, sum w, sum w^2) per rank"
def swap(s, a_new, b, a_old):
    i0, i1 = block(s, a_new, b); j0, j1 = block(old_spec, a_old, b)
    return s[:i0] + old_spec[j0:j1] + s[i1:]
v_oldpull = swap(swap(new_spec, A_PULL, B_PULL, A_PULL_OLD), A_LOOP, B_LOOP, A_LOOP_OLD)
v_olddec = swap(new_spec, A_DEC, B_DEC, A_DEC_OLD)
cu = open(os.path.join(SRC, "pmdi_cuda.cu")).read()
cu_arena = cu.replace("take(sizeof(double) * (2 * 8 * 4 + 2 * 8 * 8));", "take(sizeof(double) * 2 * 8 * 4);")
assert cu_arena != cu
a = "  if (c->arena) { cudaFree(c->arena); c->arena = nullptr; }"
assert cu_arena.count(a) == 1
cu_arena = cu_arena.replace(a, "  const size_t o_rankw = take(sizeof(double) * 2 * 8 * 8);\n" + a)
cu_arena = cu_arena.replace("  sp.rank_words = (unsigned long long*)(c->rank_part.p + 2 * 8 * 4);", "  sp.rank_words = (unsigned long long*)(c->arena + c->o_rankw);")
cu_arena = cu_arena.replace("  c->rank_part.view(A + o_rankp, 2 * 8 * 4 + 2 * 8 * 8);", "  c->rank_part.view(A + o_rankp, 2 * 8 * 4);\n  c->o_rankw = o_rankw;")
cu_arena = cu_arena.replace("  unsigned step_seq = 0;", "  size_t o_rankw = 0;\n  unsigned step_seq = 0;")
VARIANTS = {"oldpull": {"spec_kernel.cuh": v_oldpull}, "olddec": {"spec_kernel.cuh": v_olddec}, "arena": {"pmdi_cuda.cu": cu_arena}}
for v, files in VARIANTS.items():
    d = os.path.join(ROOT, "gpurun_out", "ab_build", v)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(os.path.join(d, "pkg", "csrc")); os.makedirs(os.path.join(d, "include"))
    shutil.copy(os.path.join(ROOT, "include", "pmdi_cuda.h"), os.path.join(d, "include"))
    for f in os.listdir(SRC):
        if f.endswith((".cu", ".cuh", ".h", ".inc")):
            s = files.get(f) or open(os.path.join(SRC, f)).read()
            open(os.path.join(d, "pkg", "csrc", f), "w").write(s)
    out = os.path.join(ROOT, "scripts", "ab", "lib_%s.so" % v)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                           "-fPIC", "-shared", "-w", "-o", out, "pmdi_cuda.cu"], cwd=os.path.join(d, "pkg", "csrc"))
    print("built", out, flush=True)
