"""Build the sm_100a shared library `lib/librze_b200.so` in-tree with nvcc.

`python -m reze_engine_b200.build` (or `__graft_entry__.build()`); cross-compiles
without a GPU.  One object per feature set of the deform kernel so the
instantiations compile in parallel.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "librze_b200.so")
# keep in sync with csrc/kernel_table.h RZ_FEAT_LIST
FEATS = [0, 1, 3, 4, 7, 16, 19, 20, 23, 8, 11, 15, 24, 31, 32, 39, 40, 47, 64, 71, 72, 79]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{r.stdout[-4000:]}")
    return r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in ("deform_kernel.cuh", "deform2_kernel.cuh", "kernel_table.h", "aux_kernels.cuh", "lane_plan.h", "lane_plan2.h", "mesh_tables.h")]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "rze_b200.h"))
    jobs = []
    objs = []
    api_src = os.path.join(CSRC, "rze_b200.cu")
    api_obj = os.path.join(OBJDIR, "rze_b200.o")
    objs.append(api_obj)
    if force or not _newer(api_obj, [api_src] + hdrs):
        jobs.append(([nvcc, *ARCH, *COMMON, "-c", api_src, "-o", api_obj], api_obj + ".log"))
    inst_src = os.path.join(CSRC, "deform_inst.cu")
    for f in FEATS:
        o = os.path.join(OBJDIR, f"deform_feat{f}.o")
        objs.append(o)
        if force or not _newer(o, [inst_src] + hdrs):
            jobs.append(([nvcc, *ARCH, *COMMON, f"-DRZ_FEAT={f}", "-c", inst_src, "-o", o], o + ".log"))
    v2_src = os.path.join(CSRC, "deform2_inst.cu")
    for out2 in range(5):                                    # two-vertex kernel, one object per output layout (OUT2_*)
        v2_obj = os.path.join(OBJDIR, "deform2.o" if out2 == 0 else f"deform2_out{out2}.o")
        objs.append(v2_obj)
        if force or not _newer(v2_obj, [v2_src] + hdrs):
            jobs.append(([nvcc, *ARCH, *COMMON, f"-DRZ_OUT2={out2}", "-c", v2_src, "-o", v2_obj], v2_obj + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            outs = list(ex.map(lambda j: _run(*j), jobs))
        if verbose:
            for o in outs:
                print(o)
    if jobs or not os.path.exists(LIB):
        _run([nvcc, *ARCH, "-shared", "-o", LIB, *objs], LIB + ".log")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
