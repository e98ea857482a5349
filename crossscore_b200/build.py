"""Build recipe for libcrossscore_sm100a.so (nvcc, sm_100a only, in-tree).

`python -m crossscore_b200.build` or `__graft_entry__.build()`.  The library is linked against the
static CUDA runtime and resolves the driver's cuTensorMapEncodeTiled at run time through
cudaGetDriverEntryPoint, so it loads (and its symbols can be checked) on a machine without a GPU driver.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcrossscore_sm100a.so")
SOURCES = ["xs_api.cu", "xs_gemm_tc.cu", "xs_attn_tc.cu", "xs_attn_tc2.cu", "xs_rows.cu", "xs_f32.cu", "xs_imgproc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(CSRC, "xs_common.cuh"), os.path.join(CSRC, "xs_turbo.h"),
            os.path.join(HERE, "..", "include", "crossscore_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """XS_BUILD_DEFS="-DATT_X=1 ..." + XS_BUILD_TAG=name build a development variant into
    libcrossscore_sm100a_<name>.so (load it with XS_LIB_PATH) without touching the product library."""
    tag, defs = os.environ.get("XS_BUILD_TAG", ""), os.environ.get("XS_BUILD_DEFS", "").split()
    if tag:
        return _build(os.path.join(HERE, "build_" + tag), LIB.replace(".so", f"_{tag}.so"), defs, True, verbose)
    return _build(os.path.join(HERE, "build"), LIB, [], force, verbose)


def _build(objdir: str, LIB: str, defs, force: bool, verbose: bool) -> str:
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = [NVCC, *FLAGS, *defs, "-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rebuilt = bool(procs)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        with open(os.path.join(objdir, s + ".log"), "w") as f:
            f.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    if rebuilt or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
