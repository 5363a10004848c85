"""TEST INFRASTRUCTURE — builds the reference's own CUDA kernels into oracle/_ref/.

The five kernel files of the reference hot path
    libs/pointnet_lib/src/{sampling,ball_query,group_points,interpolate}_gpu.cu
    libs/pointnet_sp/src/interpolate_gpu.cu
are compiled UNMODIFIED, from where they lie under /root/reference, with the reference's
own flag set (-O2, default -fmad) for sm_100a, into two shared objects (the two libraries
define same-named symbols):
    oracle/_ref/libref_pointnet_lib.so      oracle/_ref/libref_pointnet_sp.so
Their C++-mangled launcher symbols are called through ctypes by oracle/ref_kernels.py.
The reference's .cpp/pybind wrappers are NOT built: they include THC/THC.h, which modern
PyTorch no longer ships.  The kernel files include <torch/serialize/tensor.h> only for the
at::Tensor name in wrapper prototypes; oracle/stubs/ provides a 1-line stand-in so the
build takes seconds (no reference source is copied or edited).

Only runs where /root/reference exists (the build container).  The GPU box receives the
prebuilt .so files with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DCL_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

LIB_SRCS = ["sampling_gpu.cu", "ball_query_gpu.cu", "group_points_gpu.cu", "interpolate_gpu.cu"]
SP_SRCS = ["interpolate_gpu.cu"]


def _nvcc(srcs, src_dir, out_so):
    cmd = ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(HERE, "stubs"), "-I", src_dir,
           "-o", out_so] + [os.path.join(src_dir, s) for s in srcs]
    subprocess.run(cmd, check=True)


def build(force=False):
    """Returns True when both reference libraries exist after the call."""
    lib_so = os.path.join(OUT, "libref_pointnet_lib.so")
    sp_so = os.path.join(OUT, "libref_pointnet_sp.so")
    if not os.path.isdir(REF):
        return os.path.exists(lib_so) and os.path.exists(sp_so)
    os.makedirs(OUT, exist_ok=True)
    lib_dir = os.path.join(REF, "libs", "pointnet_lib", "src")
    sp_dir = os.path.join(REF, "libs", "pointnet_sp", "src")
    if force or not os.path.exists(lib_so):
        _nvcc(LIB_SRCS, lib_dir, lib_so)
    if force or not os.path.exists(sp_so):
        _nvcc(SP_SRCS, sp_dir, sp_so)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "unavailable (no /root/reference and no prebuilt files)")
