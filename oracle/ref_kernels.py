"""TEST INFRASTRUCTURE — the reference's OWN CUDA kernels as a parity oracle (GPU box only).

oracle/_ref/libref_pointnet_{lib,sp}.so hold the reference's five *_gpu.cu files compiled
unmodified (oracle/build_ref.py).  Their launchers are C++ symbols; we bind the mangled
names through ctypes and pass torch CUDA tensors' data_ptr() plus the current stream,
reproducing what the reference's pybind wrappers did (libs/pointnet_lib/src/*.cpp) and what
its Python Functions allocate (libs/pointnet_lib/pointnet2_utils.py).
All dist outputs are SQUARED distances.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_SO = os.path.join(_HERE, "_ref", "libref_pointnet_lib.so")
_SP_SO = os.path.join(_HERE, "_ref", "libref_pointnet_sp.so")
_lib = _sp = None

_SYM = {
    "fps": "_Z39furthest_point_sampling_kernel_launcheriiiPKfPfPiP11CUstream_st",
    "gather": "_Z34gather_points_kernel_launcher_fastiiiiPKfPKiPfP11CUstream_st",
    "gather_grad": "_Z39gather_points_grad_kernel_launcher_fastiiiiPKfPKiPfP11CUstream_st",
    "ball_query": "_Z31ball_query_kernel_launcher_fastiiifiPKfS0_PiP11CUstream_st",
    "group": "_Z33group_points_kernel_launcher_fastiiiiiPKfPKiPfP11CUstream_st",
    "group_grad": "_Z38group_points_grad_kernel_launcher_fastiiiiiPKfPKiPfP11CUstream_st",
    "three_nn": "_Z29three_nn_kernel_launcher_fastiiiPKfS0_PfPiP11CUstream_st",
    "knn": "_Z24knn_kernel_launcher_fastiiiiPKfS0_PfPiP11CUstream_st",
    "interp": "_Z38three_interpolate_kernel_launcher_fastiiiiPKfPKiS0_PfP11CUstream_st",
    "interp_grad": "_Z43three_interpolate_grad_kernel_launcher_fastiiiiPKfPKiS0_PfP11CUstream_st",
    "sp_three_nn": "_Z29three_nn_kernel_launcher_fastiiPKfS0_PfPiP11CUstream_st",
    "sp_interp": "_Z38three_interpolate_kernel_launcher_fastiiiPKfPKiS0_PfP11CUstream_st",
    "sp_interp_grad": "_Z43three_interpolate_grad_kernel_launcher_fastiiiPKfPKiS0_PfP11CUstream_st",
}


def available():
    return os.path.exists(_LIB_SO) and os.path.exists(_SP_SO)


def _libs():
    global _lib, _sp
    if _lib is None:
        _lib = ctypes.CDLL(_LIB_SO)
        _sp = ctypes.CDLL(_SP_SO)
    return _lib, _sp


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _call(libidx, key, *args):
    fn = getattr(_libs()[libidx], _SYM[key])
    fn.restype = None
    fn(*args)


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    _call(0, "fps", B, N, int(npoint), _p(xyz), _p(temp), _p(out), _st())
    return (out, temp) if return_temp else out


def gather_operation(features, idx):
    features, idx = features.contiguous(), idx.contiguous()
    B, C, N = features.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m, dtype=torch.float32, device=features.device)
    _call(0, "gather", B, C, N, m, _p(features), _p(idx), _p(out), _st())
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, idx = grad_out.contiguous(), idx.contiguous()
    B, C, m = grad_out.shape
    out = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _call(0, "gather_grad", B, C, N, m, _p(grad_out), _p(idx), _p(out), _st())
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = xyz.contiguous(), new_xyz.contiguous()
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(B, m, nsample, dtype=torch.int32, device=xyz.device)
    _call(0, "ball_query", B, N, m, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx), _st())
    return idx


def grouping_operation(features, idx):
    features, idx = features.contiguous(), idx.contiguous()
    B, C, N = features.shape
    _, npoint, nsample = idx.shape
    out = torch.empty(B, C, npoint, nsample, dtype=torch.float32, device=features.device)
    _call(0, "group", B, C, N, npoint, nsample, _p(features), _p(idx), _p(out), _st())
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, idx = grad_out.contiguous(), idx.contiguous()
    B, C, npoint, nsample = grad_out.shape
    out = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _call(0, "group_grad", B, C, N, npoint, nsample, _p(grad_out), _p(idx), _p(out), _st())
    return out


def three_nn(unknown, known):
    unknown, known = unknown.contiguous(), known.contiguous()
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    _call(0, "three_nn", B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _st())
    return d2, idx


def knn(k, unknown, known):
    unknown, known = unknown.contiguous(), known.contiguous()
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, k, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, k, dtype=torch.int32, device=unknown.device)
    _call(0, "knn", B, n, m, int(k), _p(unknown), _p(known), _p(d2), _p(idx), _st())
    return d2, idx


def three_interpolate(features, idx, weight):
    features, idx, weight = features.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, m = features.shape
    n = idx.shape[1]
    out = torch.empty(B, C, n, dtype=torch.float32, device=features.device)
    _call(0, "interp", B, C, m, n, _p(features), _p(idx), _p(weight), _p(out), _st())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = grad_out.contiguous(), idx.contiguous(), weight.contiguous()
    B, C, n = grad_out.shape
    out = torch.zeros(B, C, m, dtype=torch.float32, device=grad_out.device)
    _call(0, "interp_grad", B, C, n, m, _p(grad_out), _p(idx), _p(weight), _p(out), _st())
    return out


def sp_three_nn(unknown, known):
    unknown, known = unknown.contiguous(), known.contiguous()
    n, m = unknown.shape[0], known.shape[0]
    d2 = torch.empty(n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(n, 3, dtype=torch.int32, device=unknown.device)
    _call(1, "sp_three_nn", n, m, _p(unknown), _p(known), _p(d2), _p(idx), _st())
    return d2, idx


def sp_three_interpolate(features, idx, weight):
    features, idx, weight = features.contiguous(), idx.contiguous(), weight.contiguous()
    m, c = features.shape
    n = idx.shape[0]
    out = torch.empty(n, c, dtype=torch.float32, device=features.device)
    _call(1, "sp_interp", c, m, n, _p(features), _p(idx), _p(weight), _p(out), _st())
    return out


def sp_three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = grad_out.contiguous(), idx.contiguous(), weight.contiguous()
    n, c = grad_out.shape
    out = torch.zeros(m, c, dtype=torch.float32, device=grad_out.device)
    _call(1, "sp_interp_grad", c, n, m, _p(grad_out), _p(idx), _p(weight), _p(out), _st())
    return out
