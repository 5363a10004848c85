"""TEST INFRASTRUCTURE — ctypes front end of oracle/neighbour_oracle.c (numpy in / out).

Each function mirrors the Python-level contract of the reference wrapper it stands for
(libs/pointnet_lib/pointnet2_utils.py, libs/pointnet_sp/pointnet2_utils.py): caller-visible
allocation conventions (temp = 1e10, zeroed idx / grad buffers) are reproduced here so the
tests read like calls of the reference ops.  dist outputs are SQUARED distances (the
reference applies sqrt in Python, pointnet2_utils.py:134).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "neighbour_oracle.c")
_SO = os.path.join(_HERE, "_build", "liboracle_cpu.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fvisibility=hidden", "-shared", "-fPIC",
                        "-o", _SO, _SRC, "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def opt_n_threads(n):
    return int(lib().oracle_opt_n_threads(int(n)))


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz, px = _f(xyz)
    B, N, _ = xyz.shape
    temp, pt = _f(np.full((B, N), 1e10, np.float32))
    out, po = _i(np.zeros((B, npoint), np.int32))
    lib().oracle_furthest_point_sampling(B, N, int(npoint), px, pt, po)
    return (out, temp) if return_temp else out


def gather_operation(features, idx):
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    m = idx.shape[1]
    out, po = _f(np.empty((B, C, m), np.float32))
    lib().oracle_gather_points(B, C, N, m, pf, pi, po)
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, m = grad_out.shape
    out, po = _f(np.zeros((B, C, N), np.float32))
    lib().oracle_gather_points_grad(B, C, N, m, pg, pi, po)
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz)
    new_xyz, pn = _f(new_xyz)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx, pi = _i(np.zeros((B, m, nsample), np.int32))
    lib().oracle_ball_query(B, N, m, ctypes.c_float(radius), int(nsample), pn, px, pi)
    return idx


def grouping_operation(features, idx):
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    _, npoint, nsample = idx.shape
    out, po = _f(np.empty((B, C, npoint, nsample), np.float32))
    lib().oracle_group_points(B, C, N, npoint, nsample, pf, pi, po)
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, npoint, nsample = grad_out.shape
    out, po = _f(np.zeros((B, C, N), np.float32))
    lib().oracle_group_points_grad(B, C, N, npoint, nsample, pg, pi, po)
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2, pd = _f(np.empty((B, n, 3), np.float32))
    idx, pi = _i(np.empty((B, n, 3), np.int32))
    lib().oracle_three_nn(B, n, m, pu, pk, pd, pi)
    return d2, idx


def knn(k, unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2, pd = _f(np.empty((B, n, k), np.float32))
    idx, pi = _i(np.empty((B, n, k), np.int32))
    lib().oracle_knn(B, n, m, int(k), pu, pk, pd, pi)
    return d2, idx


def three_interpolate(features, idx, weight):
    features, pf = _f(features)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out, po = _f(np.empty((B, C, n), np.float32))
    lib().oracle_three_interpolate(B, C, m, n, pf, pi, pw, po)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, n = grad_out.shape
    out, po = _f(np.zeros((B, C, m), np.float32))
    lib().oracle_three_interpolate_grad(B, C, n, m, pg, pi, pw, po)
    return out


def sp_three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    n, m = unknown.shape[0], known.shape[0]
    d2, pd = _f(np.empty((n, 3), np.float32))
    idx, pi = _i(np.empty((n, 3), np.int32))
    lib().oracle_sp_three_nn(n, m, pu, pk, pd, pi)
    return d2, idx


def sp_three_nn_slab_model(unknown, vox, ext, off, gx):
    """Model of the product's slab-walk search (see neighbour_oracle.c); returns (dist2, idx, candidates visited)."""
    import ctypes
    unknown, pu = _f(unknown)
    vox, pv = _i(vox)
    ext, pe = _f(np.asarray(ext, np.float32))
    off, po = _f(np.asarray(off, np.float32))
    n, m = unknown.shape[0], vox.shape[0]
    d2, pd = _f(np.empty((n, 3), np.float32))
    idx, pi = _i(np.empty((n, 3), np.int32))
    visited = ctypes.c_long(0)
    lib().oracle_sp_three_nn_slab_model(n, m, int(gx), pu, pv, pe, po, pd, pi, ctypes.byref(visited))
    return d2, idx, visited.value


def sp_three_interpolate(features, idx, weight):
    features, pf = _f(features)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    m, c = features.shape
    n = idx.shape[0]
    out, po = _f(np.empty((n, c), np.float32))
    lib().oracle_sp_three_interpolate(c, m, n, pf, pi, pw, po)
    return out


def sp_three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    n, c = grad_out.shape
    out, po = _f(np.zeros((m, c), np.float32))
    lib().oracle_sp_three_interpolate_grad(c, n, m, pg, pi, pw, po)
    return out
