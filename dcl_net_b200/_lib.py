"""ctypes binding of libdcl_b200.so (C-ABI: include/dcl_b200.h).  Fails loudly."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdcl_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dcl_b200.h")

_I, _F, _P, _SZ, _I64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64

# activation / operand formats of the tensor-core kernels (include/dcl_b200.h: DCL_PM_FMT_*)
FMT_BF16X2, FMT_F16 = 0, 1

# name -> (restype, argtypes); must list every function include/dcl_b200.h declares
SIGNATURES = {
    "dcl_b200_abi_version": (_I, []),
    "dcl_b200_arch": (_I, []),
    "dcl_b200_launch_count": (ctypes.c_ulonglong, []),
    "dcl_lib_furthest_point_sampling_kernel_launcher": (_I, [_I, _I, _I, _P, _P, _P, _P]),
    "dcl_lib_gather_points_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "dcl_lib_gather_points_grad_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "dcl_lib_ball_query_kernel_launcher_fast": (_I, [_I, _I, _I, _F, _I, _P, _P, _P, _P]),
    "dcl_lib_group_points_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "dcl_lib_group_points_grad_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "dcl_lib_three_nn_kernel_launcher_fast": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_lib_knn_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_lib_three_interpolate_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_lib_three_interpolate_grad_kernel_launcher_fast": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_sp_three_nn_kernel_launcher_fast": (_I, [_I, _I, _P, _P, _P, _P, _P]),
    "dcl_sp_three_nn_workspace_bytes": (_SZ, [_I, _I]),
    "dcl_sp_three_nn_segmented": (_I, [_I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "dcl_sp_three_interpolate_kernel_launcher_fast": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_sp_three_interpolate_grad_kernel_launcher_fast": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_sp_nn_interpolate_fused": (_I, [_I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "dcl_fda_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "dcl_fda_align_fwd": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "dcl_fda_pack": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "dcl_fda_fwd_packed": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "dcl_fda_fwd_packed_jobs": (_I, [_I, _P, _I, _I, _I, _I, _I, _SZ, _P]),
    "dcl_fda_workspace_layout": (_I, [_I, _I, _I, _I, _I, _P]),
    "dcl_fda_fwd_packed_pm": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "dcl_fda_attention_map": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_svd3_project": (_I, [_I, _P, _I, _P, _P]),
    "dcl_weighted_kabsch": (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "dcl_pose_compose": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "dcl_pose_compose_pm": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "dcl_pm_gemm": (_I, [_I, _P, _I, _P]),
    "dcl_pm_pack_rows": (_I, [_I, _I, _I, _P, _P, _I, _P]),
    "dcl_pm_pack_cm": (_I, [_I, _I, _I, _P, _P, _I, _P]),
    "dcl_pm_unpack": (_I, [_I, _I, _P, _P, _I, _P]),
    "dcl_fda_fwd_packed_jobs_fmt": (_I, [_I, _P, _I, _I, _I, _I, _I, _SZ, _I, _P]),
    "dcl_fda_pack_fmt": (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _SZ, _I, _P]),
    "dcl_pose_compose_pm16": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "dcl_pm_pool_reduce": (_I, [_I, _I, _I, _P, _P, _P, _I, _P]),
    "dcl_sp_nn_interpolate_fused_pm": (_I, [_I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "dcl_sp_nn_interpolate_vox_pm": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "dcl_pose_head_workspace_bytes": (_SZ, [_I, _P, _P]),
    "dcl_pose_head": (_I, [_I, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "dcl_nearest_dist": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "dcl_conf_weights": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dcl_sp_levels_workspace_bytes": (_SZ, [_I, _P]),
    "dcl_sp_nn_interpolate_towers_pm": (_I, [_I, _P, _P, _SZ, _P]),
    "dcl_sp_nn_interpolate_levels_pm": (_I, [_I, _P, _I, _P, _P, _I, _P, _SZ, _P]),
    "dcl_spb_rows_per_instance": (_SZ, []),
    "dcl_spb_build_sets": (_I, [_I, _I, _I, _P, _F, _I, _P]),
    "dcl_spb_emit": (_I, [_I, _I, _P, _I, _P]),
    "dcl_spb_conv3": (_I, [_I, _I, _I, _I, _P, _P]),
    "dcl_spb_avgpool": (_I, [_I, _I, _I, _P, _P]),
    "dcl_voxelize_mean": (_I, [_I, _I, _I, _P, _P, _P, _P]),
    "dcl_tr_tile_pass": (_I, [_I, _P, _P]),
    "dcl_tr_bn_stats": (_I, [_I, _P, _P]),
    "dcl_tr_bn_bwd_reduce": (_I, [_I, _P, _P]),
    "dcl_tr_pack_weights": (_I, [_I, _P, _P]),
    "dcl_tr_colsum_reduce": (_I, [_I, _P, _P]),
    "dcl_fda_bwd": (_I, [_I, _P, _I, _I, _I, _I, _I, _P]),
    "dcl_debug_spconv_set_trace": (_I, [_P]),
    "dcl_debug_umma_gemm": (_I, [_I, _I, _P, _P, _P, _I, _P]),
    "dcl_debug_umma_pair_gemm": (_I, [_I, _I, _P, _P, _P, _I, _P]),
    "dcl_debug_fda_set_trace": (_I, [_P]),
    "dcl_debug_umma_ts_gemm": (_I, [_I, _I, _P, _P, _P, _I, _P]),
}



class PmGemmProblem(ctypes.Structure):
    """Mirror of dcl_pm_gemm_problem (include/dcl_b200.h)."""
    _fields_ = [("a0", _P), ("a1", _P), ("kb0", _I), ("kb_total", _I), ("w", _P), ("bias", _P), ("post_scale", _P),
                ("post_shift", _P), ("relu", _I), ("cout", _I), ("nt", _I), ("out_pm", _P), ("out_cm", _P),
                ("rows_per_inst", _I), ("pool_w", _P), ("pool_out", _P), ("dot_w", _P), ("dot_out", _P),
                ("out_qk", _P), ("qk_tile_rows", _I), ("out_v", _P), ("v_row0", _I), ("v_rows", _I),
                ("a_fmt", _I), ("out_fmt", _I), ("inst_count", _I), ("a_inst_stride", ctypes.c_longlong),
                ("w_inst_stride", ctypes.c_longlong), ("out_cm_inst_stride", ctypes.c_longlong)]


_LL = ctypes.c_longlong


class TrTile(ctypes.Structure):
    """Mirror of dcl_tr_tile (include/dcl_b200.h)."""
    _fields_ = [("x", _P), ("u", _P), ("x_sb", _LL), ("x_sc", _LL), ("x_sn", _LL), ("b", _I), ("c", _I), ("n", _I),
                ("mode", _I), ("scale", _P), ("shift", _P), ("mean", _P), ("rstd", _P), ("s1", _P), ("s2", _P),
                ("out_k", _P), ("out_t", _P), ("t_row0", _I), ("t_rows", _I), ("out_cm", _P), ("col_partial", _P),
                ("k_col0", _I), ("k_cols", _I), ("t_group", _I)]


class TrBn(ctypes.Structure):
    """Mirror of dcl_tr_bn (include/dcl_b200.h)."""
    _fields_ = [("u", _P), ("b", _I), ("c", _I), ("n", _I), ("gamma", _P), ("beta", _P), ("eps", _F), ("momentum", _F),
                ("running_mean", _P), ("running_var", _P), ("mean", _P), ("rstd", _P), ("scale", _P), ("shift", _P)]


class TrBnBwd(ctypes.Structure):
    """Mirror of dcl_tr_bn_bwd (include/dcl_b200.h)."""
    _fields_ = [("dy", _P), ("u", _P), ("dy_sb", _LL), ("dy_sc", _LL), ("b", _I), ("c", _I), ("n", _I), ("mode", _I),
                ("mean", _P), ("rstd", _P), ("scale", _P), ("shift", _P), ("s1", _P), ("s2", _P)]


class TrWpack(ctypes.Structure):
    """Mirror of dcl_tr_wpack (include/dcl_b200.h)."""
    _fields_ = [("src", _P), ("dst", _P), ("rows", _I), ("cols", _I), ("rows_pad", _I), ("k_pad", _I), ("nt", _I),
                ("transpose", _I), ("k_col0", _I), ("k_total", _I)]


class TrColsum(ctypes.Structure):
    """Mirror of dcl_tr_colsum (include/dcl_b200.h)."""
    _fields_ = [("partial", _P), ("out", _P), ("parts", _I), ("c", _I)]


class FdaBwdJob(ctypes.Structure):
    """Mirror of dcl_fda_bwd_job (include/dcl_b200.h)."""
    _fields_ = [(k, _P) for k in ("q_k", "k_k", "g_k", "v_k", "q_t", "k_t", "g_t", "lse", "dsum", "d_q", "d_k", "d_v")]


class FdaJob(ctypes.Structure):
    """Mirror of dcl_fda_job (include/dcl_b200.h)."""
    _fields_ = [("workspace", _P), ("RE_embed", _P), ("RI_embed", _P), ("RE_pm", _P), ("RI_pm", _P), ("lse", _P)]


class SpLevel(ctypes.Structure):
    """Mirror of dcl_sp_level (include/dcl_b200.h)."""
    _fields_ = [("m", _I), ("c", _I), ("out_col0", _I), ("grid_x", _I), ("vox_indices", _P), ("voxel_extent", ctypes.c_float * 3),
                ("offset", ctypes.c_float * 3), ("feats", _P)]


class SpTower(ctypes.Structure):
    """Mirror of dcl_sp_tower (include/dcl_b200.h)."""
    _fields_ = [("n", _I), ("c_total", _I), ("nlevels", _I), ("unknown", _P), ("out_pm", _P), ("levels", _P),
                ("out_fmt", _I)]


class SpbTowerIn(ctypes.Structure):
    """Mirror of dcl_spb_tower_in (include/dcl_b200.h)."""
    _fields_ = [("points", _P), ("rgb", _P), ("coords", _P), ("rows", _P), ("prefix", _P), ("counts", _P),
                ("feat16", _P), ("feat32", _P), ("occupied", _P), ("p2v", _P), ("v2p_sorted", _P), ("v2p_start", _P),
                ("errors", _P)]


class SpbTowerSets(ctypes.Structure):
    """Mirror of dcl_spb_tower_sets (include/dcl_b200.h)."""
    _fields_ = [("rows", _P), ("prefix", _P), ("counts", _P), ("offsets", _P), ("indices", _P * 9), ("cap", _I * 9),
                ("nbr", _P * 12), ("anymask", _P * 12), ("errors", _P)]


class SpbConv(ctypes.Structure):
    """Mirror of dcl_spb_conv (include/dcl_b200.h)."""
    _fields_ = [("in16", _P), ("nbr", _P), ("anymask", _P), ("offsets_out", _P), ("w", _P), ("shift", _P),
                ("out16", _P), ("out32", _P), ("cap_out", _I)]


class SpbPool(ctypes.Structure):
    """Mirror of dcl_spb_pool (include/dcl_b200.h)."""
    _fields_ = [("inp", _P), ("nbr", _P), ("offsets_out", _P), ("out32", _P), ("out16", _P), ("cap_out", _I)]


class PoseHeadMlp(ctypes.Structure):
    """Mirror of dcl_pose_head_mlp (include/dcl_b200.h)."""
    _fields_ = [("w1", _P), ("b1", _P), ("w2", _P), ("b2", _P), ("w3", _P), ("b3", _P),
                ("d_in", _I), ("d_h1", _I), ("d_h2", _I), ("d_out", _I)]


_lib = None


def load():
    """Load the shared library (no CUDA needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C dcl_net_b200/csrc`).  dcl_net_b200 has no CPU / PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.dcl_b200_abi_version() != 5:
            raise RuntimeError("libdcl_b200.so ABI version mismatch with dcl_net_b200/_lib.py")
        _lib = lib
    return _lib


def stream_ptr():
    """The current CUDA stream of the current device as a void* (the raw-handle query: torch.cuda.current_stream()
    builds a Stream object and costs ~6 us, which adds up over the ~400 launches of a training step)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def fill_structs(struct, items):
    """ctypes array of `struct` from a list of field dicts (missing pointer fields = NULL, missing numbers = 0;
    tensors must be contiguous CUDA tensors and are passed by address).  Built positionally: a setattr per field is
    several times slower, and these arrays are rebuilt for every launch."""
    spec = _STRUCT_SPECS.get(struct)
    if spec is None:
        spec = _STRUCT_SPECS[struct] = [(name, typ is _P) for name, typ in struct._fields_]
    rows = []
    for f in items:
        vals = []
        for name, is_ptr in spec:
            v = f.get(name)
            if v is None:
                vals.append(None if is_ptr else 0)
            elif isinstance(v, torch.Tensor):
                if not (v.is_cuda and v.is_contiguous()):
                    raise RuntimeError(f"dcl_net_b200: field {name} needs a contiguous CUDA tensor")
                vals.append(v.data_ptr())
            else:
                vals.append(v)
        rows.append(struct(*vals))
    return (struct * len(rows))(*rows)


_STRUCT_SPECS = {}


def ptr(t):
    """Device pointer of a tensor that must already be a contiguous CUDA tensor."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("dcl_net_b200 ops need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("dcl_net_b200 ops need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


_warned = set()


def warn_once(key, message):
    """One RuntimeWarning per process and key: used where an inference call leaves the tensor-core path for the
    PyTorch layer modules (still CUDA, never CPU) so that the switch is never silent."""
    if key not in _warned:
        _warned.add(key)
        import warnings
        warnings.warn(message, RuntimeWarning, stacklevel=3)


def check(err, what):
    if err != 0:
        raise RuntimeError(f"{what} failed: cudaError {err}")


def require(t, dtype, what):
    if t.dtype != dtype:
        raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
    return t
