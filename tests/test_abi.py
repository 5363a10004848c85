"""The C-ABI library loads on a CPU-only host and exports every function include/dcl_b200.h declares
(no compute calls here)."""
import ctypes
import re

from dcl_net_b200 import _lib


def _declared():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dcl_[a-z0-9_]+)\s*\(", src))


def test_header_and_binding_table_agree():
    declared = _declared()
    assert len(declared) >= 24
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(_declared()):
        assert hasattr(lib, name), f"libdcl_b200.so lacks {name}"


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in _lib.py have the size and field offsets a C compiler gives the structs of the header."""
    import subprocess
    structs = {"dcl_pm_gemm_problem": _lib.PmGemmProblem, "dcl_pose_head_mlp": _lib.PoseHeadMlp,
               "dcl_sp_level": _lib.SpLevel, "dcl_sp_tower": _lib.SpTower, "dcl_fda_job": _lib.FdaJob,
               "dcl_tr_tile": _lib.TrTile, "dcl_tr_bn": _lib.TrBn, "dcl_tr_bn_bwd": _lib.TrBnBwd,
               "dcl_tr_wpack": _lib.TrWpack, "dcl_tr_colsum": _lib.TrColsum, "dcl_fda_bwd_job": _lib.FdaBwdJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{_lib.HEADER_PATH}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_version_and_arch():
    lib = _lib.load()
    assert lib.dcl_b200_abi_version() == 5
    assert lib.dcl_b200_arch() == 100


def test_workspace_queries_are_host_only():
    lib = _lib.load()
    assert lib.dcl_fda_workspace_bytes(32, 64, 256, 1024, 1024) == 32 * (1024 * 64 * 2 + 1024 * 320) * 4 + 1024
    assert lib.dcl_fda_workspace_bytes(1, 96, 256, 1024, 1024) == 0      # unsupported width
    assert lib.dcl_sp_three_nn_workspace_bytes(10, 1000) > 16 * 1000


def test_no_cpu_fallback():
    """Ops refuse CPU tensors instead of silently computing somewhere else."""
    import pytest
    import torch
    from dcl_net_b200.modules import fda_align
    with pytest.raises(RuntimeError):
        fda_align(torch.zeros(1, 64, 128), torch.zeros(1, 64, 64), torch.zeros(1, 256, 64))


def test_product_does_not_import_oracle():
    """The product never imports, links or executes anything under oracle/ (test infrastructure only)."""
    import os
    root = os.path.dirname(_lib.LIB_PATH)
    py = re.compile(r"^\s*(import|from)\s+oracle\b|oracle[/.](cpu_oracle|torch_oracle|ref_kernels|_ref|_build)", re.M)
    inc = re.compile(r"#\s*include[^\n]*oracle")
    for dirpath, _, files in os.walk(root):
        for f in files:
            text = None
            if f.endswith(".py"):
                text, pat = open(os.path.join(dirpath, f)).read(), py
            elif f.endswith((".cu", ".cuh", ".h", "Makefile")):
                text, pat = open(os.path.join(dirpath, f)).read(), inc
            if text is not None:
                assert not pat.search(text), f"{f} uses the oracle"
