"""dcl_net_b200 — B200-native (sm_100a) implementation of DCL-Net's data-parallel hot path.

Host side in Python/PyTorch (device memory, streams, autograd plumbing); all arithmetic of the
path in hand-written CUDA behind the C-ABI of include/dcl_b200.h (libdcl_b200.so, loaded by
dcl_net_b200._lib).  There is no CPU fallback: calling an op without the built library or
without a CUDA device raises.

Layout (mirrors the reference's import paths for this path):
  pointnet_lib.pointnet2_utils   libs/pointnet_lib/pointnet2_utils.py   (batched (B,N,3) ops)
  pointnet_sp.pointnet2_utils    libs/pointnet_sp/pointnet2_utils.py    (flat bxyz ops)
  modules                        models/Modules.py   (Aligner, heads, point-feature glue)
  dcl_net                        models/DCL_Net.py   (ortho9d2matrix, Network)
  refiner                        models/refiner.py   (Refiner, stage-2 loop)
  backbone                       models/Modules.py:100-159 + libs/spconv + libs/pointgroup_ops
                                 (device voxelisation and the two sparse-conv towers, inference)
  engine                         host-in / host-out inference engines (static buffers, CUDA graphs)
  sharding                       instance sharding across the GPUs of one box
"""
__version__ = "0.2.0"
