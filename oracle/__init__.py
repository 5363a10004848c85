"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.

CPU / reference-side checkers for the DCL-Net hot path:

  cpu_oracle.py    ctypes front end of neighbour_oracle.c — a single-threaded C restatement
                   of the reference's CUDA-only neighbourhood kernels (bit-faithful).
  torch_oracle.py  pure-PyTorch fp32/fp64 restatements of the FDA head, the SVD pose
                   projection (ortho9d2matrix), the refiner loop and weighted Kabsch.
  ref_kernels.py   ctypes front end of oracle/_ref/*.so — the reference's own kernels,
                   compiled unmodified for sm_100a by build_ref.py (GPU box only).
  make_golden.py   imports the reference's Python modules from /root/reference (build
                   container only) and writes the fixtures under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (dcl_net_b200/) never does.
"""
