// Stand-in for <torch/serialize/tensor.h>, used ONLY to compile the reference's *_gpu.cu
// kernel files (which include it through their headers but use no torch symbol in device
// or launcher code).  It lets oracle/build_ref.py build them in seconds instead of
// minutes.  The reference sources themselves are compiled unmodified, where they lie.
#pragma once
namespace at { class Tensor {}; }
