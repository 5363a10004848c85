// See stubs/torch/serialize/tensor.h.
#pragma once
#include <cuda_runtime_api.h>
