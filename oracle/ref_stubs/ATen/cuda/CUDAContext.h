// Stub for <ATen/cuda/CUDAContext.h>; see ../../torch/serialize/tensor.h.
#pragma once
#include <cuda_runtime_api.h>
