// Stub standing in for <torch/serialize/tensor.h> when the reference's
// pointnet2/src/*_gpu.cu launchers are compiled WITHOUT libtorch (oracle/_ref).
// The reference headers only need at::Tensor as a parameter type in prototypes
// of the pybind wrappers, which this build never defines or calls.
#pragma once
namespace at { class Tensor; }
