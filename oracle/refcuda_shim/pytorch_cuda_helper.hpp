// Stand-in for mmcv's `pytorch_cuda_helper.hpp` (third_party/mmcv/mmcv/ops/csrc/common/
// pytorch_cuda_helper.hpp), which pulls in ATen / THC.  The deformable-attention kernel header
// only needs what `common_cuda_helper.hpp` (found next to it in the reference tree) defines,
// so this shim includes nothing else.  Test infrastructure: used only by oracle/Makefile to
// compile the REFERENCE's own kernels, where they lie, into oracle/_ref/.
#ifndef PAVENET_ORACLE_PYTORCH_CUDA_HELPER_SHIM
#define PAVENET_ORACLE_PYTORCH_CUDA_HELPER_SHIM
#include <cuda_runtime.h>
#include "common_cuda_helper.hpp"
#endif
