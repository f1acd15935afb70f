// Shared host helpers of the pokerrl_b200 CUDA library (error reporting for the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace prl {
// Stores `msg` for prl_last_error() and returns a non-zero status.
int fail(const char* msg);
// cudaSuccess -> 0; otherwise records "<where>: <cuda error string>" and returns the CUDA error code.
int check(cudaError_t e, const char* where);
// counts kernel launches issued by this library (prl_launch_count())
void count_launch();
}  // namespace prl
