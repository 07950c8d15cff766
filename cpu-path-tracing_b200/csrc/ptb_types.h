// ptb_types.h -- fixed-width integers and the CUDA vector types for every translation unit that sees the device
// headers: nvcc / g++ builds take them from the standard and CUDA headers, the run-time compiler (NVRTC, ptb_jit.cpp)
// has the vector types built in and gets the integer names here (it has no <cstdint>).
#pragma once

#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#endif
