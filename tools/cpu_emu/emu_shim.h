// DEVELOPMENT AID ONLY -- host stand-ins for the handful of CUDA device intrinsics used by redner_b200/csrc/*.cuh so
// that the per-sample render logic can be compiled with g++ and stepped through / compared against the oracle in a
// container without a GPU.  Nothing under tools/cpu_emu is part of the product: redner_b200/ never loads it and the bench
// never runs it; tests/test_device_code_cpu.py checks this build of the device headers against the golden fixtures.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h> // vector types + empty __host__/__device__ when compiled by g++
#ifndef __CUDACC__
#define RB_CPU_EMU 1
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
#endif
