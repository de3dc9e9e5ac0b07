// Entry points of the feature-free ("lean") instantiation of the render kernels (rb_kernels_lean.cu): scenes without an
// environment map, with a pinhole camera without lens model, rendered with channels == [radiance].  `sc` / `ka` point to
// the driver's DevScene / KernelArgs (the lean translation unit declares the same structs inside its own namespace).
#pragma once
#include <cuda_runtime.h>
namespace rb_lean_api {
enum Kernel { K_FORWARD = 0, K_BWD_TRACE, K_BWD_SEC_PICK, K_BWD_SEC_SHADE, K_BWD_SWEEP, K_PRIM_KEYS, K_PRIMARY_EDGE };
int grid(Kernel k, int device); // SMs x resident blocks per SM
void forward(const void* sc, const void* ka, int grid, cudaStream_t stream);
void bwd_trace(const void* sc, const void* ka, int grid, cudaStream_t stream);
void bwd_sec_pick(const void* sc, const void* ka, int grid, cudaStream_t stream);
void bwd_sec_pick_hier(const void* sc, const void* ka, int grid, cudaStream_t stream);
void bwd_sec_shade(const void* sc, const void* ka, int grid, cudaStream_t stream);
void bwd_sweep(const void* sc, const void* ka, int grid, cudaStream_t stream);
void prim_keys(const void* sc, const void* ka, int dim_base, long long t0, int n, unsigned* keys, unsigned* vals, int grid, cudaStream_t stream);
void primary_edge(const void* sc, const void* ka, int dim_base, long long t0, int n, const unsigned* keys, const unsigned* vals, int grid, cudaStream_t stream);
} // namespace rb_lean_api
