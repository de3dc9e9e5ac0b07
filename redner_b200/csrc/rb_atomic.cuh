// Gradient scatter: warp-aggregated atomics.
// The reference issues one global atomic per scalar per thread (src/atomic.h:41-57; call sites e.g.
// src/primary_intersection.cpp:41-65, src/path_contribution.cpp:287-292).  Neighbouring paths of a warp mostly hit
// the same triangle / material / light, so we first combine the lanes that target the same address
// (__match_any_sync on the address) with a log-depth shuffle tree and let one lane per group issue the
// red.global.add.f32.  For a warp whose 32 lanes share the target this is 32x fewer atomics.
#pragma once
#include "rb_math.cuh"

#ifdef RB_CPU_EMU
// host-compiled debug emulator (tools/cpu_emu): single "lane", plain adds
RB_D void rb_red_add(float* addr, float v) { *addr += v; }
template <int N>
RB_D void warp_agg_add(float* addr, const float (&val)[N]) {
    for (int i = 0; i < N; i++) addr[i] += val[i];
}
#else
RB_D void rb_red_add(float* addr, float v) { atomicAdd(addr, v); }

// Combine `n` consecutive floats (n <= 9) across the lanes of the current convergence group that
// have the same `addr`, then the group leader adds them to addr[0..n).
template <int N>
RB_DFN void warp_agg_add(float* addr, const float (&val)[N]) {
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, (unsigned long long)addr);
    int lane = threadIdx.x & 31;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = val[i];
    int n = __popc(peers);
    if (n > 1) {
        if (peers == 0xffffffffu) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int i = 0; i < N; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
            }
        } else {
            int rank = __popc(peers & ((1u << lane) - 1u));
            for (int off = 1; off < n; off <<= 1) {
                int src_rank = rank + off;
                bool take = (src_rank < n) && ((rank & (2 * off - 1)) == 0);
                int src_lane = take ? (int)__fns(peers, 0, src_rank + 1) : lane;
#pragma unroll
                for (int i = 0; i < N; i++) {
                    float o = __shfl_sync(peers, v[i], src_lane);
                    if (take) v[i] += o;
                }
            }
        }
    }
    if (lane == __ffs(peers) - 1) {
#pragma unroll
        for (int i = 0; i < N; i++)
            if (v[i] != 0.f) rb_red_add(addr + i, v[i]);
    }
}
#endif // RB_CPU_EMU
RB_D void agg_add3(float* addr, V3 v) {
    float a[3] = {(float)v.x, (float)v.y, (float)v.z};
    warp_agg_add<3>(addr, a);
}
RB_D void agg_add2(float* addr, V2 v) {
    float a[2] = {(float)v.x, (float)v.y};
    warp_agg_add<2>(addr, a);
}
RB_D void agg_add1(float* addr, Real v) {
    float a[1] = {(float)v};
    warp_agg_add<1>(addr, a);
}
