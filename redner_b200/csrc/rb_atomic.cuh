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
RB_D void warp_agg_add3(float* addr, float x, float y, float z) { addr[0] += x; addr[1] += y; addr[2] += z; }
#else
RB_D void rb_red_add(float* addr, float v) { atomicAdd(addr, v); } // result unused -> RED.E.ADD.F32 (fire and forget)
// (A per-block shared-memory cache in front of the REDs was tried and measured 1.4x - 4.6x SLOWER on the backward kernel:
//  the 64-bit CAS + shared atomic per scalar costs more than a fire-and-forget L2 reduction; see DESIGN.md.)

// Combine up to three consecutive floats across the lanes of the current convergence group that target the same `addr`;
// the group leader issues the reductions.  Must stay INLINE: inside a real function the callers' lanes are not
// reconverged, __activemask() degenerates to single lanes and every lane issues its own atomics (measured: 7x slower).
// The peers of one address form a linked list in lane order; pointer doubling turns every lane's value into the
// suffix sum of its list in ceil(log2 n) rounds, so the lowest lane ends with the group total.  One code path for
// every group shape and no __fns(): the previous rank-based tree made this helper ~45% of the (then fused) backward kernel's 1.6 MB of
// SASS, which then ran instruction-fetch bound (profiles/r01_ncu_fused_k_backward_summary.txt).
RB_D void warp_agg_add3(float* addr, float x, float y, float z) {
    unsigned active = __activemask();
    unsigned peers = __match_any_sync(active, (unsigned long long)addr);
    int lane = threadIdx.x & 31;
    unsigned above = peers & (0xfffffffeu << lane);
    int nxt = above ? __ffs(above) - 1 : -1;
    int n = __popc(peers);
#pragma unroll 1
    for (int span = 1; span < n; span <<= 1) {
        int src = nxt >= 0 ? nxt : lane;
        float ox = __shfl_sync(peers, x, src), oy = __shfl_sync(peers, y, src), oz = __shfl_sync(peers, z, src);
        int nn = __shfl_sync(peers, nxt, src);
        if (nxt >= 0) {
            x += ox;
            y += oy;
            z += oz;
            nxt = nn;
        }
    }
    if (lane == __ffs(peers) - 1) {
        if (x != 0.f) rb_red_add(addr, x);
        if (y != 0.f) rb_red_add(addr + 1, y);
        if (z != 0.f) rb_red_add(addr + 2, z);
    }
}
#endif // RB_CPU_EMU
// NOTE: a zero component is never written, so the 1- and 2-component variants may pass 0 for the unused lanes of the
// triple without touching memory beyond their target.
RB_D void agg_add3(float* addr, V3 v) { warp_agg_add3(addr, (float)v.x, (float)v.y, (float)v.z); }
RB_D void agg_add2(float* addr, V2 v) { warp_agg_add3(addr, (float)v.x, (float)v.y, 0.f); }
RB_D void agg_add1(float* addr, Real v) { warp_agg_add3(addr, (float)v, 0.f, 0.f); }
