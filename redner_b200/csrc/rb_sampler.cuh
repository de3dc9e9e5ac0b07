// Per-pixel random streams, generated inside the consuming kernel (no sample buffers in HBM).
//   Sobol: scrambled Joe-Kuo sequence, index = sample_id, scramble = hash64shift((seed << 32) | pixel)
//          -- reference src/sobol_sampler.cpp:10-29 (scramble), :61-76 (sample()), :97-100 (begin_sample).
//   PCG32: one stream per pixel (inc = 2*(pixel+1)+1), state advanced by LCG skip-ahead so that sample s
//          can be generated independently -- reference src/pcg_sampler.cpp:8-50.
// Dimension allocation follows the ORDER of the reference's next_* calls (src/pathtracer.cpp:260,329,340):
//   main sampler:  camera (2) | per depth: light (4: light_sel, tri_sel, u, v) then bsdf (3: u, v, w)
// Samples are produced as double (the reference's Real) because discrete decisions (light / triangle / lobe
// selection) compare them against double CDFs; continuous uses cast to Real.
#pragma once
#include "rb_types.cuh"

#define RB_SOBOL_BITS 52

RB_HD unsigned long long rb_hash64shift(unsigned long long key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

RB_HD unsigned int rb_pcg32_next(unsigned long long& state, unsigned long long inc) {
    unsigned long long old = state;
    state = old * 6364136223846793005ULL + (inc | 1ULL);
    unsigned int xorshifted = (unsigned int)(((old >> 18u) ^ old) >> 27u);
    unsigned int rot = (unsigned int)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
// Advance an LCG by `delta` steps in O(log delta) (Brown, "Random number generation with arbitrary strides").
RB_HD unsigned long long rb_pcg32_advance(unsigned long long state, unsigned long long inc, unsigned long long delta) {
    unsigned long long cur_mult = 6364136223846793005ULL, cur_plus = inc | 1ULL;
    unsigned long long acc_mult = 1ULL, acc_plus = 0ULL;
    while (delta > 0) {
        if (delta & 1ULL) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1ULL) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    return acc_mult * state + acc_plus;
}

struct Sampler {
    int type;                      // rb_sampler_type
    unsigned long long scramble;   // sobol
    unsigned int index;            // sobol: sample id
    int dim;                       // sobol: current dimension
    const unsigned long long* mat; // sobol matrices (global or shared memory), row stride = mat_stride
    int mat_stride;
    unsigned long long pcg_state, pcg_inc;

    // `draws_before` = number of PCG draws this pixel's stream has consumed before this sample.
    RB_D void init(int type_, unsigned long long seed, int pixel, unsigned int sample_id, const unsigned long long* matrices,
                   int stride, unsigned long long draws_before) {
        type = type_;
        mat = matrices;
        mat_stride = stride;
        index = sample_id;
        dim = 0;
        if (type == RB_SAMPLER_SOBOL) {
            scramble = rb_hash64shift((seed << 32) | (unsigned long long)pixel);
        } else {
            pcg_inc = (((unsigned long long)pixel + 1ULL) << 1u) | 1ULL;
            unsigned long long st = 0ULL;
            rb_pcg32_next(st, pcg_inc);
            st += (0x853c49e6748fea9bULL + seed);
            rb_pcg32_next(st, pcg_inc);
            pcg_state = rb_pcg32_advance(st, pcg_inc, draws_before);
        }
    }
    RB_D double next() {
        if (type == RB_SAMPLER_SOBOL) {
            unsigned long long result = scramble & ((1ULL << RB_SOBOL_BITS) - 1ULL);
            const unsigned long long* m = mat + (size_t)dim * mat_stride;
            unsigned int idx = index;
            int i = 0;
            while (idx) {
                if (idx & 1u) result ^= m[i];
                idx >>= 1;
                i++;
            }
            dim++;
            return (double)result * (1.0 / (double)(1ULL << RB_SOBOL_BITS));
        } else {
            unsigned int r = rb_pcg32_next(pcg_state, pcg_inc);
            unsigned long long u = ((unsigned long long)r << 20) | 0x3ff0000000000000ULL;
            return __longlong_as_double((long long)u) - 1.0;
        }
    }
    RB_D void skip(int n) {
        if (type == RB_SAMPLER_SOBOL) {
            dim += n;
        } else {
            pcg_state = rb_pcg32_advance(pcg_state, pcg_inc, (unsigned long long)n);
        }
    }
};
