// Kernels of rb_render, included by rb_kernels.cu (general instantiation, global namespace) and by rb_kernels_lean.cu
// (feature-free instantiation, namespace rb_lean, RB_LEAN defined).  No include guard on purpose.
#define RB_BLOCK 128
#ifndef RB_MIN_BLOCKS_FWD
#define RB_MIN_BLOCKS_FWD 4
#endif
#ifndef RB_MIN_BLOCKS_TRACE
#define RB_MIN_BLOCKS_TRACE 4
#endif
#ifndef RB_MIN_BLOCKS_SEC
#define RB_MIN_BLOCKS_SEC 4
#endif
#ifndef RB_MIN_BLOCKS_SWEEP
#define RB_MIN_BLOCKS_SWEEP 4
#endif
#ifndef RB_BAND_BYTES
#define RB_BAND_BYTES (1ULL << 30) // scratch budget of one backward band (records + lists)
#endif
#ifndef RB_MIN_BLOCKS_BWD
#define RB_MIN_BLOCKS_BWD 4 // k_primary_edge
#endif

// j-th owned row -> viewport row, for the round-robin stripe partition
RB_D int owned_row_to_row(const RenderParams& rp, int j) {
    int s = j / rp.rows_per_stripe, w = j % rp.rows_per_stripe;
    return (s * rp.num_parts + rp.part) * rp.rows_per_stripe + w;
}
static int count_owned_rows(int H, int part, int num_parts, int rps) {
    int n = 0;
    for (int r = 0; r < H; r++)
        if ((r / rps) % num_parts == part) n++;
    return n;
}

struct WorkItem {
    bool valid;
    int pixel;    // viewport-relative pixel id (y * vp_w + x)
    int px, py;   // viewport-relative pixel coordinates
    int sample_lane;
};
RB_D WorkItem warp_work(const RenderParams& rp, int L, int owned_rows, long long group) {
    int lane = threadIdx.x & 31;
    int P = 32 / L;
    long long k = group * P + lane / L;
    WorkItem w;
    w.sample_lane = lane % L;
    long long n = (long long)owned_rows * rp.vp_w;
    w.valid = k < n;
    if (!w.valid) k = 0;
    int j = (int)(k / rp.vp_w), x = (int)(k % rp.vp_w);
    int y = owned_row_to_row(rp, j);
    w.pixel = y * rp.vp_w + x;
    w.px = x;
    w.py = y;
    return w;
}

// ------------------------------------------------------------------------------------------------ forward
#define RB_FWD_SYNC() RB_PHASE_SYNC() // measured: k_forward 5.2 -> 3.7 ms on C2 (one I-cache miss serves the block)
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_FWD) k_forward(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    const int L = ka.lanes_per_pixel;
    const int P = 32 / L;
    long long n_px = (long long)ka.owned_rows * rp.vp_w;
    long long groups = (n_px + P - 1) / P;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nb = (rp.spp + L - 1) / L;
    for (long long g0 = 0; g0 < groups; g0 += nwarps) { // block-uniform trip count (phase barrier inside)
        long long g = g0 + warp;
        WorkItem w = warp_work(rp, L, ka.owned_rows, g < groups ? g : 0);
        if (g >= groups) w.valid = false;
        V3 acc = zero3();
        for (int b = 0; b < nb; b++) {
            int s = b * L + w.sample_lane;
            RB_FWD_SYNC();
            if (w.valid && s < rp.spp) acc += forward_sample(sc, rp, w.pixel, w.px, w.py, s);
        }
        for (int off = L >> 1; off > 0; off >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
        }
        if (w.valid && w.sample_lane == 0) {
            float* px = ka.image + (size_t)rp.nd * w.pixel + rp.rad_dim;
            px[0] += (float)acc.x;
            px[1] += (float)acc.y;
            px[2] += (float)acc.z;
        }
    }
}

#ifndef RB_LEAN // only the general instantiation renders G-buffer channels
// G-buffer forward (any channel list): one warp per pixel group like k_forward, channels reduced with shuffles; id
// channels take the value of the highest-numbered sample that hit (the reference overwrites them sample after sample).
__global__ void __launch_bounds__(RB_BLOCK, 2) k_forward_channels(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    const int L = ka.lanes_per_pixel;
    const int P = 32 / L;
    long long n_px = (long long)ka.owned_rows * rp.vp_w;
    long long groups = (n_px + P - 1) / P;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nb = (rp.spp + L - 1) / L;
    const int nd = rp.nd < RB_MAX_ND ? rp.nd : RB_MAX_ND;
    for (long long g = warp; g < groups; g += nwarps) {
        WorkItem w = warp_work(rp, L, ka.owned_rows, g);
        float acc[RB_MAX_ND];
        for (int i = 0; i < nd; i++) acc[i] = 0.f;
        int ids[3] = {-1, -1, -1};
        int last = -1;
        for (int b = 0; b < nb; b++) {
            int s = b * L + w.sample_lane;
            if (w.valid && s < rp.spp) {
                int cur[3] = {-1, -1, -1};
                if (forward_sample_channels(sc, rp, w.pixel, w.px, w.py, s, acc, cur)) {
                    last = s;
                    ids[0] = cur[0]; ids[1] = cur[1]; ids[2] = cur[2];
                }
            }
        }
        for (int off = L >> 1; off > 0; off >>= 1) {
            for (int i = 0; i < nd; i++) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
            int ol = __shfl_xor_sync(0xffffffffu, last, off);
            int o0 = __shfl_xor_sync(0xffffffffu, ids[0], off), o1 = __shfl_xor_sync(0xffffffffu, ids[1], off), o2 = __shfl_xor_sync(0xffffffffu, ids[2], off);
            if (ol > last) { last = ol; ids[0] = o0; ids[1] = o1; ids[2] = o2; }
        }
        if (w.valid && w.sample_lane == 0) {
            float* px = ka.image + (size_t)rp.nd * w.pixel;
            int d = 0;
            for (int c = 0; c < rp.num_channels; c++) {
                int ch = rp.channels[c];
                int width = (ch == RB_CH_RADIANCE || ch == RB_CH_POSITION || ch == RB_CH_GEOMETRY_NORMAL || ch == RB_CH_SHADING_NORMAL ||
                             ch == RB_CH_DIFFUSE_REFLECTANCE || ch == RB_CH_SPECULAR_REFLECTANCE || ch == RB_CH_VERTEX_COLOR) ? 3
                          : (ch == RB_CH_UV || ch == RB_CH_BARYCENTRIC) ? 2 : (ch == RB_CH_GENERIC_TEXTURE ? rp.max_generic : 1);
                if (ch == RB_CH_SHAPE_ID || ch == RB_CH_TRIANGLE_ID || ch == RB_CH_MATERIAL_ID) {
                    int v = ids[ch - RB_CH_SHAPE_ID];
                    if (last >= 0 && d < nd) px[d] = (float)v;
                } else {
                    for (int i = 0; i < width && d + i < nd; i++) px[d + i] += acc[d + i];
                }
                d += width;
            }
        }
    }
}

#endif
// ------------------------------------------------------------------------------------------------ backward (interior + first hit)
RB_D void block_reduce_camera(float* cam_smem, double* cam_accum) {
    // cam_smem: [RB_CAM_ACC][blockDim.x]; reduce each row and add to the global double accumulators
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = warp; k < RB_CAM_ACC; k += nw) {
        float s = 0.f;
        for (int i = lane; i < (int)blockDim.x; i += 32) s += cam_smem[k * blockDim.x + i];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0 && s != 0.f) atomicAdd(&cam_accum[k], (double)s);
    }
}

// Dense sample index of the band -> (pixel, px, py, s).  Consecutive lanes are consecutive samples of a pixel.
struct SampleId {
    int pixel, px, py, s;
};
RB_D SampleId band_sample(const RenderParams& rp, long long I) {
    long long k = I / rp.spp;
    SampleId id;
    id.s = (int)(I - k * rp.spp);
    int j = (int)(k / rp.vp_w);
    id.px = (int)(k - (long long)j * rp.vp_w);
    id.py = owned_row_to_row(rp, j);
    id.pixel = id.py * rp.vp_w + id.px;
    return id;
}
// The work loops below are BLOCK-uniform (every thread of a block runs the same number of iterations, idle ones with
// act == false) because the per-sample stages contain phase barriers (RB_PHASE_SYNC, rb_render.cuh).
#define RB_BLOCK_LOOP(t, n) \
    for (long long t##_base = (long long)blockIdx.x * blockDim.x, t = t##_base + threadIdx.x; t##_base < (n); \
         t##_base += (long long)gridDim.x * blockDim.x, t = t##_base + threadIdx.x)
// Stage 1: replay the primal path of every sample of the band, one VertexRec per vertex.
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_TRACE) k_bwd_trace(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    RB_BLOCK_LOOP(t, ka.band_n) {
        bool act = t < ka.band_n;
        SampleId id = band_sample(rp, ka.band_i0 + (act ? t : 0));
        int n = bwd_trace(sc, rp, id.pixel, id.px, id.py, id.s, ka.records + (size_t)(act ? t : 0) * ka.rec_per_sample, 1, act);
        if (act) ka.nrec[t] = n;
    }
}
#ifndef RB_LEAN // the compaction does not depend on scene features
// (hit << 32 | vertices) of one sample: the scan input
struct CountOp {
    __host__ __device__ unsigned long long operator()(int nrec) const { return nrec < 0 ? 0ULL : ((1ULL << 32) | (unsigned long long)nrec); }
};
// Stage 2: deterministic compaction from the exclusive scan: samples that hit something, and their vertices.
__global__ void k_bwd_compact(const __grid_constant__ KernelArgs ka) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ka.band_n; t += (long long)gridDim.x * blockDim.x) {
        int n = ka.nrec[t];
        unsigned long long o = ka.offs[t];
        if (n >= 0) {
            ka.path_list[(unsigned)(o >> 32)] = (int)t;
            unsigned v = (unsigned)(o & 0xffffffffULL);
            for (int d = 0; d < n; d++) ka.vert_list[v + d] = (int)t * ka.rec_per_sample + d;
        }
        if (t == ka.band_n - 1) {
            unsigned long long tot = o + CountOp()(n);
            *ka.totals = tot;
            // statistics for the roofline accounting (mean executed bounces per sample)
            atomicAdd(&ka.ds.cam_accum[RB_CAM_ACC], (double)(tot & 0xffffffffULL));
            atomicAdd(&ka.ds.cam_accum[RB_CAM_ACC + 1], (double)(tot >> 32));
        }
    }
}
#endif
// Stage 3a: edge pick of every path vertex (secondary edge sampling); full warps of vertices.  Key = picked edge
// (num_edges = nothing picked), value = position in the vertex list.
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_SEC) k_bwd_sec_pick(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    const long long n = ka.n_verts;
    RB_BLOCK_LOOP(t, n) {
        RB_PHASE_SYNC();
        if (t < n) {
            int e = ka.vert_list[t];
            int ts = e / ka.rec_per_sample, d = e - ts * ka.rec_per_sample;
            SampleId id = band_sample(rp, ka.band_i0 + ts);
            VertexRec cur = ka.records[e];
            EdgePick pk;
            bool ok = bwd_secondary_pick(sc, ka, id.pixel, id.s, d, cur, pk);
            if (ok) ka.picks[t] = pk;
            ka.sec_keys[t] = ok ? (unsigned)pk.edge_id : (unsigned)sc.num_edges;
            ka.sec_vals[t] = (unsigned)t;
            ka.dpos[e] = zero3();
        }
    }
}
// Stage 3b: the two edge rays and their sub-paths, in edge order (neighbouring lanes aim at the same edge).
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_SEC) k_bwd_sec_shade(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    const long long n = ka.n_verts;
    RB_BLOCK_LOOP(j, n) {
        RB_PHASE_SYNC();
        if (j < n && ka.sec_keys_sorted[j] < (unsigned)sc.num_edges) {
            unsigned t = ka.sec_vals_sorted[j];
            int e = ka.vert_list[t];
            int ts = e / ka.rec_per_sample, d = e - ts * ka.rec_per_sample;
            SampleId id = band_sample(rp, ka.band_i0 + ts);
            VertexRec cur = ka.records[e];
            EdgePick pk = ka.picks[t];
            ka.dpos[e] = bwd_secondary_shade(sc, ka, id.pixel, id.s, d, cur, pk);
        }
    }
}
// Stage 4: reverse sweep of every path, first-hit and camera adjoints.
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_SWEEP) k_bwd_sweep(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    __shared__ float cam_smem[RB_CAM_ACC * RB_BLOCK];
    for (int k = 0; k < RB_CAM_ACC; k++) cam_smem[k * RB_BLOCK + threadIdx.x] = 0.f;
    CamAcc cam_acc;
    cam_acc.base = cam_smem + threadIdx.x;
    cam_acc.stride = RB_BLOCK;
    const RenderParams& rp = ka.rp;
    const long long n = ka.n_paths;
    RB_BLOCK_LOOP(t, n) {
        bool act = t < n;
        int ts = act ? ka.path_list[t] : 0;
        SampleId id = band_sample(rp, ka.band_i0 + ts);
        size_t base = (size_t)ts * ka.rec_per_sample;
        bwd_sweep(sc, ka, id.pixel, id.px, id.py, id.s, ka.records + base, 1, act ? ka.nrec[ts] : 0, ka.dpos ? ka.dpos + base : nullptr, cam_acc, act);
    }
    block_reduce_camera(cam_smem, ka.ds.cam_accum);
}

// ------------------------------------------------------------------------------------------------ primary edges
// One thread per (edge sample i, spp sample s), in two steps: k_prim_keys computes each sample's (edge, position on
// the edge) key, a radix sort orders the band by it, and k_primary_edge shades in that order -- neighbouring lanes then
// shoot nearly identical camera rays and scatter into the same two vertices.  (In sample order every lane picks an
// unrelated edge: 13 of 32 lanes active per instruction on C2.)  Sums are order-independent, so parity is unaffected.
#define RB_PRIM_SYNC() RB_PHASE_SYNC() // measured: k_primary_edge 19.2 -> 14.3 ms on C2
// dense index t of this device's primary-edge samples -> (i, s): i with i % num_parts == part
RB_D void prim_sample_id(const RenderParams& rp, long long t, long long& i, int& s) {
    long long k = t / rp.spp;
    i = k * rp.num_parts + rp.part;
    s = (int)(t - k * rp.spp);
}
__global__ void __launch_bounds__(256) k_prim_keys(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka, int dim_base, long long t0, int n,
                                                   unsigned* keys, unsigned* vals) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        long long i;
        int s;
        prim_sample_id(ka.rp, t0 + t, i, s);
        keys[t] = primary_edge_key(sc, ka.rp, i, s, dim_base);
        vals[t] = (unsigned)t;
    }
}
__global__ void __launch_bounds__(RB_BLOCK, RB_MIN_BLOCKS_BWD) k_primary_edge(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka, int dim_base,
                                                                               long long t0, int n, const unsigned* keys, const unsigned* vals) {
    __shared__ float cam_smem[RB_CAM_ACC * RB_BLOCK];
    for (int k = 0; k < RB_CAM_ACC; k++) cam_smem[k * RB_BLOCK + threadIdx.x] = 0.f;
    CamAcc cam_acc;
    cam_acc.base = cam_smem + threadIdx.x;
    cam_acc.stride = RB_BLOCK;
    RB_BLOCK_LOOP(t, n) {
        RB_PRIM_SYNC();
        if (t < n && keys[t] != 0xffffffffu) {
            long long i;
            int s;
            prim_sample_id(ka.rp, t0 + vals[t], i, s);
            primary_edge_sample(sc, ka, i, s, dim_base, cam_acc);
        }
    }
    block_reduce_camera(cam_smem, ka.ds.cam_accum);
}

#ifndef RB_LEAN
__global__ void k_finish_camera(DevCamera cam, const double* acc, rb_dcamera out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    finish_camera(cam, acc, out);
}
#endif

