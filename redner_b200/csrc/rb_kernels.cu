// Render kernels and the rb_render driver (reference: render(), src/pathtracer.cpp:177-958).
//
// Execution model.  The reference runs a host-driven wavefront: ~20 launches and 2 host syncs per
// bounce per sample, with every per-path field (5.5 - 9.4 KB/pixel in double) streamed through managed memory between
// stage functors.  Here a path lives in registers inside a kernel and only a 128-byte record per path vertex crosses kernels:
//   forward    k_forward / k_forward_channels   camera sample -> primary hit -> emission -> bounce loop -> pixel
//   backward   per band of samples:  k_bwd_trace (primal replay, records, work lists by warp ballot) -> k_bwd_sec_pick ->
//              counting sort by edge (k_sec_offsets, k_sec_scatter) -> k_bwd_sec_shade (boundary terms) -> k_bwd_sweep (reverse
//              sweep, first-hit and camera adjoints);  then k_prim_keys -> radix sort -> k_primary_edge;  k_finish_camera
// Every kernel is small enough for the GPC instruction cache and walks the stages of a sample block-synchronously
// (RB_PHASE_SYNC): a fused megakernel of the same code ran instruction-fetch bound at 6 % issue utilisation.
// Forward: a warp owns 32/L pixels with L lanes per pixel (L = min(32, 2^floor(log2 spp))); lanes of a pixel are its
// samples, so rays of a warp are coherent in the BVH, the pixel is reduced with shuffles and written by one lane
// without atomics (deterministic image).  Gradient atomics are aggregated per warp before they reach L2.
// The per-sample logic itself lives in rb_render.cuh.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rb_lean_api.h"
#include "rb_render.cuh"
#include "rb_scene.cuh"

#include "rb_kernels_body.cuh"

// ------------------------------------------------------------------------------------------------ driver
// Backward scratch (gradient descriptors, band records and work lists): ONE grow-only allocation per device, kept for
// the life of the process so that a training loop pays no cudaMalloc / pool growth per step (measured: ~20 ms per step
// with per-call cudaMallocAsync of the 1 GiB band).  rb_render holds the lock for the whole backward pass, which also
// serialises concurrent backward passes on one device; rb_release_scratch() frees everything.
#include <mutex>
struct DeviceScratch {
    std::mutex mutex; // one backward pass per DEVICE at a time; passes on different devices of one process run concurrently
    char* ptr = nullptr;
    size_t bytes = 0;
};
static DeviceScratch g_scratch[64];
static char* scratch_ensure(DeviceScratch& sc, size_t bytes) { // (caller holds sc.mutex and has made the device current)
    if (sc.bytes >= bytes) return sc.ptr;
    if (sc.ptr) {
        cudaDeviceSynchronize();
        cudaFree(sc.ptr);
        sc.ptr = nullptr;
        sc.bytes = 0;
    }
    if (cudaMalloc((void**)&sc.ptr, bytes) != cudaSuccess) {
        sc.ptr = nullptr;
        return nullptr;
    }
    sc.bytes = bytes;
    return sc.ptr;
}
extern "C" void rb_release_scratch(void) {
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; d++) {
        std::lock_guard<std::mutex> lock(g_scratch[d].mutex);
        if (g_scratch[d].ptr) {
            cudaSetDevice(d);
            cudaDeviceSynchronize();
            cudaFree(g_scratch[d].ptr);
            g_scratch[d].ptr = nullptr;
            g_scratch[d].bytes = 0;
        }
    }
    cudaSetDevice(prev);
}
// Restores the caller's current device on EVERY exit path of rb_render (RB_CUDA_OK returns).
struct RenderGuard {
    int prev_device = -1;
    ~RenderGuard() {
        if (prev_device >= 0) cudaSetDevice(prev_device);
    }
};
// Persistent grid of a kernel: SMs x resident blocks per SM for its block size and dynamic shared memory.
static int pick_grid(const void* kernel, int device, int block = RB_BLOCK, size_t smem = 0) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

extern "C" int rb_render(const rb_scene* scene_, const rb_options* opt, float* image, const float* d_image, const rb_dscene_desc* d_scene,
                         float* screen_grad, void* stream_) {
    rb_scene* scene = const_cast<rb_scene*>(scene_);
    if (!scene || !opt) {
        rb_set_error("rb_render: null scene / options");
        return 1;
    }
    if (image == nullptr && d_image == nullptr) {
        rb_set_error("rb_render: neither rendered_image nor d_rendered_image given");
        return 1;
    }
    if (d_image != nullptr && d_scene == nullptr) {
        rb_set_error("rb_render: d_rendered_image given without d_scene");
        return 1;
    }
    if (opt->max_bounces < 0 || opt->num_samples < 0) {
        rb_set_error("rb_render: negative max_bounces / num_samples");
        return 1;
    }
    KernelArgs ka;
    memset(&ka, 0, sizeof(ka));
    RenderParams& rp = ka.rp;
    rp.seed = opt->seed;
    rp.spp = opt->num_samples;
    rp.max_bounces = opt->max_bounces;
    rp.sampler_type = opt->sampler_type;
    rp.sample_pixel_center = opt->sample_pixel_center;
    rp.num_channels = opt->num_channels;
    rp.max_generic = scene->max_generic_texture_dimension;
    rp.rad_dim = -1;
    if (opt->num_channels > RB_CH_COUNT) {
        rb_set_error("rb_render: too many channels");
        return 1;
    }
    bool only_radiance = opt->num_channels == 1 && opt->channels[0] == RB_CH_RADIANCE;
    for (int i = 0; i < opt->num_channels; i++) {
        rp.channels[i] = opt->channels[i];
        if (opt->channels[i] < 0 || opt->channels[i] >= RB_CH_COUNT) {
            rb_set_error("rb_render: unknown channel");
            return 1;
        }
        if (opt->channels[i] == RB_CH_RADIANCE) {
            if (rp.rad_dim != -1) {
                rb_set_error("Duplicated radiance channel"); // src/channels.cpp:24-26
                return 1;
            }
            // the reference stores the CHANNEL INDEX and uses it as a float offset (src/channels.cpp:27,
            // src/path_contribution.cpp:125-129); identical whenever radiance is the first channel
            rp.rad_dim = i;
        }
    }
    rp.only_radiance = only_radiance ? 1 : 0;
    // The feature-free instantiation of the kernels (rb_kernels_lean.cu) serves the common configuration.
    const bool lean_allowed = getenv("RB_NO_LEAN") == nullptr; // (test hook: force the general kernels)
    const bool lean = lean_allowed && only_radiance && !scene->dev.has_envmap && scene->dev.cam.type == RB_CAMERA_PERSPECTIVE && !scene->dev.cam.has_distortion;
    namespace la = rb_lean_api;
    rp.nd = rb_compute_num_channels(opt->channels, opt->num_channels, rp.max_generic);
    rp.rad_off = -1;
    for (int i = 0, off = 0; i < opt->num_channels; i++) {
        if (opt->channels[i] == RB_CH_RADIANCE) rp.rad_off = off;
        off += rb_channel_width(opt->channels[i], rp.max_generic);
    }
    if (rp.nd > RB_MAX_ND) {
        rb_set_error("rb_render: more than 64 image dimensions requested");
        return 1;
    }
    rp.part = scene->part;
    rp.num_parts = scene->num_parts;
    rp.rows_per_stripe = scene->rows_per_stripe;
    rp.vp_w = scene->cam.viewport_end[0] - scene->cam.viewport_beg[0];
    rp.vp_h = scene->cam.viewport_end[1] - scene->cam.viewport_beg[1];
    if (rp.vp_w <= 0 || rp.vp_h <= 0 || rp.spp == 0) return 0;
    int L = 1;
    while (L * 2 <= std::min(32, rp.spp)) L *= 2;
    ka.lanes_per_pixel = L;
    ka.owned_rows = count_owned_rows(rp.vp_h, rp.part, rp.num_parts, rp.rows_per_stripe);
    ka.image = image;
    ka.d_image = d_image;
    ka.screen_grad = screen_grad;

    RenderGuard guard;
    int prev = 0;
    RB_CUDA_OK(cudaGetDevice(&prev));
    RB_CUDA_OK(cudaSetDevice(scene->device));
    guard.prev_device = prev;
    cudaStream_t stream = (cudaStream_t)stream_;
    // per-kernel CUDA events on the render stream: [0] start, [1] after k_forward, [2] after the backward bands,
    // [3] after k_primary_edge, [4] after k_finish_camera; then 4 per backward band.  The events belong to the scene and are
    // reused by every call (creating ~40 events per call cost more than the kernels of a small render).
    EventPool& events = scene->events;
    if (!events.ensure(5)) {
        rb_set_error("rb_render: cudaEventCreate failed");
        return 1;
    }
std::vector<cudaEvent_t>& ev = events.ev;
    int launches = 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, scene->device);
    std::vector<BandCounters> host_counters;
    long long num_bands_done = 0;
    std::unique_lock<std::mutex> scratch_lock; // per-device scratch, held for the backward pass (released before the guard runs)
    RB_CUDA_OK(cudaEventRecord(ev[0], stream));
    if (image != nullptr) {
        if (only_radiance) {
            if (lean) {
                la::forward(&scene->dev, &ka, la::grid(la::K_FORWARD, scene->device), stream);
            } else {
                int grid = pick_grid((const void*)k_forward, scene->device, RB_BLOCK_FWD);
                k_forward<<<grid, RB_BLOCK_FWD, 0, stream>>>(scene->dev, ka);
            }
        } else {
            int grid = pick_grid((const void*)k_forward_channels, scene->device);
            k_forward_channels<<<grid, RB_BLOCK, 0, stream>>>(scene->dev, ka);
        }
        launches++;
    }
    RB_CUDA_OK(cudaEventRecord(ev[1], stream));
    if (d_image == nullptr) {
        for (int i = 2; i < 5; i++) RB_CUDA_OK(cudaEventRecord(ev[i], stream));
    }
    if (d_image != nullptr) {
        // device copies of the gradient descriptor
        if (d_scene->num_shapes != (int)scene->shapes.size() || d_scene->num_materials != (int)scene->materials.size() ||
            d_scene->num_lights != (int)scene->lights.size()) {
            rb_set_error("rb_render: d_scene does not match the scene (shape / material / light counts)");
            return 1;
        }
        // ---- scratch layout: gradient descriptors | camera accumulators | band (records, boundary terms, work lists)
        const bool secondary = scene->dev.use_secondary_edge && scene->dev.num_edges > 0 && scene->dev.num_lights > 0 && rp.rad_dim >= 0;
        const long long total_samples = (long long)ka.owned_rows * rp.vp_w * rp.spp;
        ka.rec_per_sample = rp.max_bounces + 1;
        if (secondary && rp.max_bounces > 64) {
            rb_set_error("rb_render: secondary edge sampling supports at most 64 bounces");
            return 1;
        }
        const size_t per_sample = (size_t)ka.rec_per_sample * (sizeof(VertexRec) + (secondary ? sizeof(V3) + sizeof(EdgePick) + 16 : 0)) + 2 * sizeof(int) +
                                  sizeof(ListCount) + (secondary ? sizeof(ulonglong2) : 0);
        size_t band_bytes = RB_BAND_BYTES;
        if (const char* env = getenv("RB_BAND_BYTES")) { // test hook: force many small bands
            long long v = atoll(env);
            if (v > 0) band_bytes = (size_t)v;
        }
        long long band = (long long)std::max<size_t>(band_bytes / per_sample, 1024);
        band = std::min<long long>(band, (1LL << 30) / ka.rec_per_sample);
        band = std::min<long long>(band, std::max<long long>(total_samples, 1));
        const long long num_bands = (total_samples + band - 1) / band;
        if (!events.ensure((size_t)(5 + 4 * num_bands))) {
            rb_set_error("rb_render: cudaEventCreate failed");
            return 1;
        }
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t nb_shapes = std::max(1, d_scene->num_shapes) * sizeof(rb_dshape), nb_mats = std::max(1, d_scene->num_materials) * sizeof(rb_material),
               nb_lights = std::max(1, d_scene->num_lights) * sizeof(float*);
        size_t o_shapes = 0, o_mats = o_shapes + al(nb_shapes), o_lights = o_mats + al(nb_mats), o_cam = o_lights + al(nb_lights);
        size_t o_cnt = o_cam + al(RB_CAM_ACC * sizeof(double));
        const size_t n_edges = (size_t)std::max(scene->dev.num_edges, 1);
        size_t o_hist = o_cnt + al((size_t)num_bands * sizeof(BandCounters)), o_eoffs = o_hist + al(secondary ? n_edges * 4 : 0);
        size_t o_ecur = o_eoffs + al(secondary ? n_edges * 4 : 0);
        size_t o_rec = o_ecur + al(secondary ? n_edges * 4 : 0), o_dpos = o_rec + al((size_t)band * ka.rec_per_sample * sizeof(VertexRec));
        size_t o_nrec = o_dpos + al(secondary ? (size_t)band * ka.rec_per_sample * sizeof(V3) : 0);
        size_t o_vmask = o_nrec + al((size_t)band * sizeof(int)), o_offs = o_vmask + al(secondary ? (size_t)band * sizeof(ulonglong2) : 0);
        auto count_it = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), ListCountOf{nullptr, nullptr});
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveScan(nullptr, scan_bytes, count_it, (ListCount*)nullptr, ListCountSum(), ListCount{0, 0, 0, 0}, (int)band, stream);
        size_t o_scan = o_offs + al((size_t)band * sizeof(ListCount)), o_paths = o_scan + al(scan_bytes);
        // boundary terms: vertex list, edge picks, (edge, vertex) per slot, slots in edge order
        const size_t vert_cap = (size_t)band * ka.rec_per_sample, slots = vert_cap + 32;
        size_t o_verts = o_paths + al((size_t)band * sizeof(int)), o_picks = o_verts + al(secondary ? vert_cap * sizeof(int) : 0);
        size_t o_sk = o_picks + al(secondary ? slots * sizeof(EdgePick) : 0), o_sv = o_sk + al(secondary ? slots * 4 : 0), o_so = o_sv + al(secondary ? slots * 4 : 0);
        size_t scratch_bytes = o_so + al(secondary ? slots * 4 : 0);
        // primary-edge pass (reuses the band area): keys/values double buffers + radix-sort temporaries
        const bool primary = scene->dev.use_primary_edge && scene->dev.num_edges > 0 && scene->dev.prim_edge_cdf != nullptr;
        const long long n_px_all = (long long)rp.vp_w * rp.vp_h;
        const long long total_e = primary ? ((n_px_all - rp.part + rp.num_parts - 1) / rp.num_parts) * rp.spp : 0; // i % num_parts == part
        const long long band_e = std::min<long long>(std::max<long long>(total_e, 1), 1LL << 26);
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr, (int)band_e, 0, 32, stream);
        size_t o_k0 = o_rec, o_k1 = o_k0 + al((size_t)band_e * 4), o_v0 = o_k1 + al((size_t)band_e * 4), o_v1 = o_v0 + al((size_t)band_e * 4);
        size_t o_sort = o_v1 + al((size_t)band_e * 4);
        if (primary) scratch_bytes = std::max(scratch_bytes, o_sort + al(sort_bytes));
        DeviceScratch& dscratch = g_scratch[scene->device & 63];
        scratch_lock = std::unique_lock<std::mutex>(dscratch.mutex);
        char* scratch = scratch_ensure(dscratch, scratch_bytes);
        if (!scratch) {
            rb_set_error("rb_render: out of device memory for the backward scratch");
            return 1;
        }
        rb_dshape* d_shapes = (rb_dshape*)(scratch + o_shapes);
        rb_material* d_mats = (rb_material*)(scratch + o_mats);
        float** d_lights = (float**)(scratch + o_lights);
        double* cam_accum = (double*)(scratch + o_cam);
        if (d_scene->num_shapes) RB_CUDA_OK(cudaMemcpyAsync(d_shapes, d_scene->shapes, d_scene->num_shapes * sizeof(rb_dshape), cudaMemcpyHostToDevice, stream));
        if (d_scene->num_materials)
            RB_CUDA_OK(cudaMemcpyAsync(d_mats, d_scene->materials, d_scene->num_materials * sizeof(rb_material), cudaMemcpyHostToDevice, stream));
        if (d_scene->num_lights)
            RB_CUDA_OK(cudaMemcpyAsync(d_lights, d_scene->light_intensity, d_scene->num_lights * sizeof(float*), cudaMemcpyHostToDevice, stream));
        // camera accumulators, per-band counters and the edge histogram are contiguous: one memset
        RB_CUDA_OK(cudaMemsetAsync(cam_accum, 0, o_eoffs - o_cam, stream));
        ka.ds.shapes = d_shapes;
        ka.ds.materials = d_mats;
        ka.ds.light_intensity = d_lights;
        ka.ds.cam_accum = cam_accum;
        memset(&ka.ds.env_values, 0, sizeof(rb_texture));
        ka.ds.env_w2e = nullptr;
        if (d_scene->envmap != nullptr) {
            ka.ds.env_values = d_scene->envmap->values;
            ka.ds.env_w2e = d_scene->envmap->world_to_env;
        } else if (scene->dev.has_envmap) {
            rb_set_error("rb_render: the scene has an environment map but d_scene has no envmap gradient buffers");
            return 1;
        }

        // ---- interior + first-hit adjoints, band by band: trace (+ work lists) -> boundary terms (pick, counting sort by edge, shade)
        //      -> sweep.  Every list size stays on the device: fixed persistent grids, no host synchronisation inside the pass.
        BandCounters* counters = (BandCounters*)(scratch + o_cnt);
        ka.records = (VertexRec*)(scratch + o_rec);
        ka.dpos = secondary ? (V3*)(scratch + o_dpos) : nullptr;
        ka.nrec = (int*)(scratch + o_nrec);
        ka.vmask = (ulonglong2*)(scratch + o_vmask);
        ListCount* list_offs = (ListCount*)(scratch + o_offs);
        ka.path_list = (int*)(scratch + o_paths);
        ka.vert_list = (int*)(scratch + o_verts);
        ka.vert_cap = (int)vert_cap;
        ka.picks = (EdgePick*)(scratch + o_picks);
        ka.sec_keys = (unsigned*)(scratch + o_sk);
        ka.sec_vals = (unsigned*)(scratch + o_sv);
        ka.sec_order = (unsigned*)(scratch + o_so);
        ka.edge_hist = (unsigned*)(scratch + o_hist);
        ka.edge_offs = (unsigned*)(scratch + o_eoffs);
        ka.edge_cursor = (unsigned*)(scratch + o_ecur);
        // Persistent-warp hierarchy pick (k_bwd_sec_pick_hier: resumable walks, finished lanes refilled from a work counter): identical
        // picks, but measured SLOWER than the plain kernel (boundary stage C2 4.5 -> 5.3 ms, teapot 44.2 -> 47.6 ms,
        // profiles/r02_persistent_pick_ab.txt): the stage waits on dependent loads, not on issue slots, so idle lanes cost nothing
        // and the refill logic is pure overhead.  Opt-in for measurements.
        ka.hier_persistent = getenv("RB_PERSISTENT_PICK") != nullptr ? 1 : 0;
        int grid_t, grid_p, grid_s, grid_w;
        if (lean) {
            grid_t = la::grid(la::K_BWD_TRACE, scene->device);
            grid_p = la::grid(la::K_BWD_SEC_PICK, scene->device);
            grid_s = la::grid(la::K_BWD_SEC_SHADE, scene->device);
            grid_w = la::grid(la::K_BWD_SWEEP, scene->device);
        } else {
            grid_t = pick_grid((const void*)k_bwd_trace, scene->device, RB_BLOCK_TRACE);
            grid_p = pick_grid((const void*)k_bwd_sec_pick, scene->device, RB_BLOCK_SEC);
            grid_s = pick_grid((const void*)k_bwd_sec_shade, scene->device, RB_BLOCK_SEC);
            grid_w = pick_grid((const void*)k_bwd_sweep, scene->device, RB_BLOCK_SWEEP, RB_SMEM_CAM(RB_BLOCK_SWEEP));
        }
        long long band_idx = 0;
        for (long long i0 = 0; i0 < total_samples; i0 += band, band_idx++) {
            ka.band_i0 = i0;
            ka.band_n = (int)std::min<long long>(band, total_samples - i0);
            ka.counters = counters + band_idx;
            cudaEvent_t* e4 = &events.ev[(size_t)(5 + 4 * band_idx)];
            RB_CUDA_OK(cudaEventRecord(e4[0], stream));
            if (lean) la::bwd_trace(&scene->dev, &ka, grid_t, stream);
            else k_bwd_trace<<<grid_t, RB_BLOCK_TRACE, 0, stream>>>(scene->dev, ka);
            RB_CUDA_OK(cudaEventRecord(e4[1], stream));
            {
                auto counts = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), ListCountOf{ka.nrec, secondary ? ka.vmask : nullptr});
                RB_CUDA_OK(cub::DeviceScan::ExclusiveScan(scratch + o_scan, scan_bytes, counts, list_offs, ListCountSum(), ListCount{0, 0, 0, 0}, ka.band_n, stream));
                k_bwd_compact<<<std::min((ka.band_n + 255) / 256, sms * 8), 256, 0, stream>>>(ka, list_offs);
            }
            launches += 4;
            if (secondary) {
                if (lean) la::bwd_sec_pick(&scene->dev, &ka, grid_p, stream);
                else k_bwd_sec_pick<<<grid_p, RB_BLOCK_SEC, 0, stream>>>(scene->dev, ka);
                if (ka.hier_persistent) {
                    if (lean) la::bwd_sec_pick_hier(&scene->dev, &ka, grid_p, stream);
                    else k_bwd_sec_pick_hier<<<grid_p, RB_BLOCK_SEC, 0, stream>>>(scene->dev, ka);
                    launches++;
                }
                k_sec_offsets<<<1, 1024, 0, stream>>>(ka, scene->dev.num_edges);
                k_sec_scatter<<<sms * 8, 256, 0, stream>>>(ka);
                if (lean) la::bwd_sec_shade(&scene->dev, &ka, grid_s, stream);
                else k_bwd_sec_shade<<<grid_s, RB_BLOCK_SEC, 0, stream>>>(scene->dev, ka);
                launches += 4;
            }
            RB_CUDA_OK(cudaEventRecord(e4[2], stream));
            if (lean) la::bwd_sweep(&scene->dev, &ka, grid_w, stream);
            else k_bwd_sweep<<<grid_w, RB_BLOCK_SWEEP, RB_SMEM_CAM(RB_BLOCK_SWEEP), stream>>>(scene->dev, ka);
            launches++;
            RB_CUDA_OK(cudaEventRecord(e4[3], stream));
        }
        num_bands_done = band_idx;
        if (num_bands_done > 0) {
            host_counters.resize((size_t)num_bands_done);
            RB_CUDA_OK(cudaMemcpyAsync(host_counters.data(), counters, (size_t)num_bands_done * sizeof(BandCounters), cudaMemcpyDeviceToHost, stream));
        }
        RB_CUDA_OK(cudaEventRecord(ev[2], stream));
        if (scene->dev.use_primary_edge && scene->dev.num_edges > 0 && scene->dev.prim_edge_cdf != nullptr) {
            int grid_e = lean ? la::grid(la::K_PRIMARY_EDGE, scene->device) : pick_grid((const void*)k_primary_edge, scene->device, RB_BLOCK_PRIM, RB_SMEM_CAM(RB_BLOCK_PRIM));
            int dim_base = primary_edge_dim_base(scene->dev, rp);
            unsigned *k0 = (unsigned*)(scratch + o_k0), *k1 = (unsigned*)(scratch + o_k1), *v0 = (unsigned*)(scratch + o_v0), *v1 = (unsigned*)(scratch + o_v1);
            int ebits = 1;
            while ((1 << ebits) < scene->dev.num_edges && ebits < 31) ebits++;
            for (long long t0 = 0; t0 < total_e; t0 += band_e) {
                int n = (int)std::min<long long>(band_e, total_e - t0);
                int grid_k = std::min((n + 255) / 256, 148 * 16);
                if (lean) la::prim_keys(&scene->dev, &ka, dim_base, t0, n, k0, v0, grid_k, stream);
                else k_prim_keys<<<grid_k, 256, 0, stream>>>(scene->dev, ka, dim_base, t0, n, k0, v0);
                // sort on the edge bits and the top 8 bits of the position only (coarser order is enough for coherence)
                int lo = std::max(0, 31 - ebits - 8);
                cub::DeviceRadixSort::SortPairs(scratch + o_sort, sort_bytes, k0, k1, v0, v1, n, lo, 32, stream);
                if (lean) la::primary_edge(&scene->dev, &ka, dim_base, t0, n, k1, v1, grid_e, stream);
                else k_primary_edge<<<grid_e, RB_BLOCK_PRIM, RB_SMEM_CAM(RB_BLOCK_PRIM), stream>>>(scene->dev, ka, dim_base, t0, n, k1, v1);
                launches += 2 + 4;
            }
        }
        RB_CUDA_OK(cudaEventRecord(ev[3], stream));
        k_finish_camera<<<1, 32, 0, stream>>>(scene->dev.cam, cam_accum, d_scene->camera);
        launches++;
        RB_CUDA_OK(cudaEventRecord(ev[4], stream));
    }
    cudaError_t err = cudaStreamSynchronize(stream);
    if (err == cudaSuccess) err = cudaGetLastError();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[4]);
    scene->last_launches = launches;
    scene->last_kernel_ms = ms;
    for (int i = 0; i < 4; i++) {
        scene->last_stage_ms[i] = 0.f;
        if (err == cudaSuccess) cudaEventElapsedTime(&scene->last_stage_ms[i], ev[i], ev[i + 1]);
    }
    for (int i = 0; i < 3; i++) scene->last_bwd_ms[i] = 0.f;
    double host_stats[2] = {0, 0};
    if (err == cudaSuccess)
        for (long long b = 0; b < num_bands_done; b++) {
            for (int i = 0; i < 3; i++) {
                float t = 0.f;
                cudaEventElapsedTime(&t, events.ev[5 + 4 * b + i], events.ev[5 + 4 * b + i + 1]);
                scene->last_bwd_ms[i] += t;
            }
            host_stats[0] += (double)host_counters[(size_t)b].total_vertices;
            host_stats[1] += (double)host_counters[(size_t)b].total_hits;
        }
    scene->last_path_vertices = host_stats[0];
    scene->last_primary_hits = host_stats[1];
    if (err != cudaSuccess) {
        rb_set_error(std::string("rb_render: kernel failure: ") + cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

// Batch of views of ONE scene (the multi-view loops of pyredner/render_utils.py:407-430 and of BASELINE config 5 rebuild the whole
// Scene per view): geometry, BVH, light tables and the edge list are shared; per view only the camera-dependent tables are rebuilt,
// on the device (rb_scene_set_camera), followed by the usual kernel set.  Gradients of all views ACCUMULATE into the buffers of
// d_scenes[k] (pass the same descriptor for every view to sum them: one gradient buffer for the batch).
extern "C" int rb_render_batch(rb_scene* scene, int num_views, const rb_camera* cameras, const rb_options* options, float* const* images,
                               const float* const* d_images, const rb_dscene_desc* const* d_scenes, void* stream) {
    if (!scene || num_views < 0 || (num_views > 0 && (!cameras || !options))) {
        rb_set_error("rb_render_batch: null argument");
        return 1;
    }
    for (int k = 0; k < num_views; k++) {
        if (rb_scene_set_camera(scene, &cameras[k])) return 1;
        if (rb_render(scene, &options[k], images ? images[k] : nullptr, d_images ? d_images[k] : nullptr, d_scenes ? d_scenes[k] : nullptr, nullptr, stream)) return 1;
    }
    return 0;
}
