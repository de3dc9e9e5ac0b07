// Render kernels and the rb_render driver (reference: render(), src/pathtracer.cpp:177-958).
//
// Execution model.  The reference runs a host-driven wavefront: ~20 launches and 2 host syncs per
// bounce per sample, with every per-path field (5.5 - 9.4 KB/pixel in double) streamed through managed memory between
// stage functors.  Here a path lives in registers inside a kernel and only a 128-byte record per path vertex crosses kernels:
//   forward    k_forward / k_forward_channels   camera sample -> primary hit -> emission -> bounce loop -> pixel
//   backward   per band of samples:  k_bwd_trace (primal replay, records) -> scan + k_bwd_compact (path / vertex lists)
//              -> k_bwd_sec_pick -> radix sort by edge -> k_bwd_sec_shade (boundary terms) -> k_bwd_sweep (reverse sweep,
//              first-hit and camera adjoints);  then k_prim_keys -> radix sort -> k_primary_edge;  k_finish_camera
// Every kernel is small enough for the GPC instruction cache and walks the stages of a sample block-synchronously
// (RB_PHASE_SYNC): a fused megakernel of the same code ran instruction-fetch bound at 6 % issue utilisation.
// Forward: a warp owns 32/L pixels with L lanes per pixel (L = min(32, 2^floor(log2 spp))); lanes of a pixel are its
// samples, so rays of a warp are coherent in the BVH, the pixel is reduced with shuffles and written by one lane
// without atomics (deterministic image).  Gradient atomics are aggregated per warp before they reach L2.
// The per-sample logic itself lives in rb_render.cuh.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

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
    char* ptr = nullptr;
    size_t bytes = 0;
};
static std::mutex g_scratch_mutex;
static DeviceScratch g_scratch[64];
static char* scratch_ensure(int device, size_t bytes) {
    DeviceScratch& sc = g_scratch[device & 63];
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
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; d++)
        if (g_scratch[d].ptr) {
            cudaSetDevice(d);
            cudaDeviceSynchronize();
            cudaFree(g_scratch[d].ptr);
            g_scratch[d] = DeviceScratch();
        }
    cudaSetDevice(prev);
}
// Destroys the timing events and restores the caller's current device on EVERY exit path of rb_render (RB_CUDA_OK returns).
struct RenderGuard {
    int prev_device = -1;
    cudaEvent_t ev[5] = {};
    int num_ev = 0;
    std::vector<cudaEvent_t> band_events; // 4 per backward band: start, after trace, after compaction+secondary, after sweep
    ~RenderGuard() {
        for (cudaEvent_t e : band_events) cudaEventDestroy(e);
        for (int i = 0; i < num_ev; i++) cudaEventDestroy(ev[i]);
        if (prev_device >= 0) cudaSetDevice(prev_device);
    }
};
static int pick_grid(const void* kernel, int device, int* blocks_per_sm_out) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RB_BLOCK, 0);
    if (per_sm < 1) per_sm = 1;
    if (blocks_per_sm_out) *blocks_per_sm_out = per_sm;
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
    // [3] after k_primary_edge, [4] after k_finish_camera
    cudaEvent_t* ev = guard.ev;
    for (int i = 0; i < 5; i++) {
        RB_CUDA_OK(cudaEventCreate(&ev[i]));
        guard.num_ev = i + 1;
    }
    int launches = 0;
    double host_stats[2] = {0, 0};
    std::unique_lock<std::mutex> scratch_lock(g_scratch_mutex, std::defer_lock); // (released before the guard runs)
    std::vector<cudaEvent_t>& band_events = guard.band_events;
    RB_CUDA_OK(cudaEventRecord(ev[0], stream));
    if (image != nullptr) {
        if (only_radiance) {
            if (lean) {
                la::forward(&scene->dev, &ka, la::grid(la::K_FORWARD, scene->device), stream);
            } else {
                int grid = pick_grid((const void*)k_forward, scene->device, nullptr);
                k_forward<<<grid, RB_BLOCK, 0, stream>>>(scene->dev, ka);
            }
        } else {
            int grid = pick_grid((const void*)k_forward_channels, scene->device, nullptr);
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
        // ---- scratch layout: gradient descriptors | camera accumulators | band (records, boundary terms, lists, scan)
        const bool secondary = scene->dev.use_secondary_edge && scene->dev.num_edges > 0 && scene->dev.num_lights > 0 && rp.rad_dim >= 0;
        const long long total_samples = (long long)ka.owned_rows * rp.vp_w * rp.spp;
        ka.rec_per_sample = rp.max_bounces + 1;
        const size_t per_sample = (size_t)ka.rec_per_sample * (sizeof(VertexRec) + (secondary ? sizeof(V3) + sizeof(EdgePick) + 16 + 8 : 0) + sizeof(int)) + 2 * sizeof(int) +
                                  sizeof(unsigned long long);
        size_t band_bytes = RB_BAND_BYTES;
        if (const char* env = getenv("RB_BAND_BYTES")) { // test hook: force many small bands
            long long v = atoll(env);
            if (v > 0) band_bytes = (size_t)v;
        }
        long long band = (long long)std::max<size_t>(band_bytes / per_sample, 1024);
        band = std::min<long long>(band, (1LL << 30) / ka.rec_per_sample);
        band = std::min<long long>(band, std::max<long long>(total_samples, 1));
        size_t scan_bytes = 0;
        cub::TransformInputIterator<unsigned long long, CountOp, const int*> probe((const int*)nullptr, CountOp());
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, probe, (unsigned long long*)nullptr, (int)band, stream);
        auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t nb_shapes = std::max(1, d_scene->num_shapes) * sizeof(rb_dshape), nb_mats = std::max(1, d_scene->num_materials) * sizeof(rb_material),
               nb_lights = std::max(1, d_scene->num_lights) * sizeof(float*);
        size_t o_shapes = 0, o_mats = o_shapes + al(nb_shapes), o_lights = o_mats + al(nb_mats), o_cam = o_lights + al(nb_lights);
        size_t o_rec = o_cam + al((RB_CAM_ACC + 2) * sizeof(double)), o_dpos = o_rec + al((size_t)band * ka.rec_per_sample * sizeof(VertexRec));
        size_t o_nrec = o_dpos + al(secondary ? (size_t)band * ka.rec_per_sample * sizeof(V3) : 0);
        size_t o_offs = o_nrec + al((size_t)band * sizeof(int)), o_paths = o_offs + al((size_t)band * sizeof(unsigned long long));
        size_t o_verts = o_paths + al((size_t)band * sizeof(int)), o_tot = o_verts + al((size_t)band * ka.rec_per_sample * sizeof(int));
        size_t o_scan = o_tot + 256;
        // boundary terms: edge picks, (edge, vertex) sort buffers, radix-sort temporaries
        const size_t max_verts = (size_t)band * ka.rec_per_sample;
        size_t sec_sort_bytes = 0;
        if (secondary)
            cub::DeviceRadixSort::SortPairs(nullptr, sec_sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr, (int)max_verts, 0, 32,
                                            stream);
        size_t o_picks = o_scan + al(scan_bytes), o_sk0 = o_picks + al(secondary ? max_verts * sizeof(EdgePick) : 0);
        size_t o_sv0 = o_sk0 + al(secondary ? max_verts * 4 : 0), o_sk1 = o_sv0 + al(secondary ? max_verts * 4 : 0), o_sv1 = o_sk1 + al(secondary ? max_verts * 4 : 0);
        size_t o_ssort = o_sv1 + al(secondary ? max_verts * 4 : 0);
        size_t scratch_bytes = o_ssort + al(sec_sort_bytes);
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
        scratch_lock.lock();
        char* scratch = scratch_ensure(scene->device, scratch_bytes);
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
        RB_CUDA_OK(cudaMemsetAsync(cam_accum, 0, (RB_CAM_ACC + 2) * sizeof(double), stream));
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

        // ---- interior + first-hit adjoints, band by band: trace -> scan/compact -> boundary terms -> sweep
        ka.records = (VertexRec*)(scratch + o_rec);
        ka.dpos = secondary ? (V3*)(scratch + o_dpos) : nullptr;
        ka.nrec = (int*)(scratch + o_nrec);
        ka.offs = (unsigned long long*)(scratch + o_offs);
        ka.path_list = (int*)(scratch + o_paths);
        ka.vert_list = (int*)(scratch + o_verts);
        ka.totals = (unsigned long long*)(scratch + o_tot);
        ka.picks = (EdgePick*)(scratch + o_picks);
        ka.sec_keys = (unsigned*)(scratch + o_sk0);
        ka.sec_vals = (unsigned*)(scratch + o_sv0);
        ka.sec_keys_sorted = (unsigned*)(scratch + o_sk1);
        ka.sec_vals_sorted = (unsigned*)(scratch + o_sv1);
        int grid_t, grid_p, grid_s, grid_w;
        if (lean) {
            grid_t = la::grid(la::K_BWD_TRACE, scene->device);
            grid_p = la::grid(la::K_BWD_SEC_PICK, scene->device);
            grid_s = la::grid(la::K_BWD_SEC_SHADE, scene->device);
            grid_w = la::grid(la::K_BWD_SWEEP, scene->device);
        } else {
            grid_t = pick_grid((const void*)k_bwd_trace, scene->device, nullptr);
            grid_p = pick_grid((const void*)k_bwd_sec_pick, scene->device, nullptr);
            grid_s = pick_grid((const void*)k_bwd_sec_shade, scene->device, nullptr);
            grid_w = pick_grid((const void*)k_bwd_sweep, scene->device, nullptr);
        }
        int edge_bits = 1; // key range of the boundary-term sort: [0, num_edges]
        while ((1LL << edge_bits) <= (long long)scene->dev.num_edges && edge_bits < 32) edge_bits++;
        for (long long i0 = 0; i0 < total_samples; i0 += band) {
            ka.band_i0 = i0;
            ka.band_n = (int)std::min<long long>(band, total_samples - i0);
            cudaEvent_t e4[4];
            for (int i = 0; i < 4; i++) {
                RB_CUDA_OK(cudaEventCreate(&e4[i]));
                band_events.push_back(e4[i]);
            }
            RB_CUDA_OK(cudaEventRecord(e4[0], stream));
            if (lean) la::bwd_trace(&scene->dev, &ka, grid_t, stream);
            else k_bwd_trace<<<grid_t, RB_BLOCK, 0, stream>>>(scene->dev, ka);
            RB_CUDA_OK(cudaEventRecord(e4[1], stream));
            cub::TransformInputIterator<unsigned long long, CountOp, const int*> counts(ka.nrec, CountOp());
            cub::DeviceScan::ExclusiveSum(scratch + o_scan, scan_bytes, counts, ka.offs, ka.band_n, stream);
            k_bwd_compact<<<std::min((ka.band_n + 255) / 256, 148 * 8), 256, 0, stream>>>(ka);
            // the list sizes come back to the host once per band (the only sync inside the pass): exact grids and sort sizes
            unsigned long long totals_h = 0;
            RB_CUDA_OK(cudaMemcpyAsync(&totals_h, ka.totals, sizeof(totals_h), cudaMemcpyDeviceToHost, stream));
            RB_CUDA_OK(cudaStreamSynchronize(stream));
            ka.n_paths = (int)(totals_h >> 32);
            ka.n_verts = (int)(totals_h & 0xffffffffULL);
            launches += 4;
            if (secondary && ka.n_verts > 0) {
                if (lean) la::bwd_sec_pick(&scene->dev, &ka, grid_p, stream);
                else k_bwd_sec_pick<<<grid_p, RB_BLOCK, 0, stream>>>(scene->dev, ka);
                cub::DeviceRadixSort::SortPairs(scratch + o_ssort, sec_sort_bytes, ka.sec_keys, ka.sec_keys_sorted, ka.sec_vals, ka.sec_vals_sorted, ka.n_verts, 0, edge_bits,
                                                stream);
                if (lean) la::bwd_sec_shade(&scene->dev, &ka, grid_s, stream);
                else k_bwd_sec_shade<<<grid_s, RB_BLOCK, 0, stream>>>(scene->dev, ka);
                launches += 2 + 4;
            }
            RB_CUDA_OK(cudaEventRecord(e4[2], stream));
            if (ka.n_paths > 0) {
                if (lean) la::bwd_sweep(&scene->dev, &ka, grid_w, stream);
                else k_bwd_sweep<<<grid_w, RB_BLOCK, 0, stream>>>(scene->dev, ka);
                launches++;
            }
            RB_CUDA_OK(cudaEventRecord(e4[3], stream));
        }
        RB_CUDA_OK(cudaEventRecord(ev[2], stream));
        if (scene->dev.use_primary_edge && scene->dev.num_edges > 0 && scene->dev.prim_edge_cdf != nullptr) {
            int grid_e = lean ? la::grid(la::K_PRIMARY_EDGE, scene->device) : pick_grid((const void*)k_primary_edge, scene->device, nullptr);
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
                else k_primary_edge<<<grid_e, RB_BLOCK, 0, stream>>>(scene->dev, ka, dim_base, t0, n, k1, v1);
                launches += 2 + 4;
            }
        }
        RB_CUDA_OK(cudaEventRecord(ev[3], stream));
        k_finish_camera<<<1, 32, 0, stream>>>(scene->dev.cam, cam_accum, d_scene->camera);
        launches++;
        RB_CUDA_OK(cudaEventRecord(ev[4], stream));
        RB_CUDA_OK(cudaMemcpyAsync(host_stats, cam_accum + RB_CAM_ACC, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
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
    if (err == cudaSuccess)
        for (size_t b = 0; b + 3 < band_events.size(); b += 4)
            for (int i = 0; i < 3; i++) {
                float t = 0.f;
                cudaEventElapsedTime(&t, band_events[b + i], band_events[b + i + 1]);
                scene->last_bwd_ms[i] += t;
            }
    scene->last_path_vertices = host_stats[0];
    scene->last_primary_hits = host_stats[1];
    if (err != cudaSuccess) {
        rb_set_error(std::string("rb_render: kernel failure: ") + cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
