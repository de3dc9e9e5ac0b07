// Render kernels and the rb_render driver (reference: render(), src/pathtracer.cpp:177-958).
//
// Execution model.  The reference runs a host-driven wavefront: ~20 launches and 2 host syncs per bounce per
// sample, with every per-path field (5.5 - 9.4 KB/pixel in double) streamed through managed memory between
// stage functors.  Here the whole per-sample pipeline is ONE persistent kernel per pass:
//   k_forward        camera sample -> primary hit -> emission -> bounce loop -> pixel
//   k_backward       forward replay (compact per-vertex records, L2-resident) -> reverse sweep over the path with
//                    the hand-derived adjoints -> first-hit / camera adjoint   [+ secondary edge sampling]
//   k_primary_edge   one thread per primary-edge sample: edge pick, two offset rays, two full sub-paths, Eq. 8
// A warp owns 32/L pixels with L lanes per pixel (L = min(32, 2^floor(log2 spp))); lanes of a pixel are its
// samples, so rays of a warp are coherent in the BVH, the pixel is reduced with shuffles and written by one lane
// without atomics (deterministic image), and gradient atomics are aggregated per warp before they reach L2.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "rb_edge.cuh"
#include "rb_path.cuh"
#include "rb_scene.cuh"

#define RB_BLOCK 128

struct KernelArgs {
    RenderParams rp;
    int lanes_per_pixel; // L
    int owned_rows;      // rows of the viewport this device renders
    float* image;        // forward
    const float* d_image;
    float* screen_grad;
    DevDScene ds;
    VertexRec* records; // [threads][max_bounces + 1]
    int rec_per_thread;
};

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
    int px, py;   // absolute pixel coordinates
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
RB_D unsigned long long main_draws_per_sample(const RenderParams& rp) {
    return (unsigned long long)((rp.sample_pixel_center ? 0 : 2) + 7 * rp.max_bounces);
}

// Camera sample -> primary ray (viewport offset applied), src/camera.cpp:8-43.
RB_D void primary_ray_for(const DevScene& sc, const RenderParams& rp, const WorkItem& w, Sampler& smp, double& sx, double& sy, Ray& ray,
                          RayDiff& rd) {
    double jx = 0.5, jy = 0.5;
    if (!rp.sample_pixel_center) {
        jx = smp.next();
        jy = smp.next();
    }
    sx = (double(w.px + sc.cam.vp_beg[0]) + jx) / double(sc.cam.width);
    sy = (double(w.py + sc.cam.vp_beg[1]) + jy) / double(sc.cam.height);
    cam_primary_ray(sc.cam, sx, sy, ray, rd);
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(RB_BLOCK) k_forward(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    const RenderParams& rp = ka.rp;
    const int L = ka.lanes_per_pixel;
    const int P = 32 / L;
    long long n_px = (long long)ka.owned_rows * rp.vp_w;
    long long groups = (n_px + P - 1) / P;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nb = (rp.spp + L - 1) / L;
    const Real weight = Real(1) / Real(rp.spp);
    for (long long g = warp; g < groups; g += nwarps) {
        WorkItem w = warp_work(rp, L, ka.owned_rows, g);
        V3 acc = zero3();
        for (int b = 0; b < nb; b++) {
            int s = b * L + w.sample_lane;
            if (w.valid && s < rp.spp) {
                Sampler smp;
                smp.init(rp.sampler_type, rp.seed, w.pixel, (unsigned)s, sc.sobol_matrices, RB_SOBOL_BITS, (unsigned long long)s * main_draws_per_sample(rp));
                double sx, sy;
                Ray ray;
                RayDiff rd;
                primary_ray_for(sc, rp, w, smp, sx, sy, ray, rd);
                Isect is = no_isect();
                if (closest_hit(sc, ray, is)) {
                    RayDiff rd_after;
                    SurfacePoint sp = make_surface_point(sc.shapes[is.shape_id], is.tri_id, ray, rd, rd_after);
                    V3 L0 = hit_emission(sc, is, sp, -ray.dir);
                    acc += weight * L0;
                    V3 Lb = trace_bounces<false>(sc, smp, ray, rd, is, mk3(1, 1, 1), Real(0), 0, rp.max_bounces, nullptr, 0, nullptr);
                    acc += weight * Lb;
                }
            }
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

__global__ void __launch_bounds__(RB_BLOCK) k_backward(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka) {
    __shared__ float cam_smem[RB_CAM_ACC * RB_BLOCK];
    for (int k = 0; k < RB_CAM_ACC; k++) cam_smem[k * RB_BLOCK + threadIdx.x] = 0.f;
    CamAcc cam_acc;
    cam_acc.base = cam_smem + threadIdx.x;
    cam_acc.stride = RB_BLOCK;

    const RenderParams& rp = ka.rp;
    const DevDScene& ds = ka.ds;
    const int L = ka.lanes_per_pixel;
    const int P = 32 / L;
    long long n_px = (long long)ka.owned_rows * rp.vp_w;
    long long groups = (n_px + P - 1) / P;
    long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long warp = gtid >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nb = (rp.spp + L - 1) / L;
    const Real weight = Real(1) / Real(rp.spp);
    VertexRec* recs = ka.records + (size_t)gtid * ka.rec_per_thread;
    for (long long g = warp; g < groups; g += nwarps) {
        WorkItem w = warp_work(rp, L, ka.owned_rows, g);
        for (int b = 0; b < nb; b++) {
            int s = b * L + w.sample_lane;
            if (!(w.valid && s < rp.spp)) continue;
            Sampler smp;
            smp.init(rp.sampler_type, rp.seed, w.pixel, (unsigned)s, sc.sobol_matrices, RB_SOBOL_BITS, (unsigned long long)s * main_draws_per_sample(rp));
            double sx, sy;
            Ray ray;
            RayDiff rd;
            primary_ray_for(sc, rp, w, smp, sx, sy, ray, rd);
            Isect is = no_isect();
            if (!closest_hit(sc, ray, is)) continue;
            const float* dpx = ka.d_image + (size_t)rp.nd * w.pixel + rp.rad_dim;
            V3 d_contrib = weight * mk3(dpx[0], dpx[1], dpx[2]);
            int nrec = 0;
            trace_bounces<true>(sc, smp, ray, rd, is, mk3(1, 1, 1), Real(0), 0, rp.max_bounces, recs, 1, &nrec);
            // reverse sweep over the interior vertices (src/pathtracer.cpp:431-714)
            VertexAdjoint adj = zero_vertex_adjoint();
            for (int d = nrec - 1; d >= 0; d--) {
                VertexRec cur = recs[d];
                VertexRec nxt = recs[d + 1];
                adj = d_vertex(sc, ds, cur, &nxt, d_contrib, adj);
            }
            // first vertex: emission adjoint (src/primary_contribution.cpp:449-466) ...
            RayDiff rd_after;
            SurfacePoint sp = make_surface_point(sc.shapes[is.shape_id], is.tri_id, ray, rd, rd_after);
            {
                const rb_shape& shape = sc.shapes[is.shape_id];
                V3 wi = -ray.dir;
                if (shape.light_id >= 0 && dot(wi, sp.shading_frame.n) > 0) {
                    const DevLight& light = sc.lights[shape.light_id];
                    if (light.directly_visible) agg_add3(ds.light_intensity[shape.light_id], d_contrib);
                }
            }
            // ... and the hit itself back to the mesh and the camera (src/primary_intersection.cpp:5-130)
            V3 d_vp[3] = {zero3(), zero3(), zero3()}, d_vn[3] = {zero3(), zero3(), zero3()}, d_vc[3] = {zero3(), zero3(), zero3()};
            V2 d_vuv[3] = {zero2(), zero2(), zero2()};
            DRay d_ray = adj.d_ray;
            RayDiff d_prd = zero_raydiff();
            d_make_surface_point(sc.shapes[is.shape_id], is.tri_id, ray, rd, adj.d_point, zero_raydiff(), d_ray, d_prd, d_vp, d_vn, d_vuv, d_vc);
            scatter_vertex_grads(sc, ds, is, d_vp, d_vn, d_vuv, d_vc);
            const Real delta = Real(1e-3);
            Real psx = Real(0.5) / sc.cam.width, psy = Real(0.5) / sc.cam.height;
            DRay d_ray_dx, d_ray_dy;
            d_ray_dx.org = d_prd.org_dx * (psx / delta);
            d_ray_dx.dir = d_prd.dir_dx * (psx / delta);
            d_ray_dy.org = d_prd.org_dy * (psy / delta);
            d_ray_dy.dir = d_prd.dir_dy * (psy / delta);
            d_ray.org += (d_prd.org_dx * (-psx) + d_prd.org_dy * (-psy)) / delta;
            d_ray.dir += (d_prd.dir_dx * (-psx) + d_prd.dir_dy * (-psy)) / delta;
            V2 d_screen = zero2();
            V2* d_screen_ptr = ka.screen_grad ? &d_screen : nullptr;
            d_cam_sample_primary(sc.cam, (Real)sx, (Real)sy, d_ray, cam_acc, d_screen_ptr);
            d_cam_sample_primary(sc.cam, (Real)sx + delta, (Real)sy, d_ray_dx, cam_acc, d_screen_ptr);
            d_cam_sample_primary(sc.cam, (Real)sx, (Real)sy + delta, d_ray_dy, cam_acc, d_screen_ptr);
            if (ka.screen_grad) {
                atomicAdd(&ka.screen_grad[2 * (size_t)w.pixel + 0], (float)d_screen.x);
                atomicAdd(&ka.screen_grad[2 * (size_t)w.pixel + 1], (float)d_screen.y);
            }
        }
    }
    block_reduce_camera(cam_smem, ds.cam_accum);
}

// ------------------------------------------------------------------------------------------------ primary edges
// Projection of an edge in double (the +-1e-6 offsets across the edge need more than fp32 screen coordinates).
struct D2 {
    double x, y;
};
RB_D D3 w2c_point(const DevCamera& cam, D3 p) {
    const double* W = cam.w2c;
    double x = W[0] * p.x + W[1] * p.y + W[2] * p.z + W[3];
    double y = W[4] * p.x + W[5] * p.y + W[6] * p.z + W[7];
    double z = W[8] * p.x + W[9] * p.y + W[10] * p.z + W[11];
    double w = W[12] * p.x + W[13] * p.y + W[14] * p.z + W[15];
    double iw = 1.0 / w;
    return d3(x * iw, y * iw, z * iw);
}
RB_D D2 cam_to_screen_d(const DevCamera& cam, D3 p) {
    const double* K = cam.intr;
    double aspect = double(cam.width) / double(cam.height);
    double ix = K[0] * p.x + K[1] * p.y + K[2] * p.z, iy = K[3] * p.x + K[4] * p.y + K[5] * p.z, iz = K[6] * p.x + K[7] * p.y + K[8] * p.z;
    D2 r;
    if (cam.type == RB_CAMERA_PERSPECTIVE) {
        r.x = (ix / iz + 1.0) * 0.5;
        r.y = (-(iy / iz) * aspect + 1.0) * 0.5;
    } else {
        r.x = (ix + 1.0) * 0.5;
        r.y = (-iy * aspect + 1.0) * 0.5;
    }
    return r;
}
RB_D bool cam_project_d(const DevCamera& cam, D3 p0, D3 p1, D2& q0, D2& q1) {
    D3 a = w2c_point(cam, p0), b = w2c_point(cam, p1);
    double cn = cam.clip_near;
    if (a.z < cn && b.z < cn) return false;
    if (a.z < cn) {
        D3 dir = d3(a.x - b.x, a.y - b.y, a.z - b.z);
        double t = -(b.z - cn) / dir.z;
        a = d3(b.x + t * dir.x, b.y + t * dir.y, b.z + t * dir.z);
    } else if (b.z < cn) {
        D3 dir = d3(b.x - a.x, b.y - a.y, b.z - a.z);
        double t = -(a.z - cn) / dir.z;
        b = d3(a.x + t * dir.x, a.y + t * dir.y, a.z + t * dir.z);
    }
    q0 = cam_to_screen_d(cam, a);
    q1 = cam_to_screen_d(cam, b);
    return true;
}
RB_D unsigned long long edge_draws_per_sample(const RenderParams& rp) { return (unsigned long long)(2 + 7 * rp.max_bounces); }

// One thread per (edge sample i, spp sample s).  Reference: primary_edge_sampler src/edge.cpp:385-625, the sub-path
// loop src/pathtracer.cpp:766-934 and primary_edge_derivatives_computer src/edge.cpp:700-783.
__global__ void __launch_bounds__(RB_BLOCK) k_primary_edge(const __grid_constant__ DevScene sc, const __grid_constant__ KernelArgs ka, int dim_base) {
    __shared__ float cam_smem[RB_CAM_ACC * RB_BLOCK];
    for (int k = 0; k < RB_CAM_ACC; k++) cam_smem[k * RB_BLOCK + threadIdx.x] = 0.f;
    CamAcc cam_acc;
    cam_acc.base = cam_smem + threadIdx.x;
    cam_acc.stride = RB_BLOCK;
    const RenderParams& rp = ka.rp;
    const DevDScene& ds = ka.ds;
    const long long n_px = (long long)rp.vp_w * rp.vp_h;
    // samples of this device: i with i % num_parts == part
    const long long n_mine = (n_px - rp.part + rp.num_parts - 1) / rp.num_parts;
    const long long total = n_mine * rp.spp;
    const Real weight = Real(1) / Real(rp.spp);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        // consecutive threads share the edge-sample index and differ in the spp sample -> coherent edge picks per warp
        long long i = (t / rp.spp) * rp.num_parts + rp.part;
        int s = (int)(t % rp.spp);
        Sampler smp;
        smp.init(rp.sampler_type, rp.seed + 131071ULL, (int)i, (unsigned)s, sc.sobol_matrices, RB_SOBOL_BITS, (unsigned long long)s * edge_draws_per_sample(rp));
        smp.dim = dim_base;
        double e_sel = smp.next(), e_t = smp.next();
        int edge_id = cdf_pick(sc.prim_edge_cdf, sc.num_edges, e_sel);
        double pmf = sc.prim_edge_pmf[edge_id];
        const Edge edge = sc.edges[edge_id];
        V3 v0 = edge_v0(sc.shapes, edge), v1 = edge_v1(sc.shapes, edge);
        D2 q0, q1;
        if (!cam_project_d(sc.cam, d3(v0.x, v0.y, v0.z), d3(v1.x, v1.y, v1.z), q0, q1)) continue;
        if (pmf <= 0) continue;
        D2 ept;
        ept.x = q0.x + e_t * (q1.x - q0.x);
        ept.y = q0.y + e_t * (q1.y - q0.y);
        if (!cam_in_screen(sc.cam, mk2((Real)ept.x, (Real)ept.y))) continue;
        // unit normal of the projected edge: get_normal(normalize(v0_ss - v1_ss)) = (d.y, -d.x)
        double ddx = q0.x - q1.x, ddy = q0.y - q1.y;
        double dl = sqrt(ddx * ddx + ddy * ddy);
        double nx = ddy / dl, ny = -ddx / dl;
        const double offset = 1e-6;
        int vp_w = rp.vp_w;
        int xi = rb_clampi(int(ept.x * sc.cam.width - sc.cam.vp_beg[0]), 0, sc.cam.vp_end[0] - sc.cam.vp_beg[0]);
        int yi = rb_clampi(int(ept.y * sc.cam.height - sc.cam.vp_beg[1]), 0, sc.cam.vp_end[1] - sc.cam.vp_beg[1]);
        const float* dpx = ka.d_image + (size_t)rp.nd * ((size_t)yi * vp_w + xi) + rp.rad_dim;
        V3 d_color = mk3(dpx[0], dpx[1], dpx[2]);
        V3 wgt = d_color / (Real)pmf;
        // ray differential of the un-offset ray, shared by both sides (src/edge.cpp:594-608)
        Ray cray;
        RayDiff rd;
        cam_primary_ray(sc.cam, ept.x, ept.y, cray, rd);
        Real contrib = 0;
        for (int side = 0; side < 2; side++) {
            double sgn = side == 0 ? 1.0 : -1.0;
            D3 o, d;
            cam_sample_primary(sc.cam, ept.x + sgn * nx * offset, ept.y + sgn * ny * offset, o, d);
            Ray ray = make_ray(o, d);
            V3 thr = side == 0 ? wgt : -wgt;
            Isect is = no_isect();
            if (!closest_hit(sc, ray, is)) continue;
            RayDiff rd_after;
            SurfacePoint sp = make_surface_point(sc.shapes[is.shape_id], is.tri_id, ray, rd, rd_after);
            contrib += sum(weight * thr * hit_emission(sc, is, sp, -ray.dir));
            Sampler sub = smp; // both sides consume the same light / bsdf samples (src/pathtracer.cpp:871-886)
            V3 Lb = trace_bounces<false>(sc, sub, ray, rd, is, thr, Real(0), 0, rp.max_bounces, nullptr, 0, nullptr);
            contrib += sum(weight * Lb);
        }
        if (contrib == 0) continue;
        // Eq. 8: gradients of the edge equation w.r.t. the projected end points
        Real d0x = (Real)(q1.y - ept.y) * contrib, d0y = (Real)(ept.x - q1.x) * contrib;
        Real d1x = (Real)(ept.y - q0.y) * contrib, d1y = (Real)(q0.x - ept.x) * contrib;
        V3 d_v0 = zero3(), d_v1 = zero3();
        d_cam_project(sc.cam, v0, v1, d0x, d0y, d1x, d1y, cam_acc, d_v0, d_v1);
        float* dv = ds.shapes[edge.shape_id].vertices;
        if (dv) {
            agg_add3(dv + 3 * (size_t)edge.v0, d_v0);
            agg_add3(dv + 3 * (size_t)edge.v1, d_v1);
        }
        if (ka.screen_grad) {
            Real dex = (Real)(q0.y - q1.y) * contrib, dey = (Real)(q1.x - q0.x) * contrib;
            size_t pix = (size_t)yi * vp_w + xi;
            atomicAdd(&ka.screen_grad[2 * pix + 0], (float)dex);
            atomicAdd(&ka.screen_grad[2 * pix + 1], (float)dey);
        }
    }
    block_reduce_camera(cam_smem, ds.cam_accum);
}

// ------------------------------------------------------------------------------------------------ camera finish
// Turns the reduced matrix gradients into the user-facing camera gradients (one thread; everything is linear in the
// accumulated matrices): d_project's world_to_cam term (src/camera.h:811-829) and d_look_at_matrix (src/transform.h:29-71).
__global__ void k_finish_camera(DevCamera cam, const double* acc, rb_dcamera out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double C[4][4], W[4][4], Dw[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            C[i][j] = acc[4 * i + j];
            Dw[i][j] = acc[16 + 4 * i + j];
            W[i][j] = cam.w2c[4 * i + j];
        }
    // d_cam_to_world += -W^T * d_W * W^T
    double tmp[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += W[k][i] * Dw[k][j];
            tmp[i][j] = s;
        }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += tmp[i][k] * W[j][k];
            C[i][j] -= s;
        }
    if (cam.use_look_at) {
        M4 d_m;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) d_m.m[i][j] = (Real)C[i][j];
        V3 pos = mk3((Real)cam.position[0], (Real)cam.position[1], (Real)cam.position[2]);
        V3 look = mk3((Real)cam.look[0], (Real)cam.look[1], (Real)cam.look[2]);
        V3 up = mk3((Real)cam.up[0], (Real)cam.up[1], (Real)cam.up[2]);
        V3 d_p = zero3(), d_l = zero3(), d_u = zero3();
        d_look_at_matrix(pos, look, up, d_m, d_p, d_l, d_u);
        if (out.position) { out.position[0] += (float)d_p.x; out.position[1] += (float)d_p.y; out.position[2] += (float)d_p.z; }
        if (out.look) { out.look[0] += (float)d_l.x; out.look[1] += (float)d_l.y; out.look[2] += (float)d_l.z; }
        if (out.up) { out.up[0] += (float)d_u.x; out.up[1] += (float)d_u.y; out.up[2] += (float)d_u.z; }
    } else if (out.cam_to_world) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) out.cam_to_world[4 * i + j] += (float)C[i][j];
    }
    if (out.intrinsic_mat_inv)
        for (int k = 0; k < 9; k++) out.intrinsic_mat_inv[k] += (float)acc[32 + k];
    if (out.intrinsic_mat)
        for (int k = 0; k < 9; k++) out.intrinsic_mat[k] += (float)acc[41 + k];
}

// ------------------------------------------------------------------------------------------------ driver
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
    for (int i = 0; i < opt->num_channels; i++) {
        rp.channels[i] = opt->channels[i];
        if (opt->channels[i] == RB_CH_RADIANCE) {
            if (rp.rad_dim != -1) {
                rb_set_error("Duplicated radiance channel"); // src/channels.cpp:24-26
                return 1;
            }
            // the reference stores the CHANNEL INDEX and uses it as a float offset (src/channels.cpp:27,
            // src/path_contribution.cpp:125-129); identical whenever radiance is the first channel
            rp.rad_dim = i;
        } else {
            rb_set_error("rb_render: only the radiance channel is implemented so far (G-buffer channels: SURVEY.md 8f rank 3)");
            return 1;
        }
    }
    if (rp.rad_dim < 0) {
        rb_set_error("rb_render: the radiance channel is required");
        return 1;
    }
    rp.nd = rb_compute_num_channels(opt->channels, opt->num_channels, rp.max_generic);
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

    int prev = 0;
    RB_CUDA_OK(cudaGetDevice(&prev));
    RB_CUDA_OK(cudaSetDevice(scene->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaEvent_t ev0, ev1;
    RB_CUDA_OK(cudaEventCreate(&ev0));
    RB_CUDA_OK(cudaEventCreate(&ev1));
    int launches = 0;
    std::vector<void*> temps;
    auto cleanup = [&]() {
        for (void* p : temps) cudaFreeAsync(p, stream);
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        cudaSetDevice(prev);
    };
    RB_CUDA_OK(cudaEventRecord(ev0, stream));
    if (image != nullptr) {
        int grid = pick_grid((const void*)k_forward, scene->device, nullptr);
        k_forward<<<grid, RB_BLOCK, 0, stream>>>(scene->dev, ka);
        launches++;
    }
    if (d_image != nullptr) {
        // device copies of the gradient descriptor
        if (d_scene->num_shapes != (int)scene->shapes.size() || d_scene->num_materials != (int)scene->materials.size() ||
            d_scene->num_lights != (int)scene->lights.size()) {
            rb_set_error("rb_render: d_scene does not match the scene (shape / material / light counts)");
            cleanup();
            return 1;
        }
        rb_dshape* d_shapes = nullptr;
        rb_material* d_mats = nullptr;
        float** d_lights = nullptr;
        double* cam_accum = nullptr;
        size_t nb_shapes = std::max(1, d_scene->num_shapes) * sizeof(rb_dshape), nb_mats = std::max(1, d_scene->num_materials) * sizeof(rb_material),
               nb_lights = std::max(1, d_scene->num_lights) * sizeof(float*);
        RB_CUDA_OK(cudaMallocAsync((void**)&d_shapes, nb_shapes, stream));
        temps.push_back(d_shapes);
        RB_CUDA_OK(cudaMallocAsync((void**)&d_mats, nb_mats, stream));
        temps.push_back(d_mats);
        RB_CUDA_OK(cudaMallocAsync((void**)&d_lights, nb_lights, stream));
        temps.push_back(d_lights);
        RB_CUDA_OK(cudaMallocAsync((void**)&cam_accum, RB_CAM_ACC * sizeof(double), stream));
        temps.push_back(cam_accum);
        if (d_scene->num_shapes) RB_CUDA_OK(cudaMemcpyAsync(d_shapes, d_scene->shapes, d_scene->num_shapes * sizeof(rb_dshape), cudaMemcpyHostToDevice, stream));
        if (d_scene->num_materials)
            RB_CUDA_OK(cudaMemcpyAsync(d_mats, d_scene->materials, d_scene->num_materials * sizeof(rb_material), cudaMemcpyHostToDevice, stream));
        if (d_scene->num_lights)
            RB_CUDA_OK(cudaMemcpyAsync(d_lights, d_scene->light_intensity, d_scene->num_lights * sizeof(float*), cudaMemcpyHostToDevice, stream));
        RB_CUDA_OK(cudaMemsetAsync(cam_accum, 0, RB_CAM_ACC * sizeof(double), stream));
        ka.ds.shapes = d_shapes;
        ka.ds.materials = d_mats;
        ka.ds.light_intensity = d_lights;
        ka.ds.cam_accum = cam_accum;

        int grid = pick_grid((const void*)k_backward, scene->device, nullptr);
        ka.rec_per_thread = rp.max_bounces + 2;
        VertexRec* recs = nullptr;
        RB_CUDA_OK(cudaMallocAsync((void**)&recs, (size_t)grid * RB_BLOCK * ka.rec_per_thread * sizeof(VertexRec), stream));
        temps.push_back(recs);
        ka.records = recs;
        k_backward<<<grid, RB_BLOCK, 0, stream>>>(scene->dev, ka);
        launches++;
        if (scene->dev.use_primary_edge && scene->dev.num_edges > 0 && scene->dev.prim_edge_cdf != nullptr) {
            int grid_e = pick_grid((const void*)k_primary_edge, scene->device, nullptr);
            int dim_base = 0;
            k_primary_edge<<<grid_e, RB_BLOCK, 0, stream>>>(scene->dev, ka, dim_base);
            launches++;
        }
        k_finish_camera<<<1, 32, 0, stream>>>(scene->dev.cam, cam_accum, d_scene->camera);
        launches++;
    }
    RB_CUDA_OK(cudaEventRecord(ev1, stream));
    cudaError_t err = cudaStreamSynchronize(stream);
    if (err == cudaSuccess) err = cudaGetLastError();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    scene->last_launches = launches;
    scene->last_kernel_ms = ms;
    cleanup();
    if (err != cudaSuccess) {
        rb_set_error(std::string("rb_render: kernel failure: ") + cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
