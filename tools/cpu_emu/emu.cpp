// DEVELOPMENT AID ONLY (see emu_shim.h).  Exports the same C ABI as libredner_b200.so, but every buffer is a HOST
// pointer and the "kernels" are plain loops over the per-sample functions of rb_render.cuh.  The triangle BVH is a
// simple median-split tree in the same node format (the GPU LBVH builder itself is exercised on the GPU only).
#include "emu_shim.h"

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../redner_b200/csrc/rb_render.cuh"
#include "../../redner_b200/csrc/rb_scene_host.hpp"

static thread_local std::string g_err;
extern "C" const char* rb_last_error(void) { return g_err.c_str(); }
extern "C" const char* rb_version(void) { return "redner_b200 CPU emulator (debug only)"; }

struct rb_scene {
    DevScene dev;
    rb_camera cam;
    std::vector<rb_shape> shapes;
    std::vector<rb_material> materials;
    std::vector<DevLight> lights;
    HostLightTables lt;
    HostEdgeTables et;
    HostEdgeTree tree;
    std::vector<float> ltc;
    std::vector<BVHNode> nodes;
    std::vector<BVHTri> tris;
    std::vector<unsigned long long> sobol;
    int max_generic = 0;
    int part = 0, num_parts = 1, rps = 16;
};

static int build_node(rb_scene* sc, std::vector<int>& order, std::vector<float>& boxes, int lo, int hi, float out_box[6]) {
    // returns child reference (>=0 inner, <0 leaf) for the range [lo, hi) of `order`
    if (hi - lo == 1) {
        for (int k = 0; k < 6; k++) out_box[k] = boxes[6 * (size_t)order[lo] + k];
        return ~order[lo];
    }
    float cb[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = lo; i < hi; i++)
        for (int a = 0; a < 3; a++) {
            float c = 0.5f * (boxes[6 * (size_t)order[i] + a] + boxes[6 * (size_t)order[i] + 3 + a]);
            cb[a] = std::min(cb[a], c);
            cb[3 + a] = std::max(cb[3 + a], c);
        }
    int axis = 0;
    for (int a = 1; a < 3; a++)
        if (cb[3 + a] - cb[a] > cb[3 + axis] - cb[axis]) axis = a;
    int mid = (lo + hi) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](int x, int y) {
        return boxes[6 * (size_t)x + axis] + boxes[6 * (size_t)x + 3 + axis] < boxes[6 * (size_t)y + axis] + boxes[6 * (size_t)y + 3 + axis];
    });
    int me = (int)sc->nodes.size();
    sc->nodes.push_back(BVHNode());
    float l[6], r[6];
    int left = build_node(sc, order, boxes, lo, mid, l);
    int right = build_node(sc, order, boxes, mid, hi, r);
    BVHNode& n = sc->nodes[me];
    n.left = left;
    n.right = right;
    n.pad0 = n.pad1 = 0;
    n.lo_x_hi_x = make_float4(l[0], l[3], r[0], r[3]);
    n.lo_y_hi_y = make_float4(l[1], l[4], r[1], r[4]);
    n.lo_z_hi_z = make_float4(l[2], l[5], r[2], r[5]);
    for (int k = 0; k < 3; k++) {
        out_box[k] = std::min(l[k], r[k]);
        out_box[3 + k] = std::max(l[3 + k], r[3 + k]);
    }
    return me;
}

extern "C" int rb_scene_create(const rb_scene_desc* desc, rb_scene** out) {
    rb_scene* sc = new rb_scene();
    memset(&sc->dev, 0, sizeof(DevScene));
    sc->cam = desc->camera;
    host_setup_camera(desc->camera, sc->dev.cam);
    sc->shapes.assign(desc->shapes, desc->shapes + desc->num_shapes);
    sc->materials.assign(desc->materials, desc->materials + desc->num_materials);
    for (int l = 0; l < desc->num_lights; l++) {
        DevLight dl;
        dl.shape_id = desc->lights[l].shape_id;
        for (int k = 0; k < 3; k++) dl.intensity[k] = desc->lights[l].intensity[k];
        dl.two_sided = desc->lights[l].two_sided;
        dl.directly_visible = desc->lights[l].directly_visible;
        sc->lights.push_back(dl);
    }
    for (const rb_material& m : sc->materials)
        if (m.generic_texture.num_levels > 0) sc->max_generic = std::max(sc->max_generic, m.generic_texture.channels);
    DevScene& d = sc->dev;
    d.edge_root_cs = d.edge_root_ncs = RB_EDGE_EMPTY;
    d.shapes = sc->shapes.data();
    d.num_shapes = (int)sc->shapes.size();
    d.materials = sc->materials.data();
    d.num_materials = (int)sc->materials.size();
    d.use_primary_edge = desc->use_primary_edge_sampling;
    d.use_secondary_edge = desc->use_secondary_edge_sampling;
    // tables
    FILE* f = fopen(RB_DATA_DIR "/sobol_joe_kuo_1024x52_u64.bin", "rb");
    if (!f) { g_err = "emu: sobol table not found"; return 1; }
    sc->sobol.resize(1024 * 52);
    if (fread(sc->sobol.data(), 8, sc->sobol.size(), f) != sc->sobol.size()) { g_err = "emu: short sobol table"; return 1; }
    fclose(f);
    d.sobol_matrices = sc->sobol.data();
    d.sobol_dims = 1024;
    // BVH
    std::vector<float> boxes;
    for (int s = 0; s < d.num_shapes; s++)
        for (int t = 0; t < sc->shapes[s].num_triangles; t++) {
            V3 v0, v1, v2;
            shape_tri_vertices(sc->shapes[s], t, v0, v1, v2);
            BVHTri tr;
            tr.v0 = make_float4((float)v0.x, (float)v0.y, (float)v0.z, __int_as_float(s));
            tr.v1 = make_float4((float)v1.x, (float)v1.y, (float)v1.z, __int_as_float(t));
            tr.v2 = make_float4((float)v2.x, (float)v2.y, (float)v2.z, 0.f);
            sc->tris.push_back(tr);
            for (int a = 0; a < 3; a++) {
                float lo = std::min((float)v0[a], std::min((float)v1[a], (float)v2[a])), hi = std::max((float)v0[a], std::max((float)v1[a], (float)v2[a]));
                float pad = std::max(std::fabs(lo), std::fabs(hi)) * 4e-7f + 1e-6f;
                boxes.push_back(lo - pad);
            }
            for (int a = 0; a < 3; a++) {
                float hi = std::max((float)v0[a], std::max((float)v1[a], (float)v2[a]));
                float lo = std::min((float)v0[a], std::min((float)v1[a], (float)v2[a]));
                float pad = std::max(std::fabs(lo), std::fabs(hi)) * 4e-7f + 1e-6f;
                boxes.push_back(hi + pad);
            }
        }
    int T = (int)sc->tris.size();
    d.num_tris = T;
    if (T > 0) {
        std::vector<int> order(T);
        std::iota(order.begin(), order.end(), 0);
        float box[6];
        d.bvh_root = build_node(sc, order, boxes, 0, T, box);
        if (sc->nodes.empty()) sc->nodes.push_back(BVHNode());
        d.bvh_nodes = sc->nodes.data();
        d.bvh_tris = sc->tris.data();
    }
    // lights + edges
    std::vector<HostMesh> meshes(d.num_shapes);
    for (int s = 0; s < d.num_shapes; s++) {
        meshes[s].vertices.assign(sc->shapes[s].vertices, sc->shapes[s].vertices + 3 * (size_t)sc->shapes[s].num_vertices);
        meshes[s].indices.assign(sc->shapes[s].indices, sc->shapes[s].indices + 3 * (size_t)sc->shapes[s].num_triangles);
    }
    d.num_lights = (int)sc->lights.size();
    d.has_envmap = desc->envmap != nullptr;
    if (d.has_envmap) {
        const rb_envmap& e = *desc->envmap;
        d.env.values = e.values;
        memcpy(d.env.w2e, e.world_to_env, sizeof(d.env.w2e));
        memcpy(d.env.e2w, e.env_to_world, sizeof(d.env.e2w));
        d.env.cdf_ys = e.sample_cdf_ys;
        d.env.cdf_xs = e.sample_cdf_xs;
        d.env.pdf_norm = e.pdf_norm;
        d.env.directly_visible = e.directly_visible;
        d.num_lights++;
    }
    if (d.num_lights > 0) {
        if (!host_build_lights(sc->lights, meshes, sc->lt, g_err, d.has_envmap != 0, d.has_envmap ? desc->envmap->pdf_norm : 0.0, host_bsphere_radius(meshes))) return 1;
        d.lights = sc->lights.data();
        d.light_pmf = sc->lt.pmf.data();
        d.light_cdf = sc->lt.cdf.data();
        d.light_areas = sc->lt.areas.data();
        d.area_cdf_pool = sc->lt.pool.data();
        d.area_cdf_offset = sc->lt.offsets.data();
    }
    if (d.use_primary_edge || d.use_secondary_edge) {
        host_build_edges(sc->shapes, meshes, d.cam, d.use_primary_edge != 0, sc->et);
        d.edges = sc->et.edges.data();
        d.num_edges = (int)sc->et.edges.size();
        d.prim_edge_pmf = sc->et.prim_pmf.data();
        d.prim_edge_cdf = sc->et.prim_cdf.data();
        d.edge_root_cs = d.edge_root_ncs = RB_EDGE_EMPTY;
        if (d.use_secondary_edge) {
            host_build_edge_tree(sc->shapes, meshes, sc->et.edges, d.cam, sc->tree);
            d.edge_nodes = sc->tree.nodes.data();
            d.edge_root_cs = sc->tree.root_cs;
            d.edge_root_ncs = sc->tree.root_ncs;
            d.edge_bounds_expand = sc->tree.expand;
            FILE* fl = fopen(RB_DATA_DIR "/ltc_blinn_phong_128x128x9_f32.bin", "rb");
            if (!fl) { g_err = "emu: ltc table not found"; return 1; }
            sc->ltc.resize(128 * 128 * 9);
            if (fread(sc->ltc.data(), 4, sc->ltc.size(), fl) != sc->ltc.size()) { g_err = "emu: short ltc table"; return 1; }
            fclose(fl);
            d.ltc_table = sc->ltc.data();
        }
    }
    *out = sc;
    return 0;
}
// Re-target at another camera: the camera-dependent tables (primary-edge distribution, both edge trees) are rebuilt with the host
// builders; geometry, BVH, lights and the edge list are kept (mirrors rb_scene_set_camera of the product, which rebuilds them on the GPU).
extern "C" int rb_scene_set_camera(rb_scene* sc, const rb_camera* cam) {
    DevScene& d = sc->dev;
    sc->cam = *cam;
    host_setup_camera(*cam, d.cam);
    if (d.num_edges > 0) {
        std::vector<HostMesh> meshes(d.num_shapes);
        for (int s = 0; s < d.num_shapes; s++) {
            meshes[s].vertices.assign(sc->shapes[s].vertices, sc->shapes[s].vertices + 3 * (size_t)sc->shapes[s].num_vertices);
            meshes[s].indices.assign(sc->shapes[s].indices, sc->shapes[s].indices + 3 * (size_t)sc->shapes[s].num_triangles);
        }
        if (d.use_primary_edge) {
            host_primary_edge_distribution(sc->shapes, meshes, d.cam, sc->et);
            d.prim_edge_pmf = sc->et.prim_pmf.data();
            d.prim_edge_cdf = sc->et.prim_cdf.data();
        }
        if (d.use_secondary_edge) {
            host_build_edge_tree(sc->shapes, meshes, sc->et.edges, d.cam, sc->tree);
            d.edge_nodes = sc->tree.nodes.data();
            d.edge_root_cs = sc->tree.root_cs;
            d.edge_root_ncs = sc->tree.root_ncs;
            d.edge_bounds_expand = sc->tree.expand;
        }
    }
    return 0;
}
extern "C" void rb_scene_destroy(rb_scene* sc) { delete sc; }
extern "C" int rb_scene_max_generic_texture_dimension(const rb_scene* sc) { return sc->max_generic; }
extern "C" int rb_compute_num_channels(const int* ch, int n, int mg) { return host_compute_num_channels(ch, n, mg); }
extern "C" int rb_scene_set_partition(rb_scene* sc, int part, int num_parts, int rps) {
    sc->part = part; sc->num_parts = num_parts; sc->rps = rps;
    return 0;
}
extern "C" int rb_scene_last_stage_stats(const rb_scene*, float* ms, double* v, double* h) {
    if (ms) for (int i = 0; i < 4; i++) ms[i] = 0;
    if (v) *v = 0;
    if (h) *h = 0;
    return 0;
}
extern "C" int rb_scene_build_ms(const rb_scene*, float* ms) { ms[0] = ms[1] = ms[2] = 0; return 0; }
extern "C" int rb_scene_last_stats(const rb_scene*, int* n, float* ms) {
    if (n) *n = 0;
    if (ms) *ms = 0;
    return 0;
}

extern "C" int rb_render(const rb_scene* scene, const rb_options* opt, float* image, const float* d_image, const rb_dscene_desc* d_scene,
                         float* screen_grad, void*) {
    KernelArgs ka;
    memset(&ka, 0, sizeof(ka));
    RenderParams& rp = ka.rp;
    rp.seed = opt->seed;
    rp.spp = opt->num_samples;
    rp.max_bounces = opt->max_bounces;
    rp.sampler_type = opt->sampler_type;
    rp.sample_pixel_center = opt->sample_pixel_center;
    rp.num_channels = opt->num_channels;
    rp.rad_dim = -1;
    for (int i = 0; i < opt->num_channels; i++)
        if (opt->channels[i] == RB_CH_RADIANCE) rp.rad_dim = i;
    rp.nd = host_compute_num_channels(opt->channels, opt->num_channels, scene->max_generic);
    rp.rad_off = -1;
    for (int i = 0, off = 0; i < opt->num_channels; i++) {
        if (opt->channels[i] == RB_CH_RADIANCE) rp.rad_off = off;
        off += rb_channel_width(opt->channels[i], scene->max_generic);
    }
    rp.part = scene->part; rp.num_parts = scene->num_parts; rp.rows_per_stripe = scene->rps; // round-robin stripes of rows, as in rb_kernels_body.cuh
    auto owned = [&](int y) { return (y / rp.rows_per_stripe) % rp.num_parts == rp.part; };
    rp.vp_w = scene->cam.viewport_end[0] - scene->cam.viewport_beg[0];
    rp.vp_h = scene->cam.viewport_end[1] - scene->cam.viewport_beg[1];
    ka.lanes_per_pixel = 1;
    while (ka.lanes_per_pixel * 2 <= std::min(32, rp.spp)) ka.lanes_per_pixel *= 2;
    ka.image = image;
    ka.d_image = d_image;
    ka.screen_grad = screen_grad;
    const DevScene& sc = scene->dev;
    for (int i = 0; i < opt->num_channels; i++) rp.channels[i] = opt->channels[i];
    rp.max_generic = scene->max_generic;
    bool only_radiance = opt->num_channels == 1 && opt->channels[0] == RB_CH_RADIANCE;
    rp.only_radiance = only_radiance ? 1 : 0;
    if (image && !only_radiance) {
        for (int y = 0; y < rp.vp_h; y++)
            for (int x = 0; x < rp.vp_w; x++) {
                if (!owned(y)) continue;
                int pixel = y * rp.vp_w + x;
                float acc[RB_MAX_ND] = {0};
                int ids[3] = {-1, -1, -1}, last = -1;
                for (int s = 0; s < rp.spp; s++) {
                    int cur[3] = {-1, -1, -1};
                    if (forward_sample_channels(sc, rp, pixel, x, y, s, acc, cur)) { last = s; ids[0] = cur[0]; ids[1] = cur[1]; ids[2] = cur[2]; }
                }
                float* px = image + (size_t)rp.nd * pixel;
                int d = 0;
                for (int c = 0; c < rp.num_channels; c++) {
                    int ch = rp.channels[c];
                    int width = (ch == RB_CH_RADIANCE || ch == RB_CH_POSITION || ch == RB_CH_GEOMETRY_NORMAL || ch == RB_CH_SHADING_NORMAL ||
                                 ch == RB_CH_DIFFUSE_REFLECTANCE || ch == RB_CH_SPECULAR_REFLECTANCE || ch == RB_CH_VERTEX_COLOR) ? 3
                              : (ch == RB_CH_UV || ch == RB_CH_BARYCENTRIC) ? 2 : (ch == RB_CH_GENERIC_TEXTURE ? rp.max_generic : 1);
                    if (ch == RB_CH_SHAPE_ID || ch == RB_CH_TRIANGLE_ID || ch == RB_CH_MATERIAL_ID) {
                        if (last >= 0) px[d] = (float)ids[ch - RB_CH_SHAPE_ID];
                    } else {
                        for (int i = 0; i < width; i++) px[d + i] += acc[d + i];
                    }
                    d += width;
                }
            }
    } else if (image) {
        for (int y = 0; y < rp.vp_h; y++)
            for (int x = 0; x < rp.vp_w; x++) {
                if (!owned(y)) continue;
                int pixel = y * rp.vp_w + x;
                V3 acc = zero3();
                for (int s = 0; s < rp.spp; s++) acc += forward_sample(sc, rp, pixel, x, y, s);
                float* px = image + (size_t)rp.nd * pixel + rp.rad_dim;
                px[0] += (float)acc.x; px[1] += (float)acc.y; px[2] += (float)acc.z;
            }
    }
    if (d_image) {
        std::vector<double> cam_accum(RB_CAM_ACC, 0.0);
        std::vector<float> cam_f(RB_CAM_ACC, 0.f);
        ka.ds.shapes = d_scene->shapes;
        ka.ds.materials = d_scene->materials;
        ka.ds.light_intensity = d_scene->light_intensity;
        ka.ds.cam_accum = cam_accum.data();
        memset(&ka.ds.env_values, 0, sizeof(rb_texture));
        ka.ds.env_w2e = nullptr;
        if (d_scene->envmap != nullptr) {
            ka.ds.env_values = d_scene->envmap->values;
            ka.ds.env_w2e = d_scene->envmap->world_to_env;
        }
        CamAcc acc;
        acc.base = cam_f.data();
        acc.stride = 1;
        std::vector<VertexRec> recs(rp.max_bounces + 2);
#ifdef RB_EMU_REF_STREAMS
        // rank of every (sample, pixel) in the reference's compacted active list of each depth (pixels in ascending order)
        const int mbr = std::max(rp.max_bounces, 1);
        const size_t npx = (size_t)rp.vp_w * rp.vp_h;
        std::vector<int> ranks((size_t)rp.spp * npx * mbr, 0);
        for (int s = 0; s < rp.spp; s++) {
            std::vector<int> counter(mbr, 0);
            for (size_t pixel = 0; pixel < npx; pixel++) {
                int nrec = bwd_trace(sc, rp, (int)pixel, (int)(pixel % rp.vp_w), (int)(pixel / rp.vp_w), s, recs.data(), 1);
                for (int d = 0; d < nrec && d < mbr; d++) ranks[((size_t)s * npx + pixel) * mbr + d] = counter[d]++;
            }
        }
#endif
        for (int y = 0; y < rp.vp_h; y++)
            for (int x = 0; x < rp.vp_w; x++)
                for (int s = 0; s < rp.spp; s++) {
                    if (!owned(y)) continue;
#ifdef RB_EMU_REF_STREAMS
                    rb_emu_rank = &ranks[((size_t)s * npx + (size_t)y * rp.vp_w + x) * mbr];
#endif
                    backward_sample(sc, ka, y * rp.vp_w + x, x, y, s, recs.data(), acc);
                    for (int k = 0; k < RB_CAM_ACC; k++) { cam_accum[k] += cam_f[k]; cam_f[k] = 0.f; }
                }
        if (sc.use_primary_edge && sc.num_edges > 0) {
            long long n_px = (long long)rp.vp_w * rp.vp_h;
            for (long long i = 0; i < n_px; i++)
                for (int s = 0; s < rp.spp; s++) {
                    if (i % rp.num_parts != rp.part) continue; // primary-edge samples are sharded by sample index
                    primary_edge_sample(sc, ka, i, s, primary_edge_dim_base(sc, rp), acc);
                    for (int k = 0; k < RB_CAM_ACC; k++) { cam_accum[k] += cam_f[k]; cam_f[k] = 0.f; }
                }
        }
        finish_camera(sc.cam, cam_accum.data(), d_scene->camera);
    }
    return 0;
}
