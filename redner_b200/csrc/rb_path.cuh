// Per-path device code: light sampling, per-vertex radiance estimate (NEE + BSDF sampling with power-2 MIS),
// the bounce loop, and the hand-derived adjoint of one path vertex.
//   light_point_sampler             src/scene.cpp:692-741
//   primary_contribs_accumulator    src/primary_contribution.cpp:6-35  (radiance channel)
//   path_contribs_accumulator       src/path_contribution.cpp:5-154
//   d_path_contribs_accumulator     src/path_contribution.cpp:156-592
//   bounce loop                     src/pathtracer.cpp:292-390
// One thread carries one path from the camera (or from an edge ray) to its end; nothing but the final pixel /
// gradient contributions leaves the SM.
#pragma once
#include "rb_bvh.cuh"
#include "rb_camera.cuh"
#include "rb_envmap.cuh"
#include "rb_material.cuh"
#include "rb_sampler.cuh"
#include "rb_shape.cuh"

// upper_bound on an ascending double table: number of entries <= x (thrust::upper_bound, src/scene.cpp:698)
RB_D int cdf_pick(const double* cdf, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return rb_clampi(lo - 1, 0, n - 1);
}

struct LightSampleRec {
    Isect isect;      // (light shape, triangle); environment map: shape_id = -1 and tri_id = bits of dir.z
    V2 uv;            // the 2-D sample used on the triangle; environment map: (dir.x, dir.y) of the sampled direction
    bool unoccluded;  // shadow ray reached the light (nee_ray.tmax >= 0)
};
// The environment-map sample keeps its world direction in the record (the adjoint pass needs exactly the direction the
// primal pass used; re-sampling from rounded random numbers would not give it).
RB_HD void light_rec_set_env_dir(LightSampleRec& r, V3 dir) {
    r.isect.shape_id = -1;
    float z = (float)dir.z;
    memcpy(&r.isect.tri_id, &z, sizeof(int));
    r.uv = mk2(dir.x, dir.y);
}
RB_HD V3 light_rec_env_dir(const LightSampleRec& r) {
    float z;
    memcpy(&z, &r.isect.tri_id, sizeof(float));
    return mk3(r.uv.x, r.uv.y, (Real)z);
}

RB_D bool closest_hit(const DevScene& sc, const Ray& ray, Isect& is) {
    float t;
    return bvh_trace<false>(sc, ray, is.shape_id, is.tri_id, t);
}
RB_D bool any_hit(const DevScene& sc, const Ray& ray) {
    int s, tr;
    float t;
    return bvh_trace<true>(sc, ray, s, tr, t);
}

// Hit point in double precision.  The reference keeps rays and hit points in double and only rounds the rays it hands to
// Embree (src/scene.cpp:556-567); a shadow ray that starts at an fp32 hit point differs from the reference's in the last
// bit and resolves differently when it grazes a blocker edge (measured: 4 of 16.8 M samples at 512x512x64 == 1.6e-4
// relative L2).  Carrying the hit point (only) in double makes the rounded shadow / bounce rays identical again.
RB_D D3 hit_point_d(const rb_shape& s, int tri, D3 o, D3 d) {
    int idx[3];
    shape_tri(s, tri, idx);
    const float *p0 = s.vertices + 3 * (size_t)idx[0], *p1 = s.vertices + 3 * (size_t)idx[1], *p2 = s.vertices + 3 * (size_t)idx[2];
    double e1x = (double)p1[0] - p0[0], e1y = (double)p1[1] - p0[1], e1z = (double)p1[2] - p0[2];
    double e2x = (double)p2[0] - p0[0], e2y = (double)p2[1] - p0[1], e2z = (double)p2[2] - p0[2];
    double pvx = d.y * e2z - d.z * e2y, pvy = d.z * e2x - d.x * e2z, pvz = d.x * e2y - d.y * e2x;
    double div = pvx * e1x + pvy * e1y + pvz * e1z;
    if (fabs(div) < 1e-8) div = div > 0 ? 1e-8 : -1e-8;
    double sx = o.x - p0[0], sy = o.y - p0[1], sz = o.z - p0[2];
    double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
    double t = (e2x * qx + e2y * qy + e2z * qz) / div;
    return d3(o.x + d.x * t, o.y + d.y * t, o.z + d.z * t);
}

// Pick a light, a triangle on it and a point on the triangle; trace the shadow ray (built in double from the double hit
// point `p_d`, then rounded once -- src/scene.cpp:692-741).
RB_D void sample_light(const DevScene& sc, D3 p_d, double light_sel, double tri_sel, double su, double sv, LightSampleRec& rec, SurfacePoint& lp) {
    int light_id = cdf_pick(sc.light_cdf, sc.num_lights, light_sel);
    if (RB_ENVMAP(sc) && light_id == sc.num_lights - 1) {
        // environment map: direction by importance sampling, shadow ray to infinity (src/scene.cpp:703-711)
        V3 dir = envmap_sample(sc.env, su, sv);
        light_rec_set_env_dir(rec, dir);
        lp = zero_point();
        Ray sh;
        sh.org = mk3((Real)p_d.x, (Real)p_d.y, (Real)p_d.z);
        sh.dir = dir;
        sh.tmin = Real(1e-3);
        sh.tmax = INFINITY;
        rec.unoccluded = !any_hit(sc, sh);
        return;
    }
    const DevLight& light = sc.lights[light_id];
    const rb_shape& shape = sc.shapes[light.shape_id];
    const double* acdf = sc.area_cdf_pool + sc.area_cdf_offset[light_id];
    int tri = cdf_pick(acdf, shape.num_triangles, tri_sel);
    rec.isect.shape_id = light.shape_id;
    rec.isect.tri_id = tri;
    rec.uv = mk2((Real)su, (Real)sv);
    lp = sample_light_triangle(shape, tri, rec.uv);
    int idx[3];
    shape_tri(shape, tri, idx);
    const float *p0 = shape.vertices + 3 * (size_t)idx[0], *p1 = shape.vertices + 3 * (size_t)idx[1], *p2 = shape.vertices + 3 * (size_t)idx[2];
    double a = sqrt(su), b1 = 1.0 - a, b2 = a * sv;
    double lx = p0[0] + ((double)p1[0] - p0[0]) * b1 + ((double)p2[0] - p0[0]) * b2;
    double ly = p0[1] + ((double)p1[1] - p0[1]) * b1 + ((double)p2[1] - p0[1]) * b2;
    double lz = p0[2] + ((double)p1[2] - p0[2]) * b1 + ((double)p2[2] - p0[2]) * b2;
    double dx = lx - p_d.x, dy = ly - p_d.y, dz = lz - p_d.z;
    double len = sqrt(dx * dx + dy * dy + dz * dz);
    Ray sh;
    sh.org = mk3((Real)p_d.x, (Real)p_d.y, (Real)p_d.z);
    sh.dir = len > 0 ? mk3((Real)(dx / len), (Real)(dy / len), (Real)(dz / len)) : zero3();
    sh.tmin = Real(1e-3);
    sh.tmax = (Real)((1 - 1e-3f) * len);
    rec.unoccluded = !any_hit(sc, sh);
}

// Emission seen along a primary / edge ray (radiance channel of accumulate_primary_contribs).
RB_D V3 hit_emission(const DevScene& sc, const Isect& is, const SurfacePoint& sp, V3 wi) {
    if (!is.valid()) return zero3();
    const rb_shape& shape = sc.shapes[is.shape_id];
    if (shape.light_id >= 0) {
        const DevLight& light = sc.lights[shape.light_id];
        if (light.directly_visible && (light.two_sided || dot(wi, sp.shading_frame.n) > 0))
            return mk3(light.intensity[0], light.intensity[1], light.intensity[2]);
    }
    return zero3();
}

// Radiance of the environment map seen along a ray that left the scene (src/primary_contribution.cpp:25-29).
RB_D V3 miss_emission(const DevScene& sc, V3 dir, const RayDiff& rd) {
    if (!RB_ENVMAP(sc) || !sc.env.directly_visible) return zero3();
    return envmap_eval(sc.env, dir, rd);
}

RB_D Real mis_power2(Real p_other, Real p_this) {
    double r = (double)p_other / (double)p_this;
    return (Real)(1.0 / (1.0 + r * r));
}

// Radiance estimate at one vertex: returns nee + scatter (not yet multiplied by the throughput) and the
// throughput factor for the next vertex.
RB_D V3 vertex_estimate(const DevScene& sc, const rb_material& mat, const SurfacePoint& sp, V3 wi, Real min_rough, const LightSampleRec& ls,
                        const SurfacePoint& lp, const Isect& bis, const SurfacePoint& bp, V3 bdir, V3& scatter_factor, bool& scatter_ok) {
    // Both estimators exist for two kinds of light (area light / environment map).  Each first settles the direction and
    // what the light contributes along it, then ONE bsdf_eval / bsdf_pdf pair serves either kind (the BSDF with its texture
    // lookups is the bulk of this function's code).
    V3 nee = zero3();
    if (ls.unoccluded) {
        V3 wo = zero3(), Le = zero3();
        Real pdf_nee = 0, G = 1;
        bool on = false;
        if (ls.isect.valid()) { // area light, src/path_contribution.cpp:28-49
            const rb_shape& lshape = sc.shapes[ls.isect.shape_id];
            V3 dir = lp.position - sp.position;
            Real dist_sq = length_sq(dir);
            wo = dir / sqrt(dist_sq);
            if (dist_sq > Real(1e-20) && lshape.light_id >= 0) {
                const DevLight& light = sc.lights[lshape.light_id];
                if (light.two_sided || dot(-wo, lp.shading_frame.n) > 0) {
                    G = fabs(dot(wo, lp.geom_normal)) / dist_sq;
                    pdf_nee = (Real)(sc.light_pmf[lshape.light_id] / sc.light_areas[lshape.light_id]);
                    Le = mk3(light.intensity[0], light.intensity[1], light.intensity[2]);
                    on = true;
                }
            }
        } else if (RB_ENVMAP(sc)) { // environment light (:51-67); the lookup is unfiltered (zero ray differential)
            wo = light_rec_env_dir(ls);
            pdf_nee = envmap_pdf(sc.env, wo) * (Real)sc.light_pmf[sc.num_lights - 1];
            if (pdf_nee > 0) {
                Le = envmap_eval(sc.env, wo, zero_raydiff());
                on = true;
            }
        }
        if (on) {
            V3 f = bsdf_eval(mat, sp, wi, wo, min_rough);
            Real pdf_b = bsdf_pdf(mat, sp, wi, wo, min_rough) * G;
            nee = (mis_power2(pdf_b, pdf_nee) * G / pdf_nee) * f * Le;
        }
    }
    V3 scatter = zero3();
    scatter_factor = zero3();
    scatter_ok = false;
    const bool hit = bis.valid();
    if (hit || RB_ENVMAP(sc)) {
        V3 wo = bdir;
        Real dist_sq = 1;
        if (hit) {
            V3 dir = bp.position - sp.position;
            dist_sq = length_sq(dir);
            wo = dir / sqrt(dist_sq);
        }
        Real pdf_b = bsdf_pdf(mat, sp, wi, wo, min_rough);
        // (hit: src/path_contribution.cpp:71-98; miss: :99-118 -- bdir is zero when the BSDF sample failed)
        if ((hit ? dist_sq > Real(1e-20) : length_sq(wo) > 0) && pdf_b > Real(1e-20)) {
            V3 f = bsdf_eval(mat, sp, wi, wo, min_rough);
            if (hit) {
                const rb_shape& bshape = sc.shapes[bis.shape_id];
                if (bshape.light_id >= 0) {
                    const DevLight& light = sc.lights[bshape.light_id];
                    if (light.two_sided || dot(-wo, bp.shading_frame.n) > 0) {
                        Real G = fabs(dot(wo, bp.geom_normal)) / dist_sq;
                        Real pdf_nee = (Real)(sc.light_pmf[bshape.light_id] * (1.0 / sc.light_areas[bshape.light_id])) / G;
                        scatter = (mis_power2(pdf_nee, pdf_b) / pdf_b) * f * mk3(light.intensity[0], light.intensity[1], light.intensity[2]);
                    }
                }
                scatter_factor = f / pdf_b;
                scatter_ok = true;
            } else {
                V3 Le = envmap_eval(sc.env, wo, zero_raydiff());
                Real pdf_nee = envmap_pdf(sc.env, wo) * (Real)sc.light_pmf[sc.num_lights - 1];
                scatter = (mis_power2(pdf_nee, pdf_b) / pdf_b) * f * Le;
            }
        }
    }
    return nee + scatter;
}

// What the adjoint pass needs to know about path vertex d (everything else is recomputed).  16-byte aligned so that
// the records move through HBM / L2 as 128-bit loads and stores (a quarter of the memory instructions).
struct alignas(16) VertexRec {
    Ray ray;       // ray that reached the vertex
    RayDiff rd_in; // its differential before the hit
    Isect isect;
    V3 thr;
    Real min_rough;
    LightSampleRec light;
};

// Follows a path from an already intersected vertex through at most (max_bounces - depth_begin) bounces and
// returns the sum of throughput-weighted vertex estimates.  `smp` must be positioned at the light-sample
// dimension of depth `depth_begin`.  If REC, vertices are written to rec[depth - depth_begin] and *num_rec is
// the number of vertices at which an estimate was formed.
template <bool REC>
RB_D V3 trace_bounces(const DevScene& sc, Sampler& smp, Ray ray, RayDiff rd_in, Isect is, V3 thr, Real min_rough, int depth_begin,
                      int max_bounces, VertexRec* rec, int rec_stride, int* num_rec, const D3* ray_org_d = nullptr, const D3* ray_dir_d = nullptr) {
    V3 L = zero3();
    int count = 0;
    // double-precision copy of the current ray (exact for camera rays, promoted fp32 otherwise)
    D3 od = ray_org_d ? *ray_org_d : d3(ray.org.x, ray.org.y, ray.org.z);
    D3 dd = ray_dir_d ? *ray_dir_d : d3(ray.dir.x, ray.dir.y, ray.dir.z);
    if (sc.num_lights > 0) {
        for (int depth = depth_begin; depth < max_bounces && is.valid(); depth++) {
            RayDiff rd;
            SurfacePoint sp = make_surface_point(sc.shapes[is.shape_id], is.tri_id, ray, rd_in, rd);
            const rb_material& mat = sc.materials[sc.shapes[is.shape_id].material_id];
            V3 wi = -ray.dir;
            double l_sel = smp.next(), t_sel = smp.next(), lu = smp.next(), lv = smp.next();
            LightSampleRec ls;
            SurfacePoint lp;
            D3 p_d = hit_point_d(sc.shapes[is.shape_id], is.tri_id, od, dd);
            sample_light(sc, p_d, l_sel, t_sel, lu, lv, ls, lp);
            double bu = smp.next(), bv = smp.next(), bw = smp.next();
            if (REC) {
                VertexRec& r = rec[(size_t)count * rec_stride];
                r.ray = ray;
                r.rd_in = rd_in;
                r.isect = is;
                r.thr = thr;
                r.min_rough = min_rough;
                r.light = ls;
            }
            RayDiff rd_b;
            Real next_rough;
            V3 dir = bsdf_sample_dir(mat, sp, wi, mk2((Real)bu, (Real)bv), bw, min_rough, rd, rd_b, next_rough);
            Ray nray;
            nray.org = mk3((Real)p_d.x, (Real)p_d.y, (Real)p_d.z);
            nray.dir = dir;
            od = p_d;
            dd = d3(dir.x, dir.y, dir.z);
            nray.tmin = Real(1e-3);
            nray.tmax = INFINITY;
            Isect bis = no_isect();
            SurfacePoint bp = zero_point();
            RayDiff rd_after;
            if (closest_hit(sc, nray, bis)) bp = make_surface_point(sc.shapes[bis.shape_id], bis.tri_id, nray, rd_b, rd_after);
            V3 factor;
            bool ok;
            V3 est = vertex_estimate(sc, mat, sp, wi, min_rough, ls, lp, bis, bp, dir, factor, ok);
            L += thr * est;
            count++;
            thr = ok ? thr * factor : zero3();
            ray = nray;
            rd_in = rd_b;
            is = bis;
            min_rough = next_rough;
        }
    }
    if (REC) {
        // terminal vertex: no estimate is formed there, but the adjoint of the previous vertex needs its hit
        VertexRec& r = rec[(size_t)count * rec_stride];
        r.ray = ray;
        r.rd_in = rd_in;
        r.isect = is;
        r.thr = thr;
        r.min_rough = min_rough;
        *num_rec = count;
    }
    return L;
}

// Gradient sinks for geometry: per-corner scatter with warp aggregation.
RB_D void scatter_vertex_grads(const DevScene& sc, const DevDScene& ds, const Isect& is, const V3 d_vp[3], const V3 d_vn[3], const V2 d_vuv[3],
                               const V3 d_vc[3]) {
    const rb_shape& s = sc.shapes[is.shape_id];
    const rb_dshape& d = ds.shapes[is.shape_id];
    TriAttribs a;
    tri_attribs(s, is.tri_id, a);
    for (int k = 0; k < 3; k++) {
        if (d.vertices) agg_add3(d.vertices + 3 * (size_t)a.ind[k], d_vp[k]);
        if (s.uvs && d.uvs) agg_add2(d.uvs + 2 * (size_t)a.uv_ind[k], d_vuv[k]);
        if (s.normals && d.normals) agg_add3(d.normals + 3 * (size_t)a.n_ind[k], d_vn[k]);
        if (s.colors && d.colors) agg_add3(d.colors + 3 * (size_t)a.ind[k], d_vc[k]);
    }
}

// Adjoint state flowing from vertex d+1 to vertex d.
struct VertexAdjoint {
    V3 d_thr;
    DRay d_ray;
    SurfacePoint d_point;
};
RB_D VertexAdjoint zero_vertex_adjoint() {
    VertexAdjoint a;
    a.d_thr = zero3();
    a.d_ray = zero_dray();
    a.d_point = zero_point();
    return a;
}

// Adjoint of vertex_estimate + throughput update at vertex `cur`, given the adjoint arriving from vertex `nxt`.
// d_contrib = weight * d_image[pixel] (radiance channels).
RB_D VertexAdjoint d_vertex(const DevScene& sc, const DevDScene& ds, const VertexRec& cur, const VertexRec* nxt, V3 d_contrib,
                            const VertexAdjoint& next) {
    VertexAdjoint out = zero_vertex_adjoint();
    const rb_shape& shape = sc.shapes[cur.isect.shape_id];
    const rb_material& mat = sc.materials[shape.material_id];
    const rb_material& d_mat = ds.materials[shape.material_id];
    RayDiff rd;
    SurfacePoint sp = make_surface_point(shape, cur.isect.tri_id, cur.ray, cur.rd_in, rd);
    V3 wi = -cur.ray.dir;
    V3 p = sp.position;
    V3 thr = cur.thr;
    Real min_rough = cur.min_rough;
    // The two estimators of this vertex (light sample, BSDF sample) share ONE rolled call of d_bsdf_eval: each prepares
    // (wo, d_f, d_wo) in a "pre" block and consumes the returned d_wo in a "post" block.  Besides halving the code of
    // the largest adjoint this reconverges the lanes of a warp that took only one of the two branches.
    bool on_l = false, on_b = false, env_l = false, env_b = false;
    V3 wo_l = zero3(), d_f_l = zero3(), d_wo_l = zero3(), dir_l = zero3();
    V3 wo_b = zero3(), d_f_b = zero3(), d_wo_b = zero3(), dir_b = zero3();
    Real dist_sq_l = 1, d_dist_sq_l = 0, dist_sq_b = 1, d_cos_l = 0;
    V3 d_lv[3] = {zero3(), zero3(), zero3()};
    // ---- next event estimation (pre): area light (src/path_contribution.cpp:212-293) or environment map (:295-337).
    // Direction and light-side quantities first, then ONE bsdf_eval / bsdf_pdf pair for either kind.
    if (cur.light.unoccluded) {
        const bool area = cur.light.isect.valid();
        const Isect& lis = cur.light.isect;
        SurfacePoint lp = zero_point();
        V3 wo = zero3(), dir = zero3(), Le = zero3();
        Real dist_sq = 1, pdf_nee = 0;
        bool ok = false;
        if (area) {
            const rb_shape& lshape = sc.shapes[lis.shape_id];
            lp = sample_light_triangle(lshape, lis.tri_id, cur.light.uv);
            dir = lp.position - p;
            dist_sq = length_sq(dir);
            wo = dir / sqrt(dist_sq);
            if (lshape.light_id >= 0) {
                const DevLight& light = sc.lights[lshape.light_id];
                if (light.two_sided || dot(-wo, lp.shading_frame.n) > 0) {
                    Le = mk3(light.intensity[0], light.intensity[1], light.intensity[2]);
                    pdf_nee = (Real)(sc.light_pmf[lshape.light_id] * (1.0 / sc.light_areas[lshape.light_id]));
                    ok = true;
                }
            }
        } else if (RB_ENVMAP(sc)) {
            wo = light_rec_env_dir(cur.light);
            pdf_nee = envmap_pdf(sc.env, wo) * (Real)sc.light_pmf[sc.num_lights - 1];
            if (pdf_nee > 0) {
                Le = envmap_eval(sc.env, wo, zero_raydiff());
                ok = true;
            }
        }
        if (ok) {
            V3 f = bsdf_eval(mat, sp, wi, wo, min_rough);
            Real pdf_b0 = bsdf_pdf(mat, sp, wi, wo, min_rough);
            V3 d_nee = d_contrib * thr;
            if (area) {
                const rb_shape& lshape = sc.shapes[lis.shape_id];
                Real cos_l = dot(wo, lp.geom_normal);
                Real G = fabs(cos_l) / dist_sq;
                Real mis = mis_power2(pdf_b0 * G, pdf_nee);
                out.d_thr += d_contrib * ((mis * G / pdf_nee) * f * Le);
                Real wgt = mis / pdf_nee;
                // derivatives of the MIS weight and of the light-selection pmf are ignored (src/path_contribution.cpp:239)
                Real d_wgt = G * sum(d_nee * f * Le);
                Real d_pdf_nee = -d_wgt * wgt / pdf_nee;
                Real d_G = wgt * sum(d_nee * f * Le);
                Real d_area = -d_pdf_nee * pdf_nee / shape_tri_area(lshape, lis.tri_id);
                d_shape_tri_area(lshape, lis.tri_id, d_area, d_lv);
                agg_add3(ds.light_intensity[lshape.light_id], wgt * G * (d_nee * f));
                d_cos_l = cos_l > 0 ? d_G / dist_sq : -d_G / dist_sq;
                dir_l = dir;
                dist_sq_l = dist_sq;
                d_dist_sq_l = -d_G * G / dist_sq;
                d_f_l = wgt * G * (d_nee * Le);
                d_wo_l = d_cos_l * lp.geom_normal;
            } else { // no dependence of the direction on the vertex position
                Real wgt = mis_power2(pdf_b0, pdf_nee) / pdf_nee;
                out.d_thr += d_contrib * (wgt * f * Le);
                RayDiff d_rd0 = zero_raydiff();
                d_envmap_eval(sc.env, wo, zero_raydiff(), wgt * (d_nee * f), ds.env_values, ds.env_w2e, d_wo_l, d_rd0);
                env_l = true;
                d_f_l = wgt * (d_nee * Le);
            }
            on_l = true;
            wo_l = wo;
        }
    }
    // ---- BSDF-sampled continuation (pre): the ray hit something (:339-518) or left the scene into the map (:520-590)
    SurfacePoint bp;
    if (nxt != nullptr && (nxt->isect.valid() || RB_ENVMAP(sc))) {
        const bool hit = nxt->isect.valid();
        const Isect& bis = nxt->isect;
        V3 wo = nxt->ray.dir, dir = zero3();
        Real dist_sq = 1;
        if (hit) {
            RayDiff rd_after;
            bp = make_surface_point(sc.shapes[bis.shape_id], bis.tri_id, nxt->ray, nxt->rd_in, rd_after);
            dir = bp.position - p;
            dist_sq = length_sq(dir);
            wo = dir / sqrt(dist_sq);
        }
        Real pdf_b = bsdf_pdf(mat, sp, wi, wo, min_rough);
        if (pdf_b > 0 && (hit || length_sq(wo) > 0)) {
            V3 f = bsdf_eval(mat, sp, wi, wo, min_rough);
            V3 d_scatter = d_contrib * thr;
            if (hit) {
                const rb_shape& bshape = sc.shapes[bis.shape_id];
                out.d_thr += next.d_thr * (f / pdf_b);
                // the derivative w.r.t. pdf_bsdf is dropped on purpose (src/path_contribution.cpp:369-376)
                V3 d_f = (next.d_thr * thr) / pdf_b;
                if (bshape.light_id >= 0) {
                    const DevLight& light = sc.lights[bshape.light_id];
                    if (light.two_sided || dot(-wo, bp.shading_frame.n) > 0) {
                        Real G = fabs(dot(wo, bp.geom_normal)) / dist_sq;
                        V3 Le = mk3(light.intensity[0], light.intensity[1], light.intensity[2]);
                        Real pdf_nee = (Real)(sc.light_pmf[bshape.light_id] * (1.0 / sc.light_areas[bshape.light_id])) / G;
                        Real wgt = mis_power2(pdf_nee, pdf_b) / pdf_b;
                        out.d_thr += d_contrib * (wgt * f * Le);
                        d_f += wgt * (d_scatter * Le);
                        agg_add3(ds.light_intensity[bshape.light_id], wgt * (d_scatter * f));
                    }
                }
                dir_b = dir;
                dist_sq_b = dist_sq;
                d_f_b = d_f;
                d_wo_b = next.d_ray.dir;
            } else { // nothing flows back into the sampling procedure
                V3 Le = envmap_eval(sc.env, wo, zero_raydiff());
                Real pdf_nee = envmap_pdf(sc.env, wo) * (Real)sc.light_pmf[sc.num_lights - 1];
                Real wgt = mis_power2(pdf_nee, pdf_b) / pdf_b;
                out.d_thr += d_contrib * (wgt * f * Le);
                RayDiff d_rd0 = zero_raydiff();
                d_envmap_eval(sc.env, wo, zero_raydiff(), wgt * (d_scatter * f), ds.env_values, ds.env_w2e, d_wo_b, d_rd0);
                env_b = true;
                d_f_b = wgt * (d_scatter * Le);
            }
            on_b = true;
            wo_b = wo;
        }
    }
    // ---- shared BSDF adjoint
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        if (!(k ? on_b : on_l)) continue;
        V3 d_wi = zero3();
        V3 d_wo = k ? d_wo_b : d_wo_l;
        d_bsdf_eval(mat, d_mat, sp, wi, k ? wo_b : wo_l, min_rough, k ? d_f_b : d_f_l, out.d_point, d_wi, d_wo);
        out.d_ray.dir -= d_wi;
        if (k) d_wo_b = d_wo; else d_wo_l = d_wo;
    }
    // ---- next event estimation (post)
    if (on_l && !env_l) {
        const Isect& lis = cur.light.isect;
        const rb_shape& lshape = sc.shapes[lis.shape_id];
        V3 d_dir = d_wo_l / sqrt(dist_sq_l);
        Real d_sqrt = -sum(d_wo_l * dir_l) / dist_sq_l;
        Real d_dist_sq = d_dist_sq_l + Real(0.5) * d_sqrt / sqrt(dist_sq_l);
        d_dir += d_length_sq(dir_l, d_dist_sq);
        SurfacePoint d_lp = zero_point();
        d_lp.geom_normal = d_cos_l * wo_l;
        d_lp.position += d_dir;
        out.d_point.position -= d_dir;
        d_sample_light_triangle(lshape, lis.tri_id, cur.light.uv, d_lp, d_lv);
        int idx[3];
        shape_tri(lshape, lis.tri_id, idx);
        float* dv = ds.shapes[lis.shape_id].vertices;
        if (dv) {
            agg_add3(dv + 3 * (size_t)idx[0], d_lv[0]);
            agg_add3(dv + 3 * (size_t)idx[1], d_lv[1]);
            agg_add3(dv + 3 * (size_t)idx[2], d_lv[2]);
        }
    }
    // ---- BSDF-sampled continuation (post)
    if (on_b && !env_b) {
        const Isect& bis = nxt->isect;
        const rb_shape& bshape = sc.shapes[bis.shape_id];
        V3 d_bvp[3] = {zero3(), zero3(), zero3()}, d_bvn[3] = {zero3(), zero3(), zero3()}, d_bvc[3] = {zero3(), zero3(), zero3()};
        V2 d_bvuv[3] = {zero2(), zero2(), zero2()};
        V3 d_dir = d_wo_b / sqrt(dist_sq_b);
        Real d_sqrt = -sum(d_wo_b * dir_b) / dist_sq_b;
        Real d_dist_sq = Real(0.5) * d_sqrt / sqrt(dist_sq_b);
        d_dir += d_length_sq(dir_b, d_dist_sq);
        SurfacePoint d_bp = next.d_point;
        d_bp.position += d_dir;
        DRay d_ray = zero_dray();
        RayDiff d_rd_b = zero_raydiff();
        Ray bray;
        bray.org = sp.position;
        bray.dir = wo_b;
        bray.tmin = Real(1e-3);
        bray.tmax = INFINITY;
        d_make_surface_point(bshape, bis.tri_id, bray, nxt->rd_in, d_bp, zero_raydiff(), d_ray, d_rd_b, d_bvp, d_bvn, d_bvuv, d_bvc);
        // position gradient through the sampled direction only below glossy vertices (src/path_contribution.cpp:447-455)
        if (min_rough > Real(0.01)) {
            out.d_point.position -= d_dir;
            out.d_point.position += d_ray.org;
        }
        scatter_vertex_grads(sc, ds, bis, d_bvp, d_bvn, d_bvuv, d_bvc);
    }
    return out;
}
