// Host-side scene preprocessing shared by rb_scene.cu and the debug emulator (tools/cpu_emu):
//   light PMF/CDF + per-light triangle-area CDFs   src/scene.cpp:197-253, compute_area_cdf :38-61 (serial, double)
//   edge list + primary-edge distribution          src/edge.cpp:43-214, :233-331
//   camera matrices in double                      src/camera.h:44-55, src/transform.h:9-27
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "rb_edge.cuh"

struct HostMesh {
    std::vector<float> vertices;
    std::vector<int> indices;
};
struct HostLightTables {
    std::vector<double> pmf, cdf, areas, pool;
    std::vector<int> offsets;
};
struct HostEdgeTables {
    std::vector<Edge> edges;
    std::vector<double> prim_pmf, prim_cdf;
};

inline bool host_build_lights(const std::vector<DevLight>& lights, const std::vector<HostMesh>& meshes, HostLightTables& out, std::string& err) {
    int L = (int)lights.size();
    out.pmf.assign(L, 0);
    out.cdf.assign(L, 0);
    out.areas.assign(L, 0);
    out.offsets.assign(L, 0);
    out.pool.clear();
    double total = 0;
    for (int l = 0; l < L; l++) {
        const DevLight& light = lights[l];
        const HostMesh& m = meshes[light.shape_id];
        int T = (int)m.indices.size() / 3;
        out.offsets[l] = (int)out.pool.size();
        std::vector<double> a(T);
        double sum_area = 0; // serial sum in triangle order == thrust::reduce on the CPP backend (src/scene.cpp:43-45)
        for (int t = 0; t < T; t++) {
            const int* id = &m.indices[3 * (size_t)t];
            double v[3][3];
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 3; c++) v[k][c] = m.vertices[3 * (size_t)id[k] + c];
            double e1[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
            double e2[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
            double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            a[t] = 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
            sum_area += a[t];
        }
        double run = 0;
        for (int t = 0; t < T; t++) { // exclusive scan, then normalise
            out.pool.push_back(run / sum_area);
            run += a[t];
        }
        out.areas[l] = sum_area;
        double lum = 0.212671f * (double)light.intensity[0] + 0.715160f * (double)light.intensity[1] + 0.072169f * (double)light.intensity[2];
        out.pmf[l] = sum_area * lum * double(M_PI);
        total += out.pmf[l];
    }
    if (!(total > 0)) {
        err = "rb_scene_create: total light importance is not positive (src/scene.cpp:243)";
        return false;
    }
    for (int l = 0; l < L; l++) out.pmf[l] /= total;
    out.cdf[0] = 0;
    for (int l = 1; l < L; l++) out.cdf[l] = out.cdf[l - 1] + out.pmf[l - 1];
    return true;
}

inline bool host_pos_less(const float* a, const float* b) { // strict lexicographic order on positions
    if (a[0] != b[0]) return a[0] < b[0];
    if (a[1] != b[1]) return a[1] < b[1];
    return a[2] < b[2];
}
inline bool host_pos_eq(const float* a, const float* b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }

// `shapes` carries device (or emulator-host) pointers; only material / light ids and the null-ness of `normals` are
// read from it here, geometry comes from `meshes`.
inline void host_build_edges(const std::vector<rb_shape>& shapes, const std::vector<HostMesh>& meshes, const DevCamera& cam, bool want_primary,
                             HostEdgeTables& out) {
    int S = (int)shapes.size();
    std::vector<rb_shape> hs(shapes);
    for (int s = 0; s < S; s++) {
        hs[s].vertices = meshes[s].vertices.data();
        hs[s].indices = meshes[s].indices.data();
    }
    std::vector<Edge> edges;
    for (int s = 0; s < S; s++) {
        const HostMesh& m = meshes[s];
        int T = (int)m.indices.size() / 3;
        std::vector<Edge> he(3 * (size_t)T);
        for (int t = 0; t < T; t++) {
            const int* id = &m.indices[3 * (size_t)t];
            for (int k = 0; k < 3; k++) {
                int a = id[k], b = id[(k + 1) % 3];
                Edge e;
                e.shape_id = s;
                e.v0 = std::min(a, b);
                e.v1 = std::max(a, b);
                e.f0 = t;
                e.f1 = -1;
                he[3 * (size_t)t + k] = e;
            }
        }
        std::stable_sort(he.begin(), he.end(), [](const Edge& x, const Edge& y) { return x.v0 != y.v0 ? x.v0 < y.v0 : x.v1 < y.v1; });
        // merge runs of equal (v0, v1): f0 of the first, f1 = f0 of the last (src/edge.cpp:86-90, :266-273)
        std::vector<Edge> merged;
        for (size_t i = 0; i < he.size();) {
            size_t j = i + 1;
            while (j < he.size() && he[j].v0 == he[i].v0 && he[j].v1 == he[i].v1) j++;
            Edge e = he[i];
            if (j - i >= 2) e.f1 = he[j - 1].f0;
            merged.push_back(e);
            i = j;
        }
        // seam repair: sort by end-point POSITIONS and pair up unmatched duplicates (src/edge.cpp:103-166, :280-288)
        const float* V = m.vertices.data();
        auto key = [&](const Edge& e, const float*& lo, const float*& hi) {
            lo = V + 3 * (size_t)e.v0;
            hi = V + 3 * (size_t)e.v1;
            if (host_pos_less(hi, lo)) std::swap(lo, hi);
        };
        std::stable_sort(merged.begin(), merged.end(), [&](const Edge& x, const Edge& y) {
            const float *xl, *xh, *yl, *yh;
            key(x, xl, xh);
            key(y, yl, yh);
            if (!host_pos_eq(xl, yl)) return host_pos_less(xl, yl);
            if (!host_pos_eq(xh, yh)) return host_pos_less(xh, yh);
            return false;
        });
        std::vector<int> new_f1(merged.size());
        for (size_t i = 0; i < merged.size(); i++) {
            new_f1[i] = merged[i].f1;
            if (merged[i].f1 != -1) continue;
            const float *l, *h, *cl, *ch;
            key(merged[i], l, h);
            if (i > 0) {
                key(merged[i - 1], cl, ch);
                if (host_pos_eq(l, cl) && host_pos_eq(h, ch)) new_f1[i] = merged[i - 1].f0;
            }
            if (i + 1 < merged.size()) {
                key(merged[i + 1], cl, ch);
                if (host_pos_eq(l, cl) && host_pos_eq(h, ch)) new_f1[i] = merged[i + 1].f0;
            }
        }
        for (size_t i = 0; i < merged.size(); i++) {
            merged[i].f1 = new_f1[i];
            edges.push_back(merged[i]);
        }
    }
    // drop edges between coplanar faces (src/edge.cpp:293-296)
    out.edges.clear();
    for (const Edge& e : edges)
        if (!edge_is_flat(hs.data(), e)) out.edges.push_back(e);
    int E = (int)out.edges.size();
    out.prim_pmf.assign(E, 0);
    out.prim_cdf.assign(E, 0);
    if (!want_primary || E == 0) return;
    // screen-space length of camera silhouettes -> PMF / CDF (src/edge.cpp:186-214, :298-331)
    double iw = 1.0 / cam.c2w[15];
    V3 org = mk3((Real)(cam.c2w[3] * iw), (Real)(cam.c2w[7] * iw), (Real)(cam.c2w[11] * iw));
    double total = 0;
    for (int i = 0; i < E; i++) {
        const Edge& e = out.edges[i];
        V3 v0 = edge_v0(hs.data(), e), v1 = edge_v1(hs.data(), e);
        V2 p0, p1, c0, c1;
        double w = 0;
        if (cam_project(cam, v0, v1, p0, p1) && clip_line_unit(p0, p1, c0, c1) && edge_is_silhouette(hs.data(), org, e)) w = length(c1 - c0);
        out.prim_pmf[i] = w;
        total += w;
    }
    double run = 0;
    for (int i = 0; i < E; i++) {
        out.prim_pmf[i] = total > 0 ? out.prim_pmf[i] / total : 0.0;
        out.prim_cdf[i] = run;
        run += out.prim_pmf[i];
    }
}

inline void host_look_at(const float* pos, const float* look, const float* up, double* m) {
    auto norm = [](double* v) {
        double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; } else { v[0] = v[1] = v[2] = 0; }
    };
    auto crs = [](const double* a, const double* b, double* c) {
        c[0] = a[1] * b[2] - a[2] * b[1];
        c[1] = a[2] * b[0] - a[0] * b[2];
        c[2] = a[0] * b[1] - a[1] * b[0];
    };
    double d[3] = {(double)look[0] - pos[0], (double)look[1] - pos[1], (double)look[2] - pos[2]};
    norm(d);
    double u[3] = {up[0], up[1], up[2]};
    norm(u);
    double r[3];
    crs(d, u, r);
    norm(r);
    double nu[3];
    crs(r, d, nu);
    norm(nu);
    double o[16] = {r[0], nu[0], d[0], pos[0], r[1], nu[1], d[1], pos[1], r[2], nu[2], d[2], pos[2], 0, 0, 0, 1};
    std::memcpy(m, o, sizeof(o));
}
inline void host_inverse4(const double* m, double* o) {
    double A[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            A[i][j] = m[4 * i + j];
            A[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        for (int k = 0; k < 8; k++) std::swap(A[c][k], A[piv][k]);
        double d = A[c][c];
        for (int k = 0; k < 8; k++) A[c][k] /= d;
        for (int r = 0; r < 4; r++)
            if (r != c) {
                double f = A[r][c];
                for (int k = 0; k < 8; k++) A[r][k] -= f * A[c][k];
            }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[4 * i + j] = A[i][4 + j];
}
// rb_camera (C ABI) -> DevCamera (double copies of the host-read parameters)
inline void host_setup_camera(const rb_camera& c, DevCamera& dc) {
    dc.width = c.width;
    dc.height = c.height;
    dc.use_look_at = c.use_look_at;
    for (int i = 0; i < 3; i++) {
        dc.position[i] = c.position[i];
        dc.look[i] = c.look[i];
        dc.up[i] = c.up[i];
    }
    if (c.use_look_at) {
        host_look_at(c.position, c.look, c.up, dc.c2w);
        host_inverse4(dc.c2w, dc.w2c);
    } else {
        for (int i = 0; i < 16; i++) {
            dc.c2w[i] = c.cam_to_world[i];
            dc.w2c[i] = c.world_to_cam[i];
        }
    }
    for (int i = 0; i < 9; i++) {
        dc.intr_inv[i] = c.intrinsic_mat_inv[i];
        dc.intr[i] = c.intrinsic_mat[i];
    }
    dc.clip_near = c.clip_near;
    dc.type = c.camera_type;
    dc.vp_beg[0] = c.viewport_beg[0];
    dc.vp_beg[1] = c.viewport_beg[1];
    dc.vp_end[0] = c.viewport_end[0];
    dc.vp_end[1] = c.viewport_end[1];
}
// compute_num_channels, src/channels.cpp:42-113
inline int host_compute_num_channels(const int* channels, int n, int max_generic) {
    int total = 0;
    for (int i = 0; i < n; i++) {
        switch (channels[i]) {
            case RB_CH_RADIANCE: case RB_CH_POSITION: case RB_CH_GEOMETRY_NORMAL: case RB_CH_SHADING_NORMAL:
            case RB_CH_DIFFUSE_REFLECTANCE: case RB_CH_SPECULAR_REFLECTANCE: case RB_CH_VERTEX_COLOR:
                total += 3;
                break;
            case RB_CH_ALPHA: case RB_CH_DEPTH: case RB_CH_ROUGHNESS: case RB_CH_SHAPE_ID: case RB_CH_TRIANGLE_ID: case RB_CH_MATERIAL_ID:
                total += 1;
                break;
            case RB_CH_UV: case RB_CH_BARYCENTRIC:
                total += 2;
                break;
            case RB_CH_GENERIC_TEXTURE:
                total += max_generic;
                break;
            default:
                return -1;
        }
    }
    return total;
}
