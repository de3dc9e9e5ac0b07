// Mesh-edge helpers shared by the scene build (host) and the edge-sampling kernels (device).
//   Edge, get_v0/v1, get_non_shared_v0/v1, is_silhouette     src/edge.h:13-204
//   clip_line (Cohen-Sutherland against the unit square)      src/line_clip.h:37-102
// All functions take an array of rb_shape whose pointers are valid in the calling address space
// (device pointers inside kernels, host mirrors during rb_scene_create).
#pragma once
#include "rb_camera.cuh"
#include "rb_shape.cuh"

RB_HD V3 edge_v0(const rb_shape* shapes, const Edge& e) { return shape_vertex(shapes[e.shape_id], e.v0); }
RB_HD V3 edge_v1(const rb_shape* shapes, const Edge& e) { return shape_vertex(shapes[e.shape_id], e.v1); }
RB_HD bool same_pos(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// third vertex of face f0 (by index), src/edge.h:84-94
RB_HD V3 edge_opposite0(const rb_shape* shapes, const Edge& e) {
    int idx[3];
    shape_tri(shapes[e.shape_id], e.f0, idx);
    for (int i = 0; i < 3; i++)
        if (idx[i] != e.v0 && idx[i] != e.v1) return shape_vertex(shapes[e.shape_id], idx[i]);
    return edge_v0(shapes, e);
}
// third vertex of face f1 (by POSITION, because f1 may come from a duplicated seam edge), src/edge.h:96-121
RB_HD V3 edge_opposite1(const rb_shape* shapes, const Edge& e) {
    int idx[3];
    shape_tri(shapes[e.shape_id], e.f1, idx);
    V3 a = edge_v0(shapes, e), b = edge_v1(shapes, e);
    for (int i = 0; i < 3; i++) {
        V3 v = shape_vertex(shapes[e.shape_id], idx[i]);
        if (!same_pos(v, a) && !same_pos(v, b)) return v;
    }
    return b;
}
RB_HD bool edge_is_silhouette(const rb_shape* shapes, V3 p, const Edge& e) {
    V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
    if (e.f0 == -1 || e.f1 == -1) {
        if (e.f0 != -1) {
            V3 o = edge_opposite0(shapes, e);
            if (length_sq(cross(v0 - o, v1 - o)) < Real(1e-20)) return false;
        }
        if (e.f1 != -1) {
            V3 o = edge_opposite1(shapes, e);
            if (length_sq(cross(v1 - o, v0 - o)) < Real(1e-20)) return false;
        }
        return true;
    }
    V3 o0 = edge_opposite0(shapes, e), o1 = edge_opposite1(shapes, e);
    V3 n0 = cross(v0 - o0, v1 - o0), n1 = cross(v1 - o1, v0 - o1);
    Real l0 = length_sq(n0), l1 = length_sq(n1);
    if (l0 < Real(1e-20) || l1 < Real(1e-20)) return false;
    n0 = n0 / sqrt(l0);
    n1 = n1 / sqrt(l1);
    if (shapes[e.shape_id].normals == nullptr) {
        // without interpolated normals every non-flat edge can be a silhouette
        return !(dot(n0, n1) >= 1 - Real(1e-6));
    }
    bool f0 = dot(p - o0, n0) > 0, f1 = dot(p - o1, n1) > 0;
    return (f0 && !f1) || (!f0 && f1);
}
// dihedral filter used when the edge list is built, src/edge.cpp:168-184
RB_HD bool edge_is_flat(const rb_shape* shapes, const Edge& e) {
    if (e.f0 == -1 || e.f1 == -1) return false;
    V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
    V3 o0 = edge_opposite0(shapes, e), o1 = edge_opposite1(shapes, e);
    V3 n0 = normalize(cross(v0 - o0, v1 - o0)), n1 = normalize(cross(v1 - o1, v0 - o1));
    return dot(n0, n1) >= (1 - Real(1e-6));
}

RB_HD int clip_code(V2 v) {
    int c = 0;
    if (v.x < 0) c |= 1; else if (v.x > 1) c |= 2;
    if (v.y < 0) c |= 4; else if (v.y > 1) c |= 8;
    return c;
}
RB_HD bool clip_line_unit(V2 v0, V2 v1, V2& a, V2& b) {
    int c0 = clip_code(v0), c1 = clip_code(v1);
    a = v0;
    b = v1;
    for (int it = 0; it < 16; it++) {
        if (!(c0 | c1)) return true;
        if (c0 & c1) return false;
        int co = c0 ? c0 : c1;
        V2 v = zero2();
        if (co & 8) {
            v.x = a.x + (b.x - a.x) * (1 - a.y) / (b.y - a.y);
            v.y = 1;
        } else if (co & 4) {
            v.x = a.x + (b.x - a.x) * (0 - a.y) / (b.y - a.y);
            v.y = 0;
        } else if (co & 2) {
            v.y = a.y + (b.y - a.y) * (1 - a.x) / (b.x - a.x);
            v.x = 1;
        } else if (co & 1) {
            v.y = a.y + (b.y - a.y) * (0 - a.x) / (b.x - a.x);
            v.x = 0;
        }
        if (co == c0) {
            a = v;
            c0 = clip_code(a);
        } else {
            b = v;
            c1 = clip_code(b);
        }
    }
    return false;
}
