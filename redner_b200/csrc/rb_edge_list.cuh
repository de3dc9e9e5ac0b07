// Edge list of a scene (topology: unique mesh edges with their one or two faces, seam repair, flat-edge filter) as data-parallel
// element steps.  What the reference does per shape in src/edge.cpp:233-296 (collect 3 half-edges per triangle :236-259, sort and
// merge :261-273, sort by end-point positions and pair seam twins :276-288, drop coplanar pairs :291-296), here over ALL shapes of a
// scene at once:
//
//   A  el_half_edge_key     one 64-bit key per half-edge: (global index of the smaller end point, of the larger one); a STABLE sort of
//                           (key, half-edge index) groups the copies of an edge, shapes in order, faces in ascending order inside a run
//   B  el_merge_run         the first element of every run of equal keys becomes an edge: f0 = its face, f1 = face of the run's LAST
//                           element when the run has two or more (src/edge.cpp:86-90)
//   C  ELPositionLess       order of the merged edges by (shape, smaller end-point POSITION, larger one).  The reference's comparator
//                           answers true for equal keys (src/edge.cpp:93-131); under its merge sort (insertion sort of <= 32 elements,
//                           then merges that take from the right run on ties) equal keys come out in exactly the REVERSE of their input
//                           order -- rb_scene_host.hpp restates that sort step by step, tests/test_edge_list_cpu.py checks the
//                           equivalence -- so: reverse the list, then sort STABLY with the strict order
//   D  el_pair_seam         an edge with one face whose sorted neighbour has the same two end-point positions takes that neighbour's
//                           face as f1; the successor wins over the predecessor (src/edge.cpp:133-166)
//   E  edge_is_flat         (rb_edge.cuh) flags the edges to drop; an exclusive scan compacts the rest in order
//
// The element functions are RB_HD: rb_edge_list.cu runs them in kernels between CUB sorts / scans, tests/edge_list_check.cpp runs the
// same functions in serial loops between std::stable_sort calls and compares the result with host_build_edges (rb_scene_host.hpp).
#pragma once
#include "rb_edge.cuh"

struct ELScene {
    const rb_shape* shapes; // pointers valid in the calling address space
    const int* tri_off;     // [S + 1] prefix sums of num_triangles
    const int* vert_off;    // [S + 1] prefix sums of num_vertices
    int S;
    int key_bits;           // bits of a global vertex index: key = (lo << key_bits) | hi
};

// shape that owns global triangle `gt`: the last s with tri_off[s] <= gt (empty shapes are skipped by the search)
RB_HD int el_shape_of_triangle(const ELScene& L, int gt) {
    int lo = 0, hi = L.S; // invariant: tri_off[lo] <= gt < tri_off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (L.tri_off[mid] <= gt) lo = mid;
        else hi = mid;
    }
    return lo;
}

// A: half-edge h = 3 * (global triangle) + k joins corners k and (k + 1) % 3
RB_HD unsigned long long el_half_edge_key(const ELScene& L, int h) {
    const int gt = h / 3, k = h - 3 * gt;
    const int s = el_shape_of_triangle(L, gt);
    int idx[3];
    shape_tri(L.shapes[s], gt - L.tri_off[s], idx);
    const int a = idx[k], b = idx[k == 2 ? 0 : k + 1];
    const unsigned long long lo = (unsigned long long)(L.vert_off[s] + (a < b ? a : b)), hi = (unsigned long long)(L.vert_off[s] + (a < b ? b : a));
    return (lo << L.key_bits) | hi;
}

RB_HD bool el_is_run_head(const unsigned long long* keys, int i) { return i == 0 || keys[i] != keys[i - 1]; }

// B: called for run heads only
RB_HD Edge el_merge_run(const ELScene& L, const unsigned long long* keys, const int* half_edges, int n, int i) {
    const unsigned long long key = keys[i];
    int j = i;
    while (j + 1 < n && keys[j + 1] == key) j++;
    const int gt = half_edges[i] / 3;
    const int s = el_shape_of_triangle(L, gt);
    Edge e;
    e.shape_id = s;
    e.v0 = (int)(key >> L.key_bits) - L.vert_off[s];
    e.v1 = (int)(key & ((1ULL << L.key_bits) - 1ULL)) - L.vert_off[s];
    e.f0 = gt - L.tri_off[s];
    e.f1 = j > i ? half_edges[j] / 3 - L.tri_off[s] : -1;
    return e;
}

RB_HD bool el_position_less(V3 a, V3 b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z < b.z;
}
// end points of an edge ordered by position
RB_HD void el_end_points(const rb_shape* shapes, const Edge& e, V3& lo, V3& hi) {
    lo = edge_v0(shapes, e);
    hi = edge_v1(shapes, e);
    if (el_position_less(hi, lo)) {
        V3 t = lo;
        lo = hi;
        hi = t;
    }
}
// C: strict order on indices into the merged list
struct ELPositionLess {
    const rb_shape* shapes;
    const Edge* merged;
    RB_HD bool operator()(int x, int y) const {
        const Edge a = merged[x], b = merged[y];
        if (a.shape_id != b.shape_id) return a.shape_id < b.shape_id;
        V3 al, ah, bl, bh;
        el_end_points(shapes, a, al, ah);
        el_end_points(shapes, b, bl, bh);
        if (!same_pos(al, bl)) return el_position_less(al, bl);
        if (!same_pos(ah, bh)) return el_position_less(ah, bh);
        return false;
    }
};

RB_HD bool el_same_segment(const rb_shape* shapes, const Edge& a, const Edge& b) {
    if (a.shape_id != b.shape_id) return false;
    V3 al, ah, bl, bh;
    el_end_points(shapes, a, al, ah);
    el_end_points(shapes, b, bl, bh);
    return same_pos(al, bl) && same_pos(ah, bh);
}
// D: the edge at position p of the position order, with its seam twin's face if it has none of its own
RB_HD Edge el_pair_seam(const rb_shape* shapes, const Edge* merged, const int* order, int M, int p) {
    Edge e = merged[order[p]];
    if (e.f1 != -1) return e;
    int f1 = -1;
    if (p > 0) {
        const Edge c = merged[order[p - 1]];
        if (el_same_segment(shapes, e, c)) f1 = c.f0;
    }
    if (p + 1 < M) {
        const Edge c = merged[order[p + 1]];
        if (el_same_segment(shapes, e, c)) f1 = c.f0;
    }
    e.f1 = f1;
    return e;
}

RB_HD int el_bits_for(long long count) { // smallest b with count <= 2^b (at least 1)
    int b = 1;
    while ((1LL << b) < count) b++;
    return b;
}
