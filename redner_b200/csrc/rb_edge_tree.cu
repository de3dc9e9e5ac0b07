// Secondary-edge hierarchy on the GPU (reference: EdgeTree::EdgeTree, src/edge_tree.cpp:724-882, Thrust-parallel there too).
//
// Builds, per rb_scene_create, the two trees the boundary sampler walks -- a 3-D tree over the camera-silhouette edges and a
// 6-D (position x Hough) tree over the rest -- entirely on the device and bit-for-bit like the host builder of
// rb_scene_host.hpp (HostTreeBuilder; tests/test_scene_build_gpu.py compares the two node arrays):
//   k_et_leaves     edge -> leaf record (position box, Hough box of the two face planes, length x exterior angle), camera-silhouette
//                   flag, sums for the mean end point                                     src/edge_tree.cpp:23-66, :749-756
//   k_et_mad        mean absolute deviation of the end points -> billboard size           :763-773
//   k_et_bounds     per-tree bounds of the leaf centres (ordered-integer atomic min / max)
//   k_et_codes      63-bit / 60-bit Morton codes, tree bit on top                         :166-266
//   radix sort      stable, by (tree, code); ties keep edge order, which the Karras split resolves by index like the reference
//   k_et_karras     radix tree, one thread per inner node                                 :282-376
//   k_et_climb<0>   bottom-up boxes / weighted lengths, atomic arrival counters           :391-445
//   k_et_climb<1>   treelet (<= 7 leaves) SAH re-optimisation, bottom-up, Karras & Aila 2013 Algorithm 2 with the reference's
//                   quirks (subset area always includes leaf 0, :491-500)                 :464-711
//   k_et_sizes / k_et_rank / k_et_emit   depth-first numbering of the inner nodes and the 128-byte EdgeNode records (both
//                   children's bounds per record) that rb_secondary.cuh walks
// Everything that decides the shape of the tree is computed in double like the reference (areas, costs, Morton quantisation).
// This translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true (redner_b200/build.py): the float parts (face
// normals) must round like the host builder's, and -Xptxas -dlcm=cg keeps the bottom-up passes' loads out of the (incoherent) L1.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "rb_edge.cuh"
#include "rb_scene.cuh"

struct ETNode { // node of the reference-shaped tree (double precision like the reference's Real); 128 bytes
    double pmin[3], pmax[3], dmin[3], dmax[3];
    double wlen, cost;
    int parent, child[2], edge_id;
};
struct ETGlobals {
    double sum[3];            // sum of all end points
    double mad[3];            // sum of |end point - mean|
    unsigned long long lo[2][6], hi[2][6]; // per tree: ordered-integer bounds of the leaf centres (position, Hough)
    int n_cs;                 // camera-silhouette edges
};

__device__ __forceinline__ unsigned long long d2ord(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord2d(unsigned long long o) {
    unsigned long long b = (o & 0x8000000000000000ULL) ? (o & 0x7fffffffffffffffULL) : ~o;
    return __longlong_as_double((long long)b);
}
// std::min / std::max semantics: the FIRST argument wins ties -- fmin / fmax order -0 below +0, and the Hough bounds are full of
// signed zeros (axis-aligned faces); the records must equal the host builder's bit for bit
__device__ __forceinline__ double et_min(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double et_max(double a, double b) { return a < b ? b : a; }
__device__ V3 et_edge_normal(const rb_shape* shapes, const Edge& e, int which) { // host_edge_normal
    V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
    V3 n;
    if (which == 0) {
        V3 o = edge_opposite0(shapes, e);
        n = cross(v0 - o, v1 - o);
    } else {
        V3 o = edge_opposite1(shapes, e);
        n = cross(v1 - o, v0 - o);
    }
    Real l2 = length_sq(n);
    if (l2 < Real(1e-20)) return zero3();
    return n / sqrt(l2);
}

__global__ void k_et_leaves(const rb_shape* shapes, const Edge* edges, int E, double cx, double cy, double cz, ETNode* leaves, unsigned char* is_cs,
                            ETGlobals* g) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s[3] = {0, 0, 0};
    int cs = 0;
    if (i < E) {
        const Edge e = edges[i];
        const double co[3] = {cx, cy, cz};
        V3 cam_org = mk3((Real)cx, (Real)cy, (Real)cz);
        V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
        ETNode n;
        V3 n0 = et_edge_normal(shapes, e, 0);
        V3 n1 = e.f1 == -1 ? -n0 : et_edge_normal(shapes, e, 1);
        double p[3], p0d = 0, p1d = 0;
        for (int k = 0; k < 3; k++) {
            s[k] = (double)v0[k] + (double)v1[k];
            p[k] = 0.5 * ((double)v0[k] + (double)v1[k]) - co[k];
        }
        for (int k = 0; k < 3; k++) {
            p0d += p[k] * (double)n0[k];
            p1d += p[k] * (double)n1[k];
        }
        for (int k = 0; k < 3; k++) {
            double h0 = (double)n0[k] * p0d, h1 = (double)n1[k] * p1d;
            n.pmin[k] = et_min((double)v0[k], (double)v1[k]);
            n.pmax[k] = et_max((double)v0[k], (double)v1[k]);
            n.dmin[k] = et_min(h0, h1);
            n.dmax[k] = et_max(h0, h1);
        }
        double ext = M_PI;
        if (e.f1 != -1) ext = acos(et_min(1.0, et_max(-1.0, (double)dot(n0, n1))));
        n.wlen = (double)length(v1 - v0) * ext;
        n.parent = -1;
        n.child[0] = n.child[1] = -1;
        n.edge_id = i;
        n.cost = 0;
        leaves[i] = n;
        cs = edge_is_silhouette(shapes, cam_org, e) ? 1 : 0;
        is_cs[i] = (unsigned char)cs;
    }
    // block sums (order of the additions differs from the host's sequential loop: `expand` agrees to ~1e-15 relative)
    __shared__ double sh[3][256];
    __shared__ int shc[256];
    for (int k = 0; k < 3; k++) sh[k][threadIdx.x] = s[k];
    shc[threadIdx.x] = cs;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
            shc[threadIdx.x] += shc[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; k++) atomicAdd(&g->sum[k], sh[k][0]);
        atomicAdd(&g->n_cs, shc[0]);
    }
}
__global__ void k_et_mad(const rb_shape* shapes, const Edge* edges, int E, const ETNode* leaves, const unsigned char* is_cs, ETGlobals* g) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s[3] = {0, 0, 0};
    if (i < E) {
        const Edge e = edges[i];
        V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
        for (int k = 0; k < 3; k++) {
            double mean = g->sum[k] / (2.0 * E);
            s[k] = fabs((double)v0[k] - mean) + fabs((double)v1[k] - mean);
        }
        // bounds of the leaf centres of this edge's tree
        const ETNode& n = leaves[i];
        int t = is_cs[i] ? 0 : 1;
        for (int k = 0; k < 3; k++) {
            atomicMin(&g->lo[t][k], d2ord(n.pmin[k]));
            atomicMax(&g->hi[t][k], d2ord(n.pmax[k]));
            atomicMin(&g->lo[t][3 + k], d2ord(n.dmin[k]));
            atomicMax(&g->hi[t][3 + k], d2ord(n.dmax[k]));
        }
    }
    __shared__ double sh[3][256];
    for (int k = 0; k < 3; k++) sh[k][threadIdx.x] = s[k];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off)
            for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; k++) atomicAdd(&g->mad[k], sh[k][0]);
}
__device__ __forceinline__ unsigned long long et_expand21(unsigned long long x) {
    x &= 0x1fffffULL;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
__device__ __forceinline__ unsigned long long et_expand10(unsigned long long x) {
    unsigned long long r = 0;
    for (int b = 0; b < 10; b++) r |= ((x >> b) & 1ULL) << (5 * b);
    return r;
}
// key = tree bit (0: camera silhouettes, 1: the rest) on top of the Morton code of the leaf centre inside its tree's bounds
__global__ void k_et_codes(int E, const ETNode* leaves, const unsigned char* is_cs, const ETGlobals* g, unsigned long long* keys, int* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const ETNode& n = leaves[i];
    int t = is_cs[i] ? 0 : 1;
    double q[6];
    for (int k = 0; k < 3; k++) {
        double lo = ord2d(g->lo[t][k]), hi = ord2d(g->hi[t][k]), lod = ord2d(g->lo[t][3 + k]), hid = ord2d(g->hi[t][3 + k]);
        double cp = 0.5 * (n.pmin[k] + n.pmax[k]), cd = 0.5 * (n.dmin[k] + n.dmax[k]);
        q[k] = hi - lo <= 0 ? 0.5 : (cp - lo) / (hi - lo);
        q[3 + k] = hid - lod <= 0 ? 0.5 : (cd - lod) / (hid - lod);
    }
    unsigned long long c;
    if (t == 0) {
        double sc = (1 << 21) - 1;
        c = (et_expand21((unsigned long long)(q[0] * sc)) << 2) | (et_expand21((unsigned long long)(q[1] * sc)) << 1) | et_expand21((unsigned long long)(q[2] * sc));
    } else {
        c = 0;
        for (int k = 0; k < 6; k++) c |= et_expand10((unsigned long long)(q[k] * 1023)) << (5 - k);
        c |= 1ULL << 63;
    }
    keys[i] = c;
    vals[i] = i;
}

// Tree t lives in nodes[base .. base + LB + L): inner nodes first (LB = max(L - 1, 1)), then the leaves in sorted order.
struct ETTree {
    int first; // first sorted position of this tree's edges
    int L;     // leaves
    int base;  // first node
    int six;
};
__device__ __forceinline__ int et_lb(const ETTree& t) { return t.L - 1 > 1 ? t.L - 1 : 1; }

__global__ void k_et_init(ETTree t, const ETNode* leaves, const int* ids_sorted, ETNode* nodes) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int LB = et_lb(t);
    if (j < LB) {
        ETNode x;
        for (int k = 0; k < 3; k++) {
            x.pmin[k] = x.dmin[k] = INFINITY;
            x.pmax[k] = x.dmax[k] = -INFINITY;
        }
        x.wlen = 0;
        x.cost = 0;
        x.parent = -1;
        x.child[0] = x.child[1] = -1;
        x.edge_id = -1;
        nodes[t.base + j] = x;
    }
    if (j < t.L) {
        ETNode x = leaves[ids_sorted[t.first + j]];
        x.parent = -1;
        double dx = x.pmax[0] - x.pmin[0], dy = x.pmax[1] - x.pmin[1], dz = x.pmax[2] - x.pmin[2];
        double s = dx * dy + dx * dz + dy * dz;
        if (t.six) {
            double ex = x.dmax[0] - x.dmin[0], ey = x.dmax[1] - x.dmin[1], ez = x.dmax[2] - x.dmin[2];
            s += ex * ey + ex * ez + ey * ez;
        }
        x.cost = 2 * s;
        nodes[t.base + LB + j] = x;
        if (t.L == 1) nodes[t.base] = x; // src/edge_tree.cpp:303-308
    }
}
__device__ __forceinline__ int et_lcp(const ETTree& t, const unsigned long long* keys, const int* ids, int i, int j) {
    if (i < 0 || i >= t.L || j < 0 || j >= t.L) return -1;
    unsigned long long a = keys[t.first + i] & 0x7fffffffffffffffULL, b = keys[t.first + j] & 0x7fffffffffffffffULL;
    if (a == b) return __clzll((long long)(a ^ b)) + __clzll((long long)((unsigned long long)ids[t.first + i] ^ (unsigned long long)ids[t.first + j]));
    return __clzll((long long)(a ^ b));
}
__global__ void k_et_karras(ETTree t, const unsigned long long* keys, const int* ids, ETNode* nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.L - 1) return;
    const int LB = et_lb(t);
    int d = (et_lcp(t, keys, ids, i, i + 1) - et_lcp(t, keys, ids, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = et_lcp(t, keys, ids, i, i - d);
    int lmax = 2;
    while (et_lcp(t, keys, ids, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int s = lmax / 2; s >= 1; s /= 2)
        if (et_lcp(t, keys, ids, i, i + (l + s) * d) > dmin) l += s;
    int j = i + l * d;
    int dnode = et_lcp(t, keys, ids, i, j);
    int s = 0, div = 2;
    for (int q = (l + (div - 1)) / div; q >= 1;) {
        if (et_lcp(t, keys, ids, i, i + (s + q) * d) > dnode) s += q;
        if (q == 1) break;
        div *= 2;
        q = (l + (div - 1)) / div;
    }
    int gamma = i + s * d + (d < 0 ? d : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    int c0 = (lo == gamma) ? LB + gamma : gamma;
    int c1 = (hi == gamma + 1) ? LB + gamma + 1 : gamma + 1;
    nodes[t.base + i].child[0] = t.base + c0;
    nodes[t.base + i].child[1] = t.base + c1;
    nodes[t.base + c0].parent = t.base + i;
    nodes[t.base + c1].parent = t.base + i;
}

// ---- node algebra of HostTreeBuilder, on global indices ----
struct ETOps {
    ETNode* n;
    int six;
    __device__ static void merge_into(ETNode& o, const ETNode& a, const ETNode& b) {
        for (int k = 0; k < 3; k++) {
            o.pmin[k] = et_min(a.pmin[k], b.pmin[k]);
            o.pmax[k] = et_max(a.pmax[k], b.pmax[k]);
            o.dmin[k] = et_min(a.dmin[k], b.dmin[k]);
            o.dmax[k] = et_max(a.dmax[k], b.dmax[k]);
        }
    }
    __device__ double area(const ETNode& a) const {
        double dx = a.pmax[0] - a.pmin[0], dy = a.pmax[1] - a.pmin[1], dz = a.pmax[2] - a.pmin[2];
        double s = dx * dy + dx * dz + dy * dz;
        if (six) {
            double ex = a.dmax[0] - a.dmin[0], ey = a.dmax[1] - a.dmin[1], ez = a.dmax[2] - a.dmin[2];
            s += ex * ey + ex * ez + ey * ez;
        }
        return 2 * s;
    }
    __device__ void refresh(int i) { // bounds, weighted length and SAH cost of an inner node from its children
        ETNode o = n[i];
        const ETNode a = n[o.child[0]], b = n[o.child[1]];
        merge_into(o, a, b);
        o.wlen = a.wlen + b.wlen;
        o.cost = area(o) + a.cost + b.cost;
        n[i] = o;
    }
    __device__ void propagate_cost(int root, const int* lv, int cnt) { // src/edge_tree.cpp:546-579
        for (int i = 0; i < cnt; i++) {
            int cur = lv[i];
            while (cur != root) {
                if (n[cur].cost < 0) {
                    if (n[n[cur].child[0]].cost >= 0 && n[n[cur].child[1]].cost >= 0) refresh(cur);
                    else break;
                }
                cur = n[cur].parent;
            }
        }
        refresh(root);
    }
    __device__ void restruct(int parent, int child_index, const int* lv, const int* inner, unsigned char partition, const unsigned char* optimal, int& index,
                             int cnt) { // src/edge_tree.cpp:586-626
        unsigned char st_part[8], st_child[8];
        int st_parent[8];
        int sp = 0;
        st_part[sp] = partition;
        st_child[sp] = (unsigned char)child_index;
        st_parent[sp] = parent;
        sp++;
        while (sp > 0) {
            sp--;
            unsigned char part = st_part[sp], ch = st_child[sp];
            int par = st_parent[sp];
            if (__popc((unsigned)part) == 1) {
                int leaf = lv[__ffs((unsigned)part) - 1];
                n[par].child[ch] = leaf;
                n[leaf].parent = par;
            } else {
                int node = inner[index++];
                n[node].cost = -1;
                n[par].child[ch] = node;
                n[node].parent = par;
                unsigned char lp = optimal[part];
                unsigned char rp = (unsigned char)((~lp) & part);
                st_part[sp] = lp;
                st_child[sp] = 0;
                st_parent[sp] = node;
                sp++;
                st_part[sp] = rp;
                st_child[sp] = 1;
                st_parent[sp] = node;
                sp++;
            }
        }
        propagate_cost(parent, lv, cnt);
    }
    // One WARP optimises one treelet: lane 0 forms it and rewires the tree afterwards; the 127 subset areas and, round by round, the
    // optimal partition of every subset of k leaves (Karras & Aila 2013, Algorithm 2, src/edge_tree.cpp:502-544) are spread over the
    // lanes.  Each subset is still evaluated by ONE lane in the reference's order, so ties resolve identically.  (A single thread
    // per treelet -- the host builder's shape -- costs ~0.2 ms per node here and the pass climbs ~30 levels.)
    struct Scratch { // shared memory, per warp
        double a[128], c_opt[128], bx[7][12];
        unsigned char optimal[128];
        int lv[7], inner[5], cnt;
    };
    __device__ void treelet_optimize_warp(int root, Scratch& w) {
        const int lane = threadIdx.x & 31;
        if (n[root].edge_id != -1) return; // (warp-uniform)
        if (lane == 0) { // src/edge_tree.cpp:627-684
            int cnt = 0, icnt = 0;
            w.lv[cnt++] = n[root].child[0];
            w.lv[cnt++] = n[root].child[1];
            int max_idx = 0;
            while (cnt < 7 && max_idx != -1) {
                max_idx = -1;
                double max_area = -1;
                for (int i = 0; i < cnt; i++)
                    if (n[w.lv[i]].edge_id == -1) {
                        double ar = area(n[w.lv[i]]);
                        if (ar > max_area) {
                            max_area = ar;
                            max_idx = i;
                        }
                    }
                if (max_idx != -1) {
                    int tmp = w.lv[max_idx];
                    w.inner[icnt++] = tmp;
                    w.lv[max_idx] = w.lv[cnt - 1];
                    w.lv[cnt - 1] = n[tmp].child[0];
                    w.lv[cnt] = n[tmp].child[1];
                    cnt++;
                }
            }
            w.cnt = cnt;
            for (int i = 0; i < cnt; i++) {
                const ETNode& l = n[w.lv[i]];
                for (int k = 0; k < 3; k++) {
                    w.bx[i][k] = l.pmin[k];
                    w.bx[i][3 + k] = l.pmax[k];
                    w.bx[i][6 + k] = l.dmin[k];
                    w.bx[i][9 + k] = l.dmax[k];
                }
                w.c_opt[1u << i] = l.cost;
            }
        }
        __syncwarp();
        const int cnt = w.cnt;
        const unsigned num_subsets = (1u << cnt) - 1;
        // a[s] = area(union(leaf 0, leaves of s)) -- the reference's union always starts from leaf 0 (src/edge_tree.cpp:491-500)
        for (unsigned s = 1 + lane; s <= num_subsets; s += 32) {
            ETNode t;
            for (int k = 0; k < 3; k++) {
                t.pmin[k] = w.bx[0][k];
                t.pmax[k] = w.bx[0][3 + k];
                t.dmin[k] = w.bx[0][6 + k];
                t.dmax[k] = w.bx[0][9 + k];
            }
            for (int i = 1; i < cnt; i++)
                if ((s >> i) & 1u)
                    for (int k = 0; k < 3; k++) {
                        t.pmin[k] = et_min(t.pmin[k], w.bx[i][k]);
                        t.pmax[k] = et_max(t.pmax[k], w.bx[i][3 + k]);
                        t.dmin[k] = et_min(t.dmin[k], w.bx[i][6 + k]);
                        t.dmax[k] = et_max(t.dmax[k], w.bx[i][9 + k]);
                    }
            w.a[s] = area(t);
        }
        __syncwarp();
        for (int k = 2; k <= cnt; k++) {
            for (unsigned s = 1 + lane; s <= num_subsets; s += 32)
                if (__popc(s) == k) {
                    double c_s = INFINITY;
                    unsigned p_s = 0;
                    unsigned d = (s - 1u) & s;
                    unsigned p = (0u - d) & s;
                    do {
                        double c = w.c_opt[p] + w.c_opt[s ^ p];
                        if (c < c_s) {
                            c_s = c;
                            p_s = p;
                        }
                        p = (p - d) & s;
                    } while (p != 0);
                    w.c_opt[s] = w.a[s] + c_s;
                    w.optimal[s] = (unsigned char)p_s;
                }
            __syncwarp();
        }
        if (lane == 0) {
            unsigned char mask = (unsigned char)((1u << cnt) - 1);
            int index = 0;
            unsigned char left = w.optimal[mask];
            restruct(root, 0, w.lv, w.inner, left, w.optimal, index, cnt);
            unsigned char right = (unsigned char)((~left) & mask);
            restruct(root, 1, w.lv, w.inner, right, w.optimal, index, cnt);
            refresh(root);
        }
        __syncwarp();
    }
};
// Bottom-up pass: every thread starts at a leaf and climbs; the SECOND thread to arrive at a node processes it (its two subtrees are
// finished then) -- boxes / weighted lengths (PASS 0), treelet optimisation (PASS 1, src/edge_tree.cpp:685-707) or inner-node counts
// of the subtrees (PASS 2).  Concurrently processed nodes sit in disjoint subtrees.
template <int PASS>
__global__ void k_et_climb(ETTree t, ETNode* nodes, int* arrived, int* inner_count) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= t.L || t.L < 2) return;
    int cur = nodes[t.base + et_lb(t) + j].parent;
    while (cur != -1) {
        __threadfence();
        if (atomicAdd(&arrived[cur], 1) == 0) break; // first arrival: the sibling subtree is not done
        __threadfence();
        if (PASS == 0) {
            ETNode o = nodes[cur];
            const ETNode a = nodes[o.child[0]], b = nodes[o.child[1]];
            ETOps::merge_into(o, a, b);
            o.wlen = a.wlen + b.wlen;
            nodes[cur] = o;
        } else {
            int c0 = nodes[cur].child[0], c1 = nodes[cur].child[1];
            inner_count[cur] = 1 + (nodes[c0].edge_id == -1 ? inner_count[c0] : 0) + (nodes[c1].edge_id == -1 ? inner_count[c1] : 0);
        }
        cur = nodes[cur].parent;
    }
}
// The treelet pass (src/edge_tree.cpp:685-707): one warp per leaf climbs; lane 0 owns the arrival counters.
#define RB_ET_WARPS 4
__global__ void __launch_bounds__(32 * RB_ET_WARPS) k_et_optimize(ETTree t, ETNode* nodes, int* arrived) {
    __shared__ ETOps::Scratch scratch[RB_ET_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int j = blockIdx.x * RB_ET_WARPS + warp;
    if (j >= t.L || t.L < 2) return;
    ETOps ops{nodes, t.six};
    int cur = nodes[t.base + et_lb(t) + j].parent;
    while (cur != -1) {
        int go = 0;
        if (lane == 0) {
            __threadfence();
            go = atomicAdd(&arrived[cur], 1) != 0;
            __threadfence();
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) break; // first arrival: the sibling subtree is not done
        ops.treelet_optimize_warp(cur, scratch[warp]);
        __threadfence();
        cur = nodes[cur].parent;
    }
}
// depth-first (left first) number of every inner node: the number of inner nodes visited before it
__global__ void k_et_rank(ETTree t, const ETNode* nodes, const int* inner_count, int rank_base, int* rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.L - 1) return;
    int node = t.base + i, r = 0;
    int cur = node;
    while (nodes[cur].parent != -1) {
        int par = nodes[cur].parent;
        r += 1; // the parent itself
        if (nodes[par].child[1] == cur) {
            int sib = nodes[par].child[0];
            if (nodes[sib].edge_id == -1) r += inner_count[sib];
        }
        cur = par;
    }
    rank[node] = rank_base + r;
}
__global__ void k_et_emit(ETTree t, const ETNode* nodes, const int* rank, EdgeNode* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.L - 1) return;
    int node = t.base + i;
    EdgeNode en;
    memset(&en, 0, sizeof(en));
    for (int c = 0; c < 2; c++) {
        int ci = nodes[node].child[c];
        const ETNode& h = nodes[ci];
        for (int k = 0; k < 3; k++) {
            en.c[c].pmin[k] = (float)h.pmin[k];
            en.c[c].pmax[k] = (float)h.pmax[k];
            en.c[c].dmin[k] = (float)h.dmin[k];
            en.c[c].dmax[k] = (float)h.dmax[k];
        }
        en.c[c].wlen = (float)h.wlen;
        en.c[c].ref = h.edge_id != -1 ? ~h.edge_id : rank[ci];
    }
    out[rank[node]] = en;
}

// Builds both trees for the scene's current edge list and camera.  Device temporaries are released in stream order.
int rb_build_edge_trees_gpu(rb_scene* sc, cudaStream_t stream) {
    const int E = sc->dev.num_edges;
    sc->dev.edge_nodes = nullptr;
    sc->num_edge_nodes = 0;
    sc->dev.edge_root_cs = sc->dev.edge_root_ncs = RB_EDGE_EMPTY;
    sc->dev.edge_bounds_expand = 0.f;
    if (E == 0) return 0;
    const double iw = 1.0 / sc->dev.cam.c2w[15];
    const double co[3] = {sc->dev.cam.c2w[3] * iw, sc->dev.cam.c2w[7] * iw, sc->dev.cam.c2w[11] * iw};
    std::vector<void*> temps;
    auto talloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(bytes, 16), stream) != cudaSuccess) return nullptr;
        temps.push_back(p);
        return p;
    };
    auto release = [&]() {
        for (void* p : temps) cudaFreeAsync(p, stream);
    };
    ETNode* leaves = (ETNode*)talloc(sizeof(ETNode) * (size_t)E);
    ETNode* nodes = (ETNode*)talloc(sizeof(ETNode) * (2 * (size_t)E + 2));
    unsigned char* is_cs = (unsigned char*)talloc(E);
    ETGlobals* g = (ETGlobals*)talloc(sizeof(ETGlobals));
    unsigned long long *keys = (unsigned long long*)talloc(8 * (size_t)E), *keys_sorted = (unsigned long long*)talloc(8 * (size_t)E);
    int *vals = (int*)talloc(4 * (size_t)E), *ids = (int*)talloc(4 * (size_t)E);
    int* counters = (int*)talloc(4 * 3 * (2 * (size_t)E + 2)); // arrival flags of the passes, inner counts / ranks
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, keys_sorted, vals, ids, E, 0, 64, stream);
    void* sort_tmp = talloc(sort_bytes);
    EdgeNode* out = sc->edge_nodes_buf; // (reused when rb_scene_set_camera rebuilds the trees: the edge list does not change)
    if (!leaves || !nodes || !is_cs || !g || !keys || !keys_sorted || !vals || !ids || !counters || !sort_tmp ||
        (out == nullptr && cudaMallocAsync((void**)&out, sizeof(EdgeNode) * (size_t)std::max(E, 1), stream) != cudaSuccess)) {
        release();
        rb_set_error("rb_scene_create: out of device memory for the edge trees");
        return 1;
    }
    if (sc->edge_nodes_buf == nullptr) {
        sc->allocs.push_back(out);
        sc->edge_nodes_buf = out;
    }
    ETGlobals g0;
    memset(&g0, 0, sizeof(g0));
    for (int t = 0; t < 2; t++)
        for (int k = 0; k < 6; k++) {
            g0.lo[t][k] = ~0ULL;
            g0.hi[t][k] = 0ULL;
        }
    RB_CUDA_OK(cudaMemcpyAsync(g, &g0, sizeof(g0), cudaMemcpyHostToDevice, stream));
    const int B = 256, G = (E + B - 1) / B;
    k_et_leaves<<<G, B, 0, stream>>>(sc->dev.shapes, sc->dev.edges, E, co[0], co[1], co[2], leaves, is_cs, g);
    k_et_mad<<<G, B, 0, stream>>>(sc->dev.shapes, sc->dev.edges, E, leaves, is_cs, g);
    k_et_codes<<<G, B, 0, stream>>>(E, leaves, is_cs, g, keys, vals);
    RB_CUDA_OK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_sorted, vals, ids, E, 0, 64, stream));
    ETGlobals gh;
    RB_CUDA_OK(cudaMemcpyAsync(&gh, g, sizeof(gh), cudaMemcpyDeviceToHost, stream));
    RB_CUDA_OK(cudaStreamSynchronize(stream)); // the tree sizes decide the launches below
    double mad[3];
    for (int k = 0; k < 3; k++) mad[k] = gh.mad[k] / E;
    sc->dev.edge_bounds_expand = (float)(0.01 * std::sqrt(mad[0] * mad[0] + mad[1] * mad[1] + mad[2] * mad[2]));
    ETTree trees[2];
    trees[0] = ETTree{0, gh.n_cs, 0, 0};
    trees[1] = ETTree{gh.n_cs, E - gh.n_cs, gh.n_cs > 0 ? std::max(gh.n_cs - 1, 1) + gh.n_cs : 0, 1};
    const size_t NN = 2 * (size_t)E + 2;
    int *arrived = counters, *inner_count = counters + NN, *rank = counters + 2 * NN;
    int roots[2] = {RB_EDGE_EMPTY, RB_EDGE_EMPTY};
    int rank_base = 0;
    std::vector<int> first_ids(2, -1);
    for (int t = 0; t < 2; t++) {
        const ETTree& tr = trees[t];
        if (tr.L == 0) continue;
        const int GL = (tr.L + B - 1) / B;
        k_et_init<<<GL, B, 0, stream>>>(tr, leaves, ids, nodes);
        if (tr.L >= 2) {
            k_et_karras<<<GL, B, 0, stream>>>(tr, keys_sorted, ids, nodes);
            for (int pass = 0; pass < 3; pass++) {
                RB_CUDA_OK(cudaMemsetAsync(arrived, 0, sizeof(int) * NN, stream));
                if (pass == 0) k_et_climb<0><<<GL, B, 0, stream>>>(tr, nodes, arrived, inner_count);
                else if (pass == 1) k_et_optimize<<<(tr.L + RB_ET_WARPS - 1) / RB_ET_WARPS, 32 * RB_ET_WARPS, 0, stream>>>(tr, nodes, arrived);
                else k_et_climb<2><<<GL, B, 0, stream>>>(tr, nodes, arrived, inner_count);
            }
            k_et_rank<<<GL, B, 0, stream>>>(tr, nodes, inner_count, rank_base, rank);
            k_et_emit<<<GL, B, 0, stream>>>(tr, nodes, rank, out);
            roots[t] = rank_base; // the root is visited first
            rank_base += tr.L - 1;
        } else {
            RB_CUDA_OK(cudaMemcpyAsync(&first_ids[t], ids + tr.first, sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
    }
    RB_CUDA_OK(cudaGetLastError());
    RB_CUDA_OK(cudaStreamSynchronize(stream));
    for (int t = 0; t < 2; t++)
        if (trees[t].L == 1) roots[t] = ~first_ids[t];
    sc->dev.edge_nodes = out;
    sc->dev.edge_root_cs = roots[0];
    sc->dev.edge_root_ncs = roots[1];
    sc->num_edge_nodes = rank_base;
    release();
    return 0;
}

// ------------------------------------------------------------------------------------------------ primary-edge distribution
// Screen-space length of every camera-silhouette edge -> PMF / CDF of the primary-edge sampler (src/edge.cpp:186-214, :298-331),
// on the device: the other camera-dependent table of a scene (rb_scene_set_camera rebuilds it together with the trees).
__global__ void k_prim_weights(DevCamera cam, const rb_shape* shapes, const Edge* edges, int E, double* w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    double iw = 1.0 / cam.c2w[15];
    V3 org = mk3((Real)(cam.c2w[3] * iw), (Real)(cam.c2w[7] * iw), (Real)(cam.c2w[11] * iw));
    const Edge e = edges[i];
    V3 v0 = edge_v0(shapes, e), v1 = edge_v1(shapes, e);
    V2 p0, p1, c0, c1;
    double x = 0;
    if (cam_project(cam, v0, v1, p0, p1) && clip_line_unit(p0, p1, c0, c1) && edge_is_silhouette(shapes, org, e)) x = length(c1 - c0);
    w[i] = x;
}
__global__ void k_prim_normalize(int E, const double* total, double* pmf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    double t = *total;
    pmf[i] = t > 0 ? pmf[i] / t : 0.0;
}
int rb_build_primary_edge_cdf_gpu(rb_scene* sc, cudaStream_t stream) {
    const int E = sc->dev.num_edges;
    if (E == 0 || !sc->dev.use_primary_edge) return 0;
    double *pmf = const_cast<double*>(sc->dev.prim_edge_pmf), *cdf = const_cast<double*>(sc->dev.prim_edge_cdf);
    if (pmf == nullptr || cdf == nullptr) {
        if (cudaMallocAsync((void**)&pmf, sizeof(double) * (size_t)E, stream) != cudaSuccess || cudaMallocAsync((void**)&cdf, sizeof(double) * (size_t)E, stream) != cudaSuccess) {
            rb_set_error("rb_scene_create: out of device memory for the primary-edge distribution");
            return 1;
        }
        sc->allocs.push_back(pmf);
        sc->allocs.push_back(cdf);
        sc->dev.prim_edge_pmf = pmf;
        sc->dev.prim_edge_cdf = cdf;
    }
    size_t b1 = 0, b2 = 0;
    cub::DeviceReduce::Sum(nullptr, b1, pmf, (double*)nullptr, E, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, pmf, cdf, E, stream);
    void* tmp = nullptr;
    if (cudaMallocAsync(&tmp, std::max(b1, b2) + 256, stream) != cudaSuccess) {
        rb_set_error("rb_scene_create: out of device memory for the primary-edge distribution");
        return 1;
    }
    double* total = (double*)tmp;
    char* work = (char*)tmp + 256;
    const int B = 256, G = (E + B - 1) / B;
    k_prim_weights<<<G, B, 0, stream>>>(sc->dev.cam, sc->dev.shapes, sc->dev.edges, E, pmf);
    RB_CUDA_OK(cub::DeviceReduce::Sum(work, b1, pmf, total, E, stream));
    k_prim_normalize<<<G, B, 0, stream>>>(E, total, pmf);
    RB_CUDA_OK(cub::DeviceScan::ExclusiveSum(work, b2, pmf, cdf, E, stream));
    RB_CUDA_OK(cudaGetLastError());
    RB_CUDA_OK(cudaFreeAsync(tmp, stream));
    return 0;
}
