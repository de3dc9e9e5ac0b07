// CPU check of redner_b200/csrc/rb_edge_list.cuh (TEST infrastructure).  The element functions that rb_edge_list.cu runs in kernels
// between CUB sorts / scans are run here in serial loops between std::stable_sort calls (a stable radix sort of keys == stable_sort by
// key; cub::DeviceMergeSort::StableSortKeys with a strict comparator == stable_sort), and the edge list is compared field by field with
// host_build_edges (rb_scene_host.hpp), the step-by-step restatement of src/edge.cpp:233-296 that every golden gradient was checked with.
//
//   edge_list_check <file.bin>...      meshes exported by tests/test_edge_list_cpu.py: int S, then per shape int nv, nt, float[3 nv], int[3 nt]
//   edge_list_check --random N         N random scenes: grids with seams, duplicated positions (large tie groups), fans sharing one edge,
//                                      degenerate and coplanar triangles, empty shapes
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "../redner_b200/csrc/rb_render.cuh"
#include "../redner_b200/csrc/rb_scene_host.hpp"
#include "../redner_b200/csrc/rb_edge_list.cuh"

static std::vector<Edge> edge_list_by_steps(const std::vector<rb_shape>& hs) {
    const int S = (int)hs.size();
    std::vector<int> tri_off(S + 1, 0), vert_off(S + 1, 0);
    for (int s = 0; s < S; s++) {
        tri_off[s + 1] = tri_off[s] + hs[s].num_triangles;
        vert_off[s + 1] = vert_off[s] + hs[s].num_vertices;
    }
    ELScene L{hs.data(), tri_off.data(), vert_off.data(), S, el_bits_for(vert_off[S])};
    const int n = 3 * tri_off[S];
    // A + stable sort of (key, half-edge)
    std::vector<unsigned long long> keys(n);
    for (int h = 0; h < n; h++) keys[h] = el_half_edge_key(L, h);
    std::vector<int> half(n);
    std::iota(half.begin(), half.end(), 0);
    std::stable_sort(half.begin(), half.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    std::vector<unsigned long long> ks(n);
    for (int i = 0; i < n; i++) ks[i] = keys[half[i]];
    // B with an inclusive scan of the head flags
    std::vector<int> incl(n);
    int M = 0;
    for (int i = 0; i < n; i++) incl[i] = (M += el_is_run_head(ks.data(), i) ? 1 : 0);
    std::vector<Edge> merged(M);
    for (int i = 0; i < n; i++)
        if (el_is_run_head(ks.data(), i)) merged[incl[i] - 1] = el_merge_run(L, ks.data(), half.data(), n, i);
    // C: reversed, then stable with the strict order
    std::vector<int> order(M);
    for (int p = 0; p < M; p++) order[p] = M - 1 - p;
    std::stable_sort(order.begin(), order.end(), ELPositionLess{hs.data(), merged.data()});
    // D + E
    std::vector<Edge> out;
    for (int p = 0; p < M; p++) {
        Edge e = el_pair_seam(hs.data(), merged.data(), order.data(), M, p);
        if (!edge_is_flat(hs.data(), e)) out.push_back(e);
    }
    return out;
}

struct Meshes {
    std::vector<HostMesh> meshes;
    std::vector<rb_shape> shapes;
    void finish() {
        shapes.assign(meshes.size(), rb_shape());
        for (size_t s = 0; s < meshes.size(); s++) {
            memset(&shapes[s], 0, sizeof(rb_shape));
            shapes[s].num_vertices = (int)meshes[s].vertices.size() / 3;
            shapes[s].num_triangles = (int)meshes[s].indices.size() / 3;
            shapes[s].vertices = meshes[s].vertices.data();
            shapes[s].indices = meshes[s].indices.data();
        }
    }
};

static int compare(const Meshes& m, const char* label, bool timing) {
    DevCamera cam;
    memset(&cam, 0, sizeof(cam));
    HostEdgeTables t;
    auto t0 = std::chrono::high_resolution_clock::now();
    host_build_edges(m.shapes, m.meshes, cam, false, t);
    auto t1 = std::chrono::high_resolution_clock::now();
    std::vector<Edge> e = edge_list_by_steps(m.shapes);
    auto t2 = std::chrono::high_resolution_clock::now();
    int bad = e.size() != t.edges.size();
    for (size_t i = 0; !bad && i < e.size(); i++) {
        const Edge &a = e[i], &b = t.edges[i];
        if (a.shape_id != b.shape_id || a.v0 != b.v0 || a.v1 != b.v1 || a.f0 != b.f0 || a.f1 != b.f1) {
            printf("MISMATCH %s edge %zu: steps (%d %d %d %d %d) host (%d %d %d %d %d)\n", label, i, a.shape_id, a.v0, a.v1, a.f0, a.f1, b.shape_id, b.v0, b.v1, b.f0, b.f1);
            bad = 1;
        }
    }
    if (e.size() != t.edges.size()) printf("MISMATCH %s: %zu edges by steps, %zu on the host\n", label, e.size(), t.edges.size());
    if (timing)
        printf("ok %s shapes %zu edges %zu host_ms %.3f steps_ms %.3f\n", label, m.shapes.size(), e.size(), std::chrono::duration<double, std::milli>(t1 - t0).count(),
               std::chrono::duration<double, std::milli>(t2 - t1).count());
    return bad;
}

static bool load(const char* path, Meshes& m) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    int S = 0;
    if (fread(&S, 4, 1, f) != 1) return false;
    m.meshes.assign(S, HostMesh());
    for (int s = 0; s < S; s++) {
        int nv[2];
        if (fread(nv, 4, 2, f) != 2) return false;
        m.meshes[s].vertices.resize(3 * (size_t)nv[0]);
        m.meshes[s].indices.resize(3 * (size_t)nv[1]);
        if (fread(m.meshes[s].vertices.data(), 4, 3 * (size_t)nv[0], f) != 3 * (size_t)nv[0]) return false;
        if (fread(m.meshes[s].indices.data(), 4, 3 * (size_t)nv[1], f) != 3 * (size_t)nv[1]) return false;
    }
    fclose(f);
    m.finish();
    return true;
}

// Random scene generator: every shape is one of a few families chosen to hit the corner cases of the list.
static void random_scene(std::mt19937& rng, Meshes& m) {
    auto U = [&](int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); };
    auto F = [&]() { return std::uniform_real_distribution<float>(-1.f, 1.f)(rng); };
    const int S = U(1, 5);
    m.meshes.assign(S, HostMesh());
    for (int s = 0; s < S; s++) {
        HostMesh& h = m.meshes[s];
        auto add_v = [&](float x, float y, float z) {
            h.vertices.push_back(x);
            h.vertices.push_back(y);
            h.vertices.push_back(z);
            return (int)h.vertices.size() / 3 - 1;
        };
        auto add_t = [&](int a, int b, int c) {
            h.indices.push_back(a);
            h.indices.push_back(b);
            h.indices.push_back(c);
        };
        switch (U(0, 5)) {
        case 0: { // grid with unshared vertices along some rows (seams), optionally flat (coplanar pairs are dropped)
            const int nx = U(2, 14), ny = U(2, 14);
            const bool flat = U(0, 2) == 0, seams = U(0, 1) == 1;
            std::vector<float> height((size_t)(nx + 1) * (ny + 1));
            for (float& z : height) z = flat ? 0.f : 0.3f * F();
            std::vector<int> id((size_t)(nx + 1) * (ny + 1));
            for (int y = 0; y <= ny; y++)
                for (int x = 0; x <= nx; x++) id[(size_t)y * (nx + 1) + x] = add_v((float)x, (float)y, height[(size_t)y * (nx + 1) + x]);
            for (int y = 0; y < ny; y++) {
                std::vector<int> top(nx + 1);
                for (int x = 0; x <= nx; x++) top[x] = id[(size_t)(y + 1) * (nx + 1) + x];
                if (seams && (y % 3) == 1) // this row of quads uses its own copies of the upper vertices
                    for (int x = 0; x <= nx; x++) top[x] = add_v((float)x, (float)(y + 1), height[(size_t)(y + 1) * (nx + 1) + x]);
                for (int x = 0; x < nx; x++) {
                    int a = id[(size_t)y * (nx + 1) + x], b = id[(size_t)y * (nx + 1) + x + 1];
                    add_t(a, b, top[x + 1]);
                    add_t(a, top[x + 1], top[x]);
                }
            }
            break;
        }
        case 1: { // triangle soup on a tiny lattice of positions: every vertex is its own copy -> large groups of equal segments
            const int T = U(1, 120), lattice = U(2, 3);
            for (int t = 0; t < T; t++) {
                int v[3];
                for (int k = 0; k < 3; k++) v[k] = add_v((float)U(0, lattice - 1), (float)U(0, lattice - 1), (float)U(0, 1));
                add_t(v[0], v[1], v[2]);
            }
            break;
        }
        case 2: { // fan: many faces share ONE index pair (runs of more than two half-edges), plus a degenerate triangle
            const int nf = U(1, 9);
            int a = add_v(0, 0, 0), b = add_v(0, 0, 1);
            for (int k = 0; k < nf; k++) add_t(a, b, add_v(std::cos(0.7f * k), std::sin(0.7f * k), 0.5f));
            add_t(a, a, b);
            break;
        }
        case 3: { // closed box with shared vertices (12 triangles, coplanar pairs on every face)
            int v[8];
            for (int k = 0; k < 8; k++) v[k] = add_v((float)(k & 1), (float)((k >> 1) & 1), (float)(k >> 2));
            const int q[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 5, 7, 3}};
            for (auto& f : q) {
                add_t(v[f[0]], v[f[1]], v[f[2]]);
                add_t(v[f[0]], v[f[2]], v[f[3]]);
            }
            break;
        }
        case 4: { // random indexed soup
            const int nv = U(3, 40), T = U(1, 80);
            for (int k = 0; k < nv; k++) add_v(F(), F(), F());
            for (int t = 0; t < T; t++) add_t(U(0, nv - 1), U(0, nv - 1), U(0, nv - 1));
            break;
        }
        default: // an empty shape (no triangles); sometimes with vertices
            if (U(0, 1)) add_v(F(), F(), F());
            break;
        }
    }
    m.finish();
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc >= 3 && std::string(argv[1]) == "--random") {
        const int N = atoi(argv[2]);
        std::mt19937 rng(20260923u);
        size_t edges = 0;
        for (int i = 0; i < N; i++) {
            Meshes m;
            random_scene(rng, m);
            bad += compare(m, ("random" + std::to_string(i)).c_str(), false);
            edges += 1;
        }
        printf("random scenes %d mismatching %d\n", N, bad);
        return bad != 0;
    }
    for (int a = 1; a < argc; a++) {
        Meshes m;
        if (!load(argv[a], m)) {
            printf("cannot read %s\n", argv[a]);
            return 2;
        }
        bad += compare(m, argv[a], true);
    }
    return bad != 0;
}
