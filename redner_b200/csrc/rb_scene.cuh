// Host-side scene object behind the C ABI (rb_scene).  Mirrors what Scene::Scene builds in the reference
// (src/scene.cpp:63-307) with B200-native replacements: GPU LBVH instead of Embree/OptiX Prime, stream-ordered
// pool allocations instead of cudaMallocManaged buffers.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "rb_types.cuh"

// Timing events owned by a scene and reused by every rb_render call on it.
struct EventPool {
    std::vector<cudaEvent_t> ev;
    bool ensure(size_t n) {
        while (ev.size() < n) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return false;
            ev.push_back(e);
        }
        return true;
    }
    void destroy() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
    }
};

struct rb_scene {
    int device = 0;
    cudaStream_t stream = 0; // stream of rb_scene_create_on_stream: builds, rb_scene_set_camera and the frees of rb_scene_destroy
    EventPool events;
    DevScene dev;   // passed by value to kernels
    rb_camera cam;  // host copy of the descriptor camera
    std::vector<void*> allocs; // device allocations owned by the scene
    std::vector<rb_shape> shapes;
    std::vector<rb_material> materials;
    std::vector<DevLight> lights;
    int max_generic_texture_dimension = 0;
    int has_envmap = 0;
    // partition (multi-GPU)
    int part = 0, num_parts = 1, rows_per_stripe = 16;
    // stats of the last render
    int last_launches = 0;
    float last_kernel_ms = 0.f;
    float last_stage_ms[4] = {0.f, 0.f, 0.f, 0.f}; // k_forward, backward bands, k_primary_edge, k_finish_camera
    float last_bwd_ms[3] = {0.f, 0.f, 0.f};        // inside the bands: k_bwd_trace (+ work lists), boundary stage (pick, sort by edge, shade), k_bwd_sweep
    double last_path_vertices = 0, last_primary_hits = 0;
    int num_edge_nodes = 0; // records of the secondary-edge trees (dev.edge_nodes)
    bool edge_list_on_device = false;   // this scene's edge list was built by rb_edge_list.cu (else on the host)
    EdgeNode* edge_nodes_buf = nullptr; // device buffer of the GPU tree builder (num_edges records), reused by rb_scene_set_camera
    // scene-build timings (ms, host clock) for reporting
    float build_ms_bvh = 0.f, build_ms_lights = 0.f, build_ms_edges = 0.f;
};

void rb_set_error(const std::string& msg);
#define RB_CUDA_OK(expr)                                                                                       \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            rb_set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" +      \
                         std::to_string(__LINE__) + " (" #expr ")");                                           \
            return 1;                                                                                          \
        }                                                                                                      \
    } while (0)

// tables embedded into the shared object (rb_tables.cpp)
extern "C" const unsigned char rb_sobol_table_begin[];
extern "C" const unsigned char rb_sobol_table_end[];
extern "C" const unsigned char rb_ltc_table_begin[];
extern "C" const unsigned char rb_ltc_table_end[];

int rb_build_bvh(rb_scene* sc, cudaStream_t stream);
int rb_build_lights(rb_scene* sc, cudaStream_t stream);
int rb_build_edges(rb_scene* sc, cudaStream_t stream);
int rb_build_edge_list_gpu(rb_scene* sc, cudaStream_t stream);  // rb_edge_list.cu
int rb_build_edge_trees_gpu(rb_scene* sc, cudaStream_t stream); // rb_edge_tree.cu
int rb_build_primary_edge_cdf_gpu(rb_scene* sc, cudaStream_t stream);
