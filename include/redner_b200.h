/*
 * redner_b200 -- C ABI of the B200-native differentiable path tracer.
 *
 * This is the drop-in boundary for the one hot path of BachiLi/redner:
 *   pyredner.RenderFunction.forward/backward  ->  redner.render(...)
 * Every entry point below replaces one piece of the reference's pybind11 surface
 * (reference file:line given per item; paths relative to the reference checkout).
 * Signatures use plain pointers and sizes only (no torch / pybind types).
 *
 * Memory convention (same as the reference, pyredner/render_pytorch.py:314-617):
 *   - the CALLER owns every buffer; the library stores raw pointers only;
 *   - mesh, texture, image and gradient buffers are DEVICE pointers (cuda:<gpu_index>);
 *   - camera parameters, light intensities and other small "host-read" parameters are
 *     passed BY VALUE inside the descriptors (the reference copies them at construction
 *     time: src/camera.h:44-65, src/area_light.h:18-20);
 *   - image and gradient buffers are pre-zeroed by the caller and are ACCUMULATED into.
 *
 * Error convention: every call returns 0 on success, non-zero on failure;
 * rb_last_error() returns a thread-local message (the reference assert()/exit(1)s instead:
 * src/cuda_utils.h:9-13, src/redner.h:173).
 */
#ifndef REDNER_B200_H
#define REDNER_B200_H

#include <stdint.h>

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RB_MAX_MIP_LEVELS 8 /* src/texture.h:11 max_num_texels */

/* src/camera.h:12-17 */
enum rb_camera_type { RB_CAMERA_PERSPECTIVE = 0, RB_CAMERA_ORTHOGRAPHIC = 1, RB_CAMERA_FISHEYE = 2, RB_CAMERA_PANORAMA = 3 };
/* src/pathtracer.h:11-14 */
enum rb_sampler_type { RB_SAMPLER_INDEPENDENT = 0, RB_SAMPLER_SOBOL = 1 };
/* src/channels.h:6-23 */
enum rb_channel {
    RB_CH_RADIANCE = 0, RB_CH_ALPHA, RB_CH_DEPTH, RB_CH_POSITION, RB_CH_GEOMETRY_NORMAL, RB_CH_SHADING_NORMAL,
    RB_CH_UV, RB_CH_BARYCENTRIC, RB_CH_DIFFUSE_REFLECTANCE, RB_CH_SPECULAR_REFLECTANCE, RB_CH_ROUGHNESS,
    RB_CH_GENERIC_TEXTURE, RB_CH_VERTEX_COLOR, RB_CH_SHAPE_ID, RB_CH_TRIANGLE_ID, RB_CH_MATERIAL_ID, RB_CH_COUNT
};

/* Camera -- src/camera.h:19-83 (constructor semantics: cam_to_world given <=> use_look_at == 0). */
typedef struct rb_camera {
    int width, height;
    int use_look_at;
    float position[3], look[3], up[3]; /* valid if use_look_at */
    float cam_to_world[16];            /* row-major; valid if !use_look_at */
    float world_to_cam[16];            /* row-major; valid if !use_look_at */
    float intrinsic_mat_inv[9];        /* row-major 3x3 */
    float intrinsic_mat[9];
    int has_distortion;
    float distortion[8]; /* k1..k6, p1, p2 -- src/camera.h:56-62 */
    float clip_near;
    int camera_type; /* rb_camera_type */
    int viewport_beg[2], viewport_end[2]; /* (x, y) -- src/camera.h:82 */
} rb_camera;

/* Shape -- src/shape.h:9-63.  All pointers are device pointers; optional ones may be NULL. */
typedef struct rb_shape {
    const float* vertices;     /* [num_vertices, 3] */
    const int* indices;        /* [num_triangles, 3] */
    const float* uvs;          /* [num_uv_vertices, 2] or NULL */
    const float* normals;      /* [num_normal_vertices, 3] or NULL */
    const int* uv_indices;     /* [num_triangles, 3] or NULL */
    const int* normal_indices; /* [num_triangles, 3] or NULL */
    const float* colors;       /* [num_vertices, 3] or NULL */
    int num_vertices, num_uv_vertices, num_normal_vertices, num_triangles;
    int material_id, light_id;
} rb_shape;

/* Texture<N> -- src/texture.h:14-46.  Constant texture <=> width[0] == 0 && height[0] == 0 (src/texture.h:342). */
typedef struct rb_texture {
    float* texels[RB_MAX_MIP_LEVELS]; /* device pointers, [h, w, channels] per level, or [channels] if constant */
    int width[RB_MAX_MIP_LEVELS];
    int height[RB_MAX_MIP_LEVELS];
    int channels;
    int num_levels;  /* 0 == texture absent */
    float* uv_scale; /* device pointer to 2 floats (may be NULL when num_levels == 0) */
} rb_texture;

/* Material -- src/material.h:12-91; DMaterial (src/material.h:93-99) uses the same layout with gradient buffers. */
typedef struct rb_material {
    rb_texture diffuse_reflectance;  /* 3 channels */
    rb_texture specular_reflectance; /* 3 channels */
    rb_texture roughness;            /* 1 channel */
    rb_texture generic_texture;      /* N channels, optional */
    rb_texture normal_map;           /* 3 channels, optional */
    int compute_specular_lighting, two_sided, use_vertex_color;
} rb_material;

/* AreaLight -- src/area_light.h:8-36 */
typedef struct rb_area_light {
    int shape_id;
    float intensity[3];
    int two_sided, directly_visible;
} rb_area_light;

/* EnvironmentMap -- src/envmap.h:19-51, constructor src/redner.cpp:169-178.  `values` is the [h, w, 3] mip pyramid, the two
 * tables are the caller's importance-sampling CDFs (pyredner/envmap.py:36-61); all device memory, matrices row-major. */
typedef struct rb_envmap {
    rb_texture values;
    float env_to_world[16], world_to_env[16];
    const float* sample_cdf_ys;
    const float* sample_cdf_xs;
    float pdf_norm;
    int directly_visible;
} rb_envmap;

/* Scene -- constructor arguments of src/scene.cpp:63-75 / src/redner.cpp:62-73 */
typedef struct rb_scene_desc {
    rb_camera camera;
    int num_shapes;
    const rb_shape* shapes; /* host array */
    int num_materials;
    const rb_material* materials; /* host array */
    int num_lights;
    const rb_area_light* lights; /* host array */
    const rb_envmap* envmap;     /* host pointer or NULL */
    int use_gpu;                 /* must be 1: there is no CPU fallback */
    int gpu_index;               /* -1 == current device (src/pathtracer.cpp:186-191) */
    int use_primary_edge_sampling;
    int use_secondary_edge_sampling;
} rb_scene_desc;

/* RenderOptions -- src/pathtracer.h:16-23 */
typedef struct rb_options {
    uint64_t seed;
    int num_samples;
    int max_bounces;
    int num_channels;
    const int* channels; /* host array of rb_channel */
    int sampler_type;    /* rb_sampler_type */
    int sample_pixel_center;
} rb_options;

/* DShape -- src/shape.h:65-80 (device pointers, any may be NULL) */
typedef struct rb_dshape {
    float *vertices, *uvs, *normals, *colors;
} rb_dshape;

/* DCamera -- src/camera.h:85-112 (device pointers) */
typedef struct rb_dcamera {
    float *position, *look, *up;      /* 3 floats each, used when the camera uses look-at */
    float *cam_to_world, *world_to_cam; /* 16 floats each */
    float *intrinsic_mat_inv, *intrinsic_mat; /* 9 floats each */
    float* distortion;                /* 8 floats or NULL */
} rb_dcamera;

/* DEnvironmentMap -- src/envmap.h:53-61: gradient mip pyramid and the 16 floats of d(world_to_env) (device memory) */
typedef struct rb_denvmap {
    rb_texture values;
    float* world_to_env;
} rb_denvmap;

/* DScene -- src/scene.h DScene / src/redner.cpp:75-82 */
typedef struct rb_dscene_desc {
    rb_dcamera camera;
    int num_shapes;
    const rb_dshape* shapes; /* host array */
    int num_materials;
    const rb_material* materials; /* host array; texel pointers are gradient buffers */
    int num_lights;
    float* const* light_intensity; /* host array of device pointers (3 floats each) -- src/area_light.h:38-43 */
    const rb_denvmap* envmap;      /* host pointer or NULL */
} rb_dscene_desc;

typedef struct rb_scene rb_scene;

/* Scene::Scene (src/scene.cpp:63-307): flattens the scene, builds the triangle BVH (replaces Embree / OptiX Prime,
 * src/scene.cpp:78-155), the light PMF/CDF and per-light area CDFs (src/scene.cpp:197-253), the edge list and
 * primary-edge distribution (src/edge.cpp:233-383). */
int rb_scene_create(const rb_scene_desc* desc, rb_scene** out);
/* The same on a caller-chosen CUDA stream (a cudaStream_t; NULL == legacy default stream, which is what rb_scene_create uses): uploads,
 * mesh read-back and build kernels are ordered after the work already queued on that stream -- pass the stream the geometry tensors
 * were produced on.  rb_scene_set_camera and rb_scene_destroy keep using it. */
int rb_scene_create_on_stream(const rb_scene_desc* desc, rb_scene** out, void* stream);
void rb_scene_destroy(rb_scene* scene);
/* Scene::max_generic_texture_dimension (src/scene.cpp:293-300, bound at src/redner.cpp:72) */
int rb_scene_max_generic_texture_dimension(const rb_scene* scene);

/* compute_num_channels (src/channels.cpp:42-113, bound at src/redner.cpp:201) */
int rb_compute_num_channels(const int* channels, int num_channels, int max_generic_texture_dimension);

/* render (src/pathtracer.cpp:177-958, declared src/pathtracer.h:25-31, bound at src/redner.cpp:257).
 * Forward pass <=> rendered_image != NULL; backward pass <=> d_rendered_image != NULL (then d_scene is required).
 * screen_gradient_image may be NULL.  `stream` is a cudaStream_t (NULL == legacy default stream); the call
 * synchronises the stream before returning, like the reference (src/pathtracer.cpp:947-949). */
int rb_render(const rb_scene* scene, const rb_options* options, float* rendered_image, const float* d_rendered_image,
              const rb_dscene_desc* d_scene, float* screen_gradient_image, void* stream);

/* Re-target a scene at another camera.  Only what depends on the camera is rebuilt, on the device: the primary-edge distribution
 * (src/edge.cpp:298-331) and the two secondary-edge trees (src/edge_tree.cpp:724-882); geometry, BVH, light tables and the edge
 * list are kept.  (The reference rebuilds the whole Scene per view, pyredner/render_pytorch.py:608-617.) */
int rb_scene_set_camera(rb_scene* scene, const rb_camera* camera);

/* A batch of views of one scene -- the native form of the per-view Python loops of pyredner/render_utils.py:407-430 and of
 * BASELINE config 5: for k in [0, num_views): rb_scene_set_camera(cameras[k]) then rb_render(options[k], images[k], d_images[k],
 * d_scenes[k]).  images / d_images / d_scenes may be NULL (or hold NULL entries) like the arguments of rb_render; gradients
 * ACCUMULATE, so one descriptor passed for every view sums the batch's gradients into one set of buffers. */
int rb_render_batch(rb_scene* scene, int num_views, const rb_camera* cameras, const rb_options* options, float* const* images,
                    const float* const* d_images, const rb_dscene_desc* const* d_scenes, void* stream);

/* Multi-GPU tile sharding (no reference counterpart).  Restricts subsequent rb_render calls on
 * this scene to the rows r with (r / rows_per_stripe) % num_parts == part of the viewport, while samplers stay
 * seeded by the full-viewport pixel index, so the union over parts equals the single-GPU result.  Primary-edge
 * samples are sharded by sample index.  num_parts == 1 restores the full image. */
int rb_scene_set_partition(rb_scene* scene, int part, int num_parts, int rows_per_stripe);

/* Statistics of the last rb_render on this scene: number of kernels launched and device milliseconds (CUDA events on
 * the render stream) spent inside the traced kernels. */
int rb_scene_last_stats(const rb_scene* scene, int* num_kernel_launches, float* kernel_ms);

/* Per-kernel device times of the last rb_render (CUDA events on the render stream), in launch order
 * { k_forward, backward bands (trace + boundary terms + sweep), k_primary_edge, k_finish_camera } (0 for kernels that
 * did not run), the number of path
 * vertices at which the last backward pass formed a radiance estimate and its number of primary hits (mean executed
 * bounces per sample = path_vertices / (W*H*spp)). */
int rb_scene_last_stage_stats(const rb_scene* scene, float* stage_ms4, double* path_vertices, double* primary_hits);
/* Split of the backward bands of the last rb_render, summed over the bands:
 * { k_bwd_trace, scan + compaction + k_bwd_secondary, k_bwd_sweep } in milliseconds. */
int rb_scene_last_backward_stats(const rb_scene* scene, float* bwd_ms3);
/* rb_render keeps one grow-only scratch allocation per device for the backward pass (gradient descriptors, path
 * records, work lists; at most ~1 GiB + small).  This frees them all; the next backward pass allocates again. */
void rb_release_scratch(void);
/* Host wall-clock milliseconds rb_scene_create spent in { BVH build, light tables, edge list + edge tree }. */
int rb_scene_build_ms(const rb_scene* scene, float* bvh_lights_edges3);

/* Test hook: the secondary-edge trees as the kernels see them.  info3 = { number of 128-byte records, root reference of the
 * camera-silhouette tree, root reference of the other tree } (reference >= 0: record index, < 0: ~edge id, INT_MIN: empty tree);
 * *expand = billboard size (src/edge_tree.cpp:773); records_out (may be NULL) receives up to records_bytes of the records. */
int rb_scene_edge_trees(const rb_scene* scene, int* info3, float* expand, void* records_out, size_t records_bytes);
/* Test hook: the scene's edge list (what collect_edges builds, src/edge.cpp:233-296): *num_edges, and up to edges_bytes of
 * { shape, v0, v1, f0, f1 } int records into edges_out (may be NULL). */
int rb_scene_edge_list(const rb_scene* scene, int* num_edges, int* edges_out, size_t edges_bytes);

const char* rb_last_error(void);
const char* rb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* REDNER_B200_H */
