#include "emu_shim.h"
#include <cstdio>
#include "../../redner_b200/csrc/rb_render.cuh"
#include "../../redner_b200/csrc/rb_scene_host.hpp"
int main() {
    rb_camera c; memset(&c, 0, sizeof(c));
    c.width = 64; c.height = 64; c.use_look_at = 1;
    float pos[3] = {0,0,-5}, look[3] = {0,0,0}, up[3] = {0,1,0};
    memcpy(c.position, pos, 12); memcpy(c.look, look, 12); memcpy(c.up, up, 12);
    float f = 1.0f / tanf(0.5f * 45.f * 3.14159265f / 180.f);
    float K[9] = {f,0,0, 0,f,0, 0,0,1}; float Ki[9] = {1/f,0,0, 0,1/f,0, 0,0,1};
    memcpy(c.intrinsic_mat, K, 36); memcpy(c.intrinsic_mat_inv, Ki, 36);
    c.clip_near = 1e-2f; c.camera_type = 0; c.viewport_end[0] = 64; c.viewport_end[1] = 64;
    DevCamera dc; host_setup_camera(c, dc);
    for (int i = 0; i < 16; i++) printf("%g ", dc.c2w[i]); printf("\n");
    Ray r; RayDiff rd; cam_primary_ray(dc, 0.5, 0.5, r, rd);
    printf("org %g %g %g dir %g %g %g\n", r.org.x, r.org.y, r.org.z, r.dir.x, r.dir.y, r.dir.z);
    // triangle test
    BVHTri t; t.v0 = make_float4(-2,1.5,0.3,0); t.v1 = make_float4(0.9,1.2,-0.3,0); t.v2 = make_float4(-0.4,-1.4,0.2,0);
    float tt; bool h = tri_test(f3(0,0,-5), f3(0,0,1), 1e-3f, INFINITY, t, tt);
    printf("hit %d t %g\n", (int)h, tt);
    return 0;
}
