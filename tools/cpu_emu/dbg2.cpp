#include "emu.cpp"
#include <cstdio>
int main() {
    float verts0[9] = {-2.0f, 1.5f, 0.3f, 0.9f, 1.2f, -0.3f, -0.4f, -1.4f, 0.2f}; int idx0[3] = {0,1,2};
    float verts1[12] = {-1,-1,-7, 1,-1,-7, -1,1,-7, 1,1,-7}; int idx1[6] = {0,1,2, 1,3,2};
    rb_shape sh[2]; memset(sh, 0, sizeof(sh));
    sh[0].vertices = verts0; sh[0].indices = idx0; sh[0].num_vertices = 3; sh[0].num_triangles = 1; sh[0].material_id = 0; sh[0].light_id = -1;
    sh[1].vertices = verts1; sh[1].indices = idx1; sh[1].num_vertices = 4; sh[1].num_triangles = 2; sh[1].material_id = 0; sh[1].light_id = 0;
    float kd[3] = {0.5f,0.5f,0.5f}, ks[3] = {0,0,0}, rg[1] = {1.f}, uvs[2] = {1,1};
    rb_material m; memset(&m, 0, sizeof(m));
    m.diffuse_reflectance.texels[0] = kd; m.diffuse_reflectance.num_levels = 1; m.diffuse_reflectance.channels = 3; m.diffuse_reflectance.uv_scale = uvs;
    m.specular_reflectance.texels[0] = ks; m.specular_reflectance.num_levels = 1; m.specular_reflectance.channels = 3; m.specular_reflectance.uv_scale = uvs;
    m.roughness.texels[0] = rg; m.roughness.num_levels = 1; m.roughness.channels = 1; m.roughness.uv_scale = uvs;
    rb_area_light al; al.shape_id = 1; al.intensity[0] = al.intensity[1] = al.intensity[2] = 20; al.two_sided = 0; al.directly_visible = 1;
    rb_scene_desc d; memset(&d, 0, sizeof(d));
    rb_camera& c = d.camera;
    c.width = 64; c.height = 64; c.use_look_at = 1;
    float pos[3] = {0,0,-5}, look[3] = {0,0,0}, up[3] = {0,1,0};
    memcpy(c.position, pos, 12); memcpy(c.look, look, 12); memcpy(c.up, up, 12);
    float f = 1.0f / tanf(0.5f * 45.f * 3.14159265f / 180.f);
    float K[9] = {f,0,0, 0,f,0, 0,0,1}; float Ki[9] = {1/f,0,0, 0,1/f,0, 0,0,1};
    memcpy(c.intrinsic_mat, K, 36); memcpy(c.intrinsic_mat_inv, Ki, 36);
    c.clip_near = 1e-2f; c.camera_type = 0; c.viewport_end[0] = 64; c.viewport_end[1] = 64;
    d.num_shapes = 2; d.shapes = sh; d.num_materials = 1; d.materials = &m; d.num_lights = 1; d.lights = &al; d.use_gpu = 1;
    rb_scene* sc; if (rb_scene_create(&d, &sc)) { printf("err %s\n", rb_last_error()); return 1; }
    printf("root %d tris %d nodes %zu\n", sc->dev.bvh_root, sc->dev.num_tris, sc->nodes.size());
    Ray r; RayDiff rd; cam_primary_ray(sc->dev.cam, 0.4, 0.5, r, rd);
    Isect is = no_isect(); bool h = closest_hit(sc->dev, r, is);
    printf("hit %d shape %d tri %d\n", h, is.shape_id, is.tri_id);
    for (auto& n : sc->nodes) printf("node l %d r %d  L[%g %g|%g %g|%g %g] R[%g %g|%g %g|%g %g]\n", n.left, n.right, n.lo_x_hi_x.x, n.lo_x_hi_x.y, n.lo_y_hi_y.x, n.lo_y_hi_y.y, n.lo_z_hi_z.x, n.lo_z_hi_z.y, n.lo_x_hi_x.z, n.lo_x_hi_x.w, n.lo_y_hi_y.z, n.lo_y_hi_y.w, n.lo_z_hi_z.z, n.lo_z_hi_z.w);
    RenderParams rp; memset(&rp, 0, sizeof(rp)); rp.seed = 1; rp.spp = 4; rp.max_bounces = 1; rp.sampler_type = 1; rp.nd = 3; rp.vp_w = 64; rp.vp_h = 64; rp.num_parts = 1; rp.rows_per_stripe=16;
    V3 a = forward_sample(sc->dev, rp, 32*64+25, 25, 32, 0);
    printf("sample %g %g %g\n", a.x, a.y, a.z);
    return 0;
}
