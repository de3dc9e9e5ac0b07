// TEST INFRASTRUCTURE.  The reference's own finite-difference checks (north_star: "gradients pass the repo's own finite-difference
// checks"), restated for OUR adjoint functions: the rb_*.cuh device headers are compiled for the host with Real = double
// (g++ -DRB_REAL_DOUBLE -include tools/cpu_emu/emu_shim.h) and every hand-derived adjoint is compared with central differences of
// its primal, with the reference's inputs, step sizes and tolerance (equal_or_error, src/test_utils.h:15-23: |fd - analytic| < 1e-3):
//   test_d_intersect        src/shape.cpp:5-270     -> make_surface_point / d_make_surface_point
//   test_d_sample_shape     src/shape.cpp:272-331   -> sample_light_triangle / d_sample_light_triangle
//   test_d_bsdf             src/material.cpp:6-149  -> bsdf_eval / d_bsdf_eval
//   test_d_sample_primary_rays src/camera.cpp:98-276 -> cam_sample_primary / d_cam_sample_primary + finish_camera (pose adjoints)
//   test_d_camera_to_screen src/camera.cpp:278-439  -> cam_project / d_cam_project (point adjoints)
// Prints one line per check and exits non-zero on the first failure.  Built and run by tests/test_fd_functions_cpu.py.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../redner_b200/csrc/rb_render.cuh"
#include "../redner_b200/csrc/rb_scene_host.hpp"

static int g_checks = 0;
static void check(const char* what, int i, double fd, double analytic, double tol = 1e-3) {
    g_checks++;
    if (!(fabs(fd - analytic) <= tol)) {
        fprintf(stderr, "FD check failed: %s[%d]: finite difference %.9g, adjoint %.9g\n", what, i, fd, analytic);
        exit(1);
    }
}
static Real sum2(V2 v) { return v.x + v.y; }
static Real point_sum(const SurfacePoint& p, const RayDiff& rd) { // the reference sums these outputs (d_point = all ones)
    return sum(p.position) + sum(p.geom_normal) + sum(p.shading_frame.x) + sum(p.shading_frame.y) + sum(p.shading_frame.n) + sum2(p.uv) + sum2(p.du_dxy) +
           sum2(p.dv_dxy) + sum(p.dn_dx) + sum(p.dn_dy) + sum(rd.org_dx) + sum(rd.org_dy) + sum(rd.dir_dx) + sum(rd.dir_dy);
}
static SurfacePoint ones_point() {
    SurfacePoint d = zero_point();
    V3 one = mk3(1, 1, 1);
    d.position = d.geom_normal = d.dn_dx = d.dn_dy = d.color = one;
    d.shading_frame.x = d.shading_frame.y = d.shading_frame.n = one;
    d.uv = d.du_dxy = d.dv_dxy = mk2(1, 1);
    return d;
}

static void test_d_make_surface_point() {
    float vertices[9] = {-1.f, 0.f, 1.f, 1.f, 0.f, 1.f, 0.f, 1.f, 1.f};
    int indices[3] = {0, 1, 2};
    rb_shape shape;
    memset(&shape, 0, sizeof(shape));
    shape.vertices = vertices;
    shape.indices = indices;
    shape.num_vertices = 3;
    shape.num_triangles = 1;
    shape.material_id = 0;
    shape.light_id = -1;
    Ray ray;
    ray.org = mk3(0, 0, 0);
    ray.dir = mk3(0, 0, 1);
    ray.tmin = Real(1e-3);
    ray.tmax = INFINITY;
    RayDiff rd;
    rd.org_dx = rd.org_dy = rd.dir_dx = rd.dir_dy = mk3(1, 1, 1);
    SurfacePoint d_point = ones_point();
    RayDiff d_rd_out;
    d_rd_out.org_dx = d_rd_out.org_dy = d_rd_out.dir_dx = d_rd_out.dir_dy = mk3(1, 1, 1);
    DRay d_ray;
    d_ray.org = d_ray.dir = zero3();
    RayDiff d_rd = zero_raydiff();
    V3 d_vp[3] = {zero3(), zero3(), zero3()}, d_vn[3] = {zero3(), zero3(), zero3()}, d_vc[3] = {zero3(), zero3(), zero3()};
    V2 d_vuv[3] = {zero2(), zero2(), zero2()};
    d_make_surface_point(shape, 0, ray, rd, d_point, d_rd_out, d_ray, d_rd, d_vp, d_vn, d_vuv, d_vc);
    const Real h = Real(1e-4);
    auto eval = [&](const Ray& r, const RayDiff& q) {
        RayDiff out;
        SurfacePoint p = make_surface_point(shape, 0, r, q, out);
        return point_sum(p, out);
    };
    for (int i = 0; i < 3; i++) { // ray origin and direction
        Ray a = ray, b = ray;
        a.org[i] += h;
        b.org[i] -= h;
        check("d_ray.org", i, (eval(a, rd) - eval(b, rd)) / (2 * h), d_ray.org[i]);
        a = ray;
        b = ray;
        a.dir[i] += h;
        b.dir[i] -= h;
        check("d_ray.dir", i, (eval(a, rd) - eval(b, rd)) / (2 * h), d_ray.dir[i]);
    }
    V3* rdv[4] = {&rd.org_dx, &rd.org_dy, &rd.dir_dx, &rd.dir_dy};
    V3* d_rdv[4] = {&d_rd.org_dx, &d_rd.org_dy, &d_rd.dir_dx, &d_rd.dir_dy};
    const char* rdn[4] = {"d_ray_diff.org_dx", "d_ray_diff.org_dy", "d_ray_diff.dir_dx", "d_ray_diff.dir_dy"};
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 3; i++) {
            Real keep = (*rdv[k])[i];
            (*rdv[k])[i] = keep + h;
            Real fp = eval(ray, rd);
            (*rdv[k])[i] = keep - h;
            Real fn = eval(ray, rd);
            (*rdv[k])[i] = keep;
            check(rdn[k], i, (fp - fn) / (2 * h), (*d_rdv[k])[i]);
        }
    for (int v = 0; v < 3; v++) // vertex positions (float buffers: a larger step, like src/shape.cpp:196-233)
        for (int i = 0; i < 3; i++) {
            const float hv = 1e-2f;
            float keep = vertices[3 * v + i];
            vertices[3 * v + i] = keep + hv;
            Real fp = eval(ray, rd);
            vertices[3 * v + i] = keep - hv;
            Real fn = eval(ray, rd);
            vertices[3 * v + i] = keep;
            check("d_vertex", 3 * v + i, (fp - fn) / (2 * (Real)hv), d_vp[v][i], 5e-3);
        }
}

static void test_d_sample_light_triangle() {
    float vertices[9] = {-1.f, 0.f, 1.f, 1.f, 0.f, 1.f, 0.f, 1.f, 1.f};
    int indices[3] = {0, 1, 2};
    rb_shape shape;
    memset(&shape, 0, sizeof(shape));
    shape.vertices = vertices;
    shape.indices = indices;
    shape.num_vertices = 3;
    shape.num_triangles = 1;
    V2 smp = mk2(Real(0.5), Real(0.5));
    SurfacePoint d_p = zero_point();
    d_p.position = d_p.geom_normal = mk3(1, 1, 1);
    d_p.shading_frame.x = d_p.shading_frame.y = d_p.shading_frame.n = mk3(1, 1, 1);
    V3 d_v[3] = {zero3(), zero3(), zero3()};
    d_sample_light_triangle(shape, 0, smp, d_p, d_v);
    auto eval = [&]() {
        SurfacePoint p = sample_light_triangle(shape, 0, smp);
        return sum(p.position) + sum(p.geom_normal) + sum(p.shading_frame.x) + sum(p.shading_frame.y) + sum(p.shading_frame.n);
    };
    for (int v = 0; v < 3; v++)
        for (int i = 0; i < 3; i++) {
            const float hv = 1e-2f;
            float keep = vertices[3 * v + i];
            vertices[3 * v + i] = keep + hv;
            Real fp = eval();
            vertices[3 * v + i] = keep - hv;
            Real fn = eval();
            vertices[3 * v + i] = keep;
            check("d_light_vertex", 3 * v + i, (fp - fn) / (2 * (Real)hv), d_v[v][i], 5e-3);
        }
}

static rb_texture const_tex(float* data, int channels, float* uv_scale) {
    rb_texture t;
    memset(&t, 0, sizeof(t));
    t.texels[0] = data;
    t.width[0] = t.height[0] = 0;
    t.channels = channels;
    t.num_levels = 1;
    t.uv_scale = uv_scale;
    return t;
}
static void test_d_bsdf_eval() {
    float kd[3] = {0.5f, 0.4f, 0.3f}, ks[3] = {0.2f, 0.3f, 0.4f}, ro[1] = {0.5f}, uvs[2] = {1.f, 1.f};
    float d_kd[3] = {0, 0, 0}, d_ks[3] = {0, 0, 0}, d_ro[1] = {0}, d_uvs[2] = {0, 0};
    rb_material m, d_m;
    memset(&m, 0, sizeof(m));
    memset(&d_m, 0, sizeof(d_m));
    m.diffuse_reflectance = const_tex(kd, 3, uvs);
    m.specular_reflectance = const_tex(ks, 3, uvs);
    m.roughness = const_tex(ro, 1, uvs);
    m.compute_specular_lighting = 1;
    d_m.diffuse_reflectance = const_tex(d_kd, 3, d_uvs);
    d_m.specular_reflectance = const_tex(d_ks, 3, d_uvs);
    d_m.roughness = const_tex(d_ro, 1, d_uvs);
    d_m.compute_specular_lighting = 1;
    SurfacePoint p = zero_point();
    p.geom_normal = mk3(0, 1, 0);
    p.shading_frame = frame_from_normal(mk3(0, 1, 0));
    p.dpdu = mk3(1, 0, 0);
    p.uv = mk2(Real(0.5), Real(0.5));
    V3 wi = normalize(mk3(Real(0.5), 1, Real(0.5))), wo = normalize(mk3(Real(-0.5), 1, Real(-0.5)));
    SurfacePoint d_p = zero_point();
    V3 d_wi = zero3(), d_wo = zero3();
    d_bsdf_eval(m, d_m, p, wi, wo, Real(0), mk3(1, 1, 1), d_p, d_wi, d_wo);
    auto eval = [&](const SurfacePoint& q, V3 a, V3 b) { return sum(bsdf_eval(m, q, a, b, Real(0))); };
    const float hf = 1e-3f; // (float texels: the reference perturbs them by 1e-6 in double arithmetic; fp32 storage needs a larger step)
    for (int i = 0; i < 3; i++) {
        float keep = kd[i];
        kd[i] = keep + hf;
        Real fp = eval(p, wi, wo);
        kd[i] = keep - hf;
        Real fn = eval(p, wi, wo);
        kd[i] = keep;
        check("d_diffuse", i, (fp - fn) / (2 * (Real)hf), d_kd[i]);
        keep = ks[i];
        ks[i] = keep + hf;
        fp = eval(p, wi, wo);
        ks[i] = keep - hf;
        fn = eval(p, wi, wo);
        ks[i] = keep;
        check("d_specular", i, (fp - fn) / (2 * (Real)hf), d_ks[i]);
    }
    {
        float keep = ro[0];
        ro[0] = keep + hf;
        Real fp = eval(p, wi, wo);
        ro[0] = keep - hf;
        Real fn = eval(p, wi, wo);
        ro[0] = keep;
        check("d_roughness", 0, (fp - fn) / (2 * (Real)hf), d_ro[0], 2e-2); // (d_smithG1's 2.557 vs 2.577 slip of the reference is reproduced, src/material.h:581)
    }
    const Real h = Real(1e-6);
    V3* fr[3] = {&p.shading_frame.x, &p.shading_frame.y, &p.shading_frame.n};
    V3* d_fr[3] = {&d_p.shading_frame.x, &d_p.shading_frame.y, &d_p.shading_frame.n};
    const char* frn[3] = {"d_frame.x", "d_frame.y", "d_frame.n"};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) {
            Real keep = (*fr[k])[i];
            (*fr[k])[i] = keep + h;
            Real fp = eval(p, wi, wo);
            (*fr[k])[i] = keep - h;
            Real fn = eval(p, wi, wo);
            (*fr[k])[i] = keep;
            check(frn[k], i, (fp - fn) / (2 * h), (*d_fr[k])[i]);
        }
    for (int i = 0; i < 3; i++) {
        V3 a = wi, b = wi;
        a[i] += h;
        b[i] -= h;
        check("d_wi", i, (eval(p, a, wo) - eval(p, b, wo)) / (2 * h), d_wi[i]);
        a = wo;
        b = wo;
        a[i] += h;
        b[i] -= h;
        check("d_wo", i, (eval(p, wi, a) - eval(p, wi, b)) / (2 * h), d_wo[i]);
    }
    check("d_position", 0, 0, sum(d_p.position));
    check("d_geom_normal", 0, 0, sum(d_p.geom_normal));
}

static rb_camera test_camera(const float* pos, const float* look, const float* up) {
    rb_camera c;
    memset(&c, 0, sizeof(c));
    c.width = c.height = 64;
    c.use_look_at = 1;
    for (int i = 0; i < 3; i++) {
        c.position[i] = pos[i];
        c.look[i] = look[i];
        c.up[i] = up[i];
    }
    // perspective intrinsics of a 45 degree field of view (pyredner/camera.py:108-117)
    float f = 1.f / tanf(0.5f * 45.f * 3.14159265f / 180.f);
    float K[9] = {f, 0, 0, 0, f, 0, 0, 0, 1}, Ki[9] = {1 / f, 0, 0, 0, 1 / f, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++) {
        c.intrinsic_mat[i] = K[i];
        c.intrinsic_mat_inv[i] = Ki[i];
    }
    c.clip_near = 1e-2f;
    c.camera_type = RB_CAMERA_PERSPECTIVE;
    c.viewport_end[0] = c.viewport_end[1] = 64;
    return c;
}
// test_d_sample_primary_rays, src/camera.cpp:98-276: the primary ray w.r.t. the camera pose (position / look / up), through the
// matrix accumulators and finish_camera (d_look_at_matrix, src/transform.h:29-71)
static void test_d_cam_sample_primary() {
    float pos[3] = {0.3f, 0.4f, -5.f}, look[3] = {0.1f, -0.2f, 0.f}, up[3] = {0.f, 1.f, 0.f};
    const Real sx = Real(0.3), sy = Real(0.6);
    rb_camera c = test_camera(pos, look, up);
    DevCamera cam;
    host_setup_camera(c, cam);
    DRay d_ray;
    d_ray.org = d_ray.dir = mk3(1, 1, 1);
    float accf[RB_CAM_ACC];
    for (float& a : accf) a = 0.f;
    CamAcc acc;
    acc.base = accf;
    acc.stride = 1;
    d_cam_sample_primary(cam, sx, sy, d_ray, acc, nullptr);
    double accd[RB_CAM_ACC];
    for (int i = 0; i < RB_CAM_ACC; i++) accd[i] = accf[i];
    float d_pos[3] = {0, 0, 0}, d_look[3] = {0, 0, 0}, d_up[3] = {0, 0, 0};
    rb_dcamera out;
    memset(&out, 0, sizeof(out));
    out.position = d_pos;
    out.look = d_look;
    out.up = d_up;
    finish_camera(cam, accd, out);
    auto eval = [&](const float* p, const float* l, const float* u) {
        rb_camera cc = test_camera(p, l, u);
        DevCamera dc;
        host_setup_camera(cc, dc);
        D3 o, d;
        cam_sample_primary(dc, (double)sx, (double)sy, o, d);
        return o.x + o.y + o.z + d.x + d.y + d.z;
    };
    float* params[3] = {pos, look, up};
    float* grads[3] = {d_pos, d_look, d_up};
    const char* names[3] = {"d_camera.position", "d_camera.look", "d_camera.up"};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) {
            const float h = 1e-2f;
            float keep = params[k][i];
            params[k][i] = keep + h;
            double fp = eval(pos, look, up);
            params[k][i] = keep - h;
            double fn = eval(pos, look, up);
            params[k][i] = keep;
            check(names[k], i, (fp - fn) / (2.0 * h), grads[k][i], 2e-3);
        }
}
// test_d_camera_to_screen, src/camera.cpp:278-439: screen position of a segment's end points w.r.t. the points
static void test_d_cam_project() {
    float pos[3] = {0.3f, 0.4f, -5.f}, look[3] = {0.1f, -0.2f, 0.f}, up[3] = {0.f, 1.f, 0.f};
    rb_camera c = test_camera(pos, look, up);
    DevCamera cam;
    host_setup_camera(c, cam);
    V3 p0 = mk3(Real(-0.7), Real(0.4), Real(0.3)), p1 = mk3(Real(0.8), Real(-0.3), Real(1.1));
    float accf[RB_CAM_ACC];
    for (float& a : accf) a = 0.f;
    CamAcc acc;
    acc.base = accf;
    acc.stride = 1;
    V3 d_p0 = zero3(), d_p1 = zero3();
    d_cam_project(cam, p0, p1, 1, 1, 1, 1, acc, d_p0, d_p1);
    auto eval = [&](V3 a, V3 b) {
        V2 q0, q1;
        if (!cam_project(cam, a, b, q0, q1)) return Real(0);
        return q0.x + q0.y + q1.x + q1.y;
    };
    const Real h = Real(1e-5);
    for (int i = 0; i < 3; i++) {
        V3 a = p0, b = p0;
        a[i] += h;
        b[i] -= h;
        check("d_project.p0", i, (eval(a, p1) - eval(b, p1)) / (2 * h), d_p0[i]);
        a = p1;
        b = p1;
        a[i] += h;
        b[i] -= h;
        check("d_project.p1", i, (eval(p0, a) - eval(p0, b)) / (2 * h), d_p1[i]);
    }
}

int main() {
    test_d_make_surface_point();
    printf("ok make_surface_point / d_make_surface_point (src/shape.cpp:5-270)\n");
    test_d_sample_light_triangle();
    printf("ok sample_light_triangle / d_sample_light_triangle (src/shape.cpp:272-331)\n");
    test_d_bsdf_eval();
    printf("ok bsdf_eval / d_bsdf_eval (src/material.cpp:6-149)\n");
    test_d_cam_sample_primary();
    printf("ok cam_sample_primary / d_cam_sample_primary + finish_camera (src/camera.cpp:98-276)\n");
    test_d_cam_project();
    printf("ok cam_project / d_cam_project (src/camera.cpp:278-439)\n");
    printf("checks %d\n", g_checks);
    return 0;
}
