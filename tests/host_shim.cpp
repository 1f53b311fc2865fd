// TEST INFRASTRUCTURE ONLY.  Compiles fresnel_b200/csrc/frb_math.h (the arithmetic the CUDA
// kernels inline) with g++ so that the CPU-only test tier can check it against the oracle:
// bit-exact projection / rectangles / visibility and the hand-derived projection backward.
// Nothing in the product imports or links this file.
//   g++ -O2 -ffp-contract=off -shared -fPIC tests/host_shim.cpp -o tests/_host_shim.so
#include <cstring>

#include "../fresnel_b200/csrc/frb_math.h"
#include "../fresnel_b200/csrc/frb_head.h"

static FrbCamera cam_from(const float* c) {
    FrbCamera cam;
    std::memcpy(cam.V, c, 12 * sizeof(float));
    cam.fx = c[12]; cam.fy = c[13]; cam.cx = c[14]; cam.cy = c[15];
    cam.width = c[16]; cam.height = c[17]; cam.near_ = c[18]; cam.far_ = c[19];
    return cam;
}

extern "C" {

// out_f: [u, v, depth, a, b, c, d, radius, A, B, C] per Gaussian; out_i: [visible, x0, x1, y0, y1]
void shim_project_fwd(int n, const float* p, const float* s, const float* q, const float* cam20,
                      float max_radius, float* out_f, int* out_i) {
    FrbCamera cam = cam_from(cam20);
    for (int i = 0; i < n; ++i) {
        FrbProjTmp t;
        FrbProjected o;
        frb_project_core(p + 3 * i, s + 3 * i, q + 4 * i, cam, t, o);
        frb_project_finish(cam, max_radius, o);
        float* f = out_f + 11 * i;
        f[0] = o.u; f[1] = o.v; f[2] = o.depth; f[3] = o.a; f[4] = o.b; f[5] = o.c; f[6] = o.d;
        f[7] = o.radius; f[8] = o.A; f[9] = o.B; f[10] = o.C;
        int* k = out_i + 5 * i;
        k[0] = o.visible; k[1] = o.x0; k[2] = o.x1; k[3] = o.y0; k[4] = o.y1;
    }
}

// g2d: [g_u, g_v, g_A, g_B, g_C, g_depth] per Gaussian (A, B, C are the PRE-SCALED conic)
void shim_project_bwd(int n, const float* p, const float* s, const float* q, const float* cam20,
                      const float* g2d, float* gp, float* gs, float* gq) {
    FrbCamera cam = cam_from(cam20);
    for (int i = 0; i < n; ++i) {
        const float* g = g2d + 6 * i;
        frb_project_bwd_one(p + 3 * i, s + 3 * i, q + 4 * i, cam, g[0], g[1], g[2], g[3], g[4], g[5],
                            gp + 3 * i, gs + 3 * i, gq + 4 * i);
    }
}

// mode-aware variants (FRB_MODE_DENSE / FRB_MODE_FOURIER: the other renderers of the reference file)
void shim_project_fwd_mode(int n, const float* p, const float* s, const float* q, const float* cam20,
                           float max_radius, int mode, float* out_f, int* out_i) {
    FrbCamera cam = cam_from(cam20);
    for (int i = 0; i < n; ++i) {
        FrbProjTmp t;
        FrbProjected o;
        frb_project_core(p + 3 * i, s + 3 * i, q + 4 * i, cam, t, o);
        frb_project_finish(cam, max_radius, o, mode);
        float* f = out_f + 11 * i;
        f[0] = o.u; f[1] = o.v; f[2] = o.depth; f[3] = o.a; f[4] = o.b; f[5] = o.c; f[6] = o.d;
        f[7] = o.radius; f[8] = o.A; f[9] = o.B; f[10] = o.C;
        int* k = out_i + 5 * i;
        k[0] = o.visible; k[1] = o.x0; k[2] = o.x1; k[3] = o.y0; k[4] = o.y1;
    }
}

void shim_project_bwd_mode(int n, const float* p, const float* s, const float* q, const float* cam20,
                           const float* g2d, int mode, float* gp, float* gs, float* gq) {
    FrbCamera cam = cam_from(cam20);
    for (int i = 0; i < n; ++i) {
        const float* g = g2d + 6 * i;
        frb_project_bwd_one(p + 3 * i, s + 3 * i, q + 4 * i, cam, g[0], g[1], g[2], g[3], g[4], g[5],
                            gp + 3 * i, gs + 3 * i, gq + 4 * i, mode);
    }
}

// decoder output head (frb_head.h): out = [pos 3 | scl 3 | rot 4 | col 3 | opa 1] per Gaussian
void shim_head_fwd(int n, const float* raw, const float* base_xyz, const float* edge, float esf, float eob,
                   float* out) {
    for (int i = 0; i < n; ++i) {
        FrbHeadOut o;
        frb_head_fwd_one(raw + 16 * i, base_xyz[3 * i], base_xyz[3 * i + 1], base_xyz[3 * i + 2], edge[i], esf, eob, o);
        float* f = out + 14 * i;
        for (int k = 0; k < 3; ++k) { f[k] = o.pos[k]; f[3 + k] = o.scl[k]; f[10 + k] = o.col[k]; }
        for (int k = 0; k < 4; ++k) f[6 + k] = o.rot[k];
        f[13] = o.opa;
    }
}

void shim_head_bwd(int n, const float* raw, const float* edge, float esf, float eob, const float* g_out,
                   float* g_raw, float* g_z) {
    for (int i = 0; i < n; ++i) {
        FrbHeadOut g;
        const float* f = g_out + 14 * i;
        for (int k = 0; k < 3; ++k) { g.pos[k] = f[k]; g.scl[k] = f[3 + k]; g.col[k] = f[10 + k]; }
        for (int k = 0; k < 4; ++k) g.rot[k] = f[6 + k];
        g.opa = f[13];
        float g_edge;
        frb_head_bwd_one(raw + 16 * i, edge[i], esf, eob, g, g_raw + 16 * i, g_z[i], g_edge);
    }
}

// the full head as head.cu sequences it: zone snap of the depth, head, pose rotation; and its backward with the
// edge-strength gradient.  depth: per-Gaussian depth-grid value; trig: per-Gaussian (cos az, sin az, cos el, sin el).
void shim_head_full_fwd(int n, const float* raw, const float* base_xy, const float* depth, float depth_offset,
                        const float* edge, float esf, float eob, int num_zones, const float* boundaries,
                        const float* centers, const float* trig, float* out) {
    for (int i = 0; i < n; ++i) {
        float d = depth[i];
        if (num_zones > 0) d = frb_zone_center(d, boundaries, centers, num_zones);
        const float z = depth_offset + d * -2.0f;
        FrbHeadOut o;
        frb_head_fwd_one(raw + 16 * i, base_xy[2 * i], base_xy[2 * i + 1], z, edge[i], esf, eob, o);
        if (trig) {
            float r[3];
            frb_pose_rotate(o.pos, trig + 4 * i, r);
            o.pos[0] = r[0]; o.pos[1] = r[1]; o.pos[2] = r[2];
        }
        float* f = out + 14 * i;
        for (int k = 0; k < 3; ++k) { f[k] = o.pos[k]; f[3 + k] = o.scl[k]; f[10 + k] = o.col[k]; }
        for (int k = 0; k < 4; ++k) f[6 + k] = o.rot[k];
        f[13] = o.opa;
    }
}

void shim_head_full_bwd(int n, const float* raw, const float* edge, float esf, float eob, const float* trig,
                        const float* g_out, float* g_raw, float* g_z, float* g_edge) {
    for (int i = 0; i < n; ++i) {
        FrbHeadOut g;
        const float* f = g_out + 14 * i;
        for (int k = 0; k < 3; ++k) { g.pos[k] = f[k]; g.scl[k] = f[3 + k]; g.col[k] = f[10 + k]; }
        for (int k = 0; k < 4; ++k) g.rot[k] = f[6 + k];
        g.opa = f[13];
        if (trig) {
            float gp[3];
            frb_pose_rotate_bwd(g.pos, trig + 4 * i, gp);
            g.pos[0] = gp[0]; g.pos[1] = gp[1]; g.pos[2] = gp[2];
        }
        frb_head_bwd_one(raw + 16 * i, edge[i], esf, eob, g, g_raw + 16 * i, g_z[i], g_edge[i]);
    }
}
}
