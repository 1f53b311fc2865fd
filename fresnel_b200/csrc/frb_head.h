/*
 * Decoder output head: the 16 raw MLP outputs of one Gaussian -> renderer parameters, and the backward.
 * Written once as __host__ __device__ inline functions (as frb_math.h): the CUDA kernels in head.cu are
 * the only product callers; tests/host_shim.cpp compiles it with g++ to check the hand-derived backward
 * against autograd of the PyTorch restatement (fresnel_b200/training.py), without a GPU.
 *
 * Reference (GM = scripts/models/gaussian_decoder_models.py):
 *   DirectPatchDecoder.forward tail GM:807-948: grid + offset positions, Fresnel zone snap of the depth GM:833-838,
 *   pose rotation GM:51-104 / GM:860, softplus scales, 6D rotations, sigmoid colour / opacity and the edge-aware
 *   modulation GM:881-895
 *   rotation_6d_to_quaternion GM:186-276 (the +-1e-8 random sign jitter of GM:208 is a fixed +1e-8)
 * Raw layout (GM:795-800): [0:3] position offset (z unused: depth is locked, GM:844-850), [3:6] scale,
 * [6:12] 6D rotation, [12:15] colour, [15] opacity.
 */
#ifndef FRB_HEAD_H
#define FRB_HEAD_H

#include <math.h>

#if defined(__CUDACC__)
#define FRB_HEAD_HD __host__ __device__ __forceinline__
#else
#define FRB_HEAD_HD inline
#endif

#define FRB_HEAD_RAW 16

struct FrbHeadOut {
    float pos[3], scl[3], rot[4], col[3], opa;
};

/* Intermediates of the 6D -> quaternion map that the backward needs. */
struct FrbRot6dTmp {
    float b1[3], b2[3], b3[3];
    float n1, n2, n3raw, n3;      /* norms before clamping (n3raw: of b1 x b2, n3: of the selected vector) */
    float d;                      /* b1 . a2 */
    int degenerate;               /* b1 x b2 replaced by e_z (GM:213) */
    int kase;                     /* selected branch 0..3 (GM:262-271) */
    float s;                      /* s_k of the selected branch */
    float t;                      /* argument of the sqrt of the selected branch, before the 1e-10 clamp */
    float qs[4];                  /* selected quaternion before normalisation */
    float nq;                     /* its norm */
};

FRB_HEAD_HD float frb_sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

FRB_HEAD_HD float frb_norm3(const float v[3]) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* y = x / max(||x||, eps)  (F.normalize); returns ||x|| */
FRB_HEAD_HD float frb_normalize3(const float x[3], float eps, float y[3]) {
    float n = frb_norm3(x);
    float den = n > eps ? n : eps;
    y[0] = x[0] / den; y[1] = x[1] / den; y[2] = x[2] / den;
    return n;
}

/* backward of y = x / max(n, eps): clamp(min) passes its gradient where n >= eps */
FRB_HEAD_HD void frb_normalize3_bwd(const float y[3], float n, float eps, const float gy[3], float gx[3]) {
    if (n >= eps) {
        float dot = y[0] * gy[0] + y[1] * gy[1] + y[2] * gy[2];
        float inv = 1.0f / n;
        gx[0] = (gy[0] - y[0] * dot) * inv; gx[1] = (gy[1] - y[1] * dot) * inv; gx[2] = (gy[2] - y[2] * dot) * inv;
    } else {
        float inv = 1.0f / eps;
        gx[0] = gy[0] * inv; gx[1] = gy[1] * inv; gx[2] = gy[2] * inv;
    }
}

FRB_HEAD_HD void frb_rot6d_to_quat(const float r6[6], float q[4], FrbRot6dTmp& t) {
    const float* a1 = r6;
    const float* a2 = r6 + 3;
    t.n1 = frb_normalize3(a1, 1e-6f, t.b1);                                   /* GM:205 */
    t.d = t.b1[0] * a2[0] + t.b1[1] * a2[1] + t.b1[2] * a2[2];
    float v[3] = {a2[0] - t.d * t.b1[0] + 1e-8f, a2[1] - t.d * t.b1[1] + 1e-8f, a2[2] - t.d * t.b1[2] + 1e-8f};
    t.n2 = frb_normalize3(v, 1e-6f, t.b2);                                    /* GM:206-209 */
    float c[3] = {t.b1[1] * t.b2[2] - t.b1[2] * t.b2[1], t.b1[2] * t.b2[0] - t.b1[0] * t.b2[2],
                  t.b1[0] * t.b2[1] - t.b1[1] * t.b2[0]};                     /* GM:210 */
    t.n3raw = frb_norm3(c);
    t.degenerate = t.n3raw < 1e-6f;                                           /* GM:212-213 */
    if (t.degenerate) { c[0] = 0.0f; c[1] = 0.0f; c[2] = 1.0f; }
    t.n3 = frb_normalize3(c, 1e-6f, t.b3);                                    /* GM:214 */

    /* R = [b1 b2 b3] as columns (GM:217) */
    const float R00 = t.b1[0], R01 = t.b2[0], R02 = t.b3[0];
    const float R10 = t.b1[1], R11 = t.b2[1], R12 = t.b3[1];
    const float R20 = t.b1[2], R21 = t.b2[2], R22 = t.b3[2];
    const float trace = R00 + R11 + R22;
    if (trace > 0.0f) {                                                       /* GM:230-235 */
        t.kase = 0; t.t = trace + 1.0f;
        t.s = sqrtf(t.t > 1e-10f ? t.t : 1e-10f) * 2.0f;
        t.qs[0] = 0.25f * t.s; t.qs[1] = (R21 - R12) / t.s; t.qs[2] = (R02 - R20) / t.s; t.qs[3] = (R10 - R01) / t.s;
    } else if (R00 > R11 && R00 > R22) {                                      /* GM:237-242 */
        t.kase = 1; t.t = 1.0f + R00 - R11 - R22;
        t.s = sqrtf(t.t > 1e-10f ? t.t : 1e-10f) * 2.0f;
        t.qs[0] = (R21 - R12) / t.s; t.qs[1] = 0.25f * t.s; t.qs[2] = (R01 + R10) / t.s; t.qs[3] = (R02 + R20) / t.s;
    } else if (R11 > R22) {                                                   /* GM:244-249 */
        t.kase = 2; t.t = 1.0f + R11 - R00 - R22;
        t.s = sqrtf(t.t > 1e-10f ? t.t : 1e-10f) * 2.0f;
        t.qs[0] = (R02 - R20) / t.s; t.qs[1] = (R01 + R10) / t.s; t.qs[2] = 0.25f * t.s; t.qs[3] = (R12 + R21) / t.s;
    } else {                                                                  /* GM:251-256 */
        t.kase = 3; t.t = 1.0f + R22 - R00 - R11;
        t.s = sqrtf(t.t > 1e-10f ? t.t : 1e-10f) * 2.0f;
        t.qs[0] = (R10 - R01) / t.s; t.qs[1] = (R02 + R20) / t.s; t.qs[2] = (R12 + R21) / t.s; t.qs[3] = 0.25f * t.s;
    }
    t.nq = sqrtf(t.qs[0] * t.qs[0] + t.qs[1] * t.qs[1] + t.qs[2] * t.qs[2] + t.qs[3] * t.qs[3]);
    const float den = t.nq > 1e-6f ? t.nq : 1e-6f;                            /* GM:274 */
    q[0] = t.qs[0] / den; q[1] = t.qs[1] / den; q[2] = t.qs[2] / den; q[3] = t.qs[3] / den;
}

/* g_r6 = (d q / d r6)^T g_q */
FRB_HEAD_HD void frb_rot6d_to_quat_bwd(const float r6[6], const float gq[4], float g_r6[6]) {
    float q[4];
    FrbRot6dTmp t;
    frb_rot6d_to_quat(r6, q, t);
    const float* a2 = r6 + 3;

    /* quaternion normalisation */
    float gs[4];
    if (t.nq >= 1e-6f) {
        float dot = q[0] * gq[0] + q[1] * gq[1] + q[2] * gq[2] + q[3] * gq[3];
        float inv = 1.0f / t.nq;
        for (int k = 0; k < 4; ++k) gs[k] = (gq[k] - q[k] * dot) * inv;
    } else {
        for (int k = 0; k < 4; ++k) gs[k] = gq[k] * 1e6f;
    }

    /* selected branch: q_k = 0.25 s at the diagonal slot, (Ra +- Rb) / s elsewhere */
    float gR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    const float is = 1.0f / t.s;
    float g_s = 0.0f;
    /* off-diagonal combinations: numerator index pairs and signs per branch and slot */
    /* kase 0: w=.25s, x=(R21-R12), y=(R02-R20), z=(R10-R01)
       kase 1: w=(R21-R12), x=.25s, y=(R01+R10), z=(R02+R20)
       kase 2: w=(R02-R20), x=(R01+R10), y=.25s, z=(R12+R21)
       kase 3: w=(R10-R01), x=(R02+R20), y=(R12+R21), z=.25s */
    const int PA[4][4][2] = {{{-1, -1}, {2, 1}, {0, 2}, {1, 0}},
                             {{2, 1}, {-1, -1}, {0, 1}, {0, 2}},
                             {{0, 2}, {0, 1}, {-1, -1}, {1, 2}},
                             {{1, 0}, {0, 2}, {1, 2}, {-1, -1}}};
    const float SG[4][4] = {{0.f, -1.f, -1.f, -1.f}, {-1.f, 0.f, 1.f, 1.f}, {-1.f, 1.f, 0.f, 1.f}, {-1.f, 1.f, 1.f, 0.f}};
    for (int k = 0; k < 4; ++k) {
        const int i = PA[t.kase][k][0], j = PA[t.kase][k][1];
        if (i < 0) {
            g_s += 0.25f * gs[k];
        } else {
            /* q_k = (R[i][j] + sg * R[j][i]) / s */
            const float sg = SG[t.kase][k];
            gR[i][j] += gs[k] * is;
            gR[j][i] += sg * gs[k] * is;
            g_s -= gs[k] * t.qs[k] * is;
        }
    }
    /* s = 2 sqrt(max(t, 1e-10)) */
    const float g_t = (t.t >= 1e-10f) ? g_s * 2.0f / t.s : 0.0f;      /* d(2 sqrt t)/dt = 1/sqrt t = 2/s */
    {
        const float sgn[4][3] = {{1.f, 1.f, 1.f}, {1.f, -1.f, -1.f}, {-1.f, 1.f, -1.f}, {-1.f, -1.f, 1.f}};
        gR[0][0] += sgn[t.kase][0] * g_t;
        gR[1][1] += sgn[t.kase][1] * g_t;
        gR[2][2] += sgn[t.kase][2] * g_t;
    }
    /* columns of R */
    float g_b1[3] = {gR[0][0], gR[1][0], gR[2][0]};
    float g_b2[3] = {gR[0][1], gR[1][1], gR[2][1]};
    float g_b3[3] = {gR[0][2], gR[1][2], gR[2][2]};

    /* b3 = normalize(sel(b1 x b2)) */
    float g_c[3];
    frb_normalize3_bwd(t.b3, t.n3, 1e-6f, g_b3, g_c);
    if (!t.degenerate) {
        /* c = b1 x b2: g_b1 += b2 x g_c, g_b2 += g_c x b1 */
        g_b1[0] += t.b2[1] * g_c[2] - t.b2[2] * g_c[1];
        g_b1[1] += t.b2[2] * g_c[0] - t.b2[0] * g_c[2];
        g_b1[2] += t.b2[0] * g_c[1] - t.b2[1] * g_c[0];
        g_b2[0] += g_c[1] * t.b1[2] - g_c[2] * t.b1[1];
        g_b2[1] += g_c[2] * t.b1[0] - g_c[0] * t.b1[2];
        g_b2[2] += g_c[0] * t.b1[1] - g_c[1] * t.b1[0];
    }
    /* b2 = normalize(v), v = a2 - d b1 + 1e-8, d = b1 . a2 */
    float g_v[3];
    frb_normalize3_bwd(t.b2, t.n2, 1e-6f, g_b2, g_v);
    const float g_d = -(t.b1[0] * g_v[0] + t.b1[1] * g_v[1] + t.b1[2] * g_v[2]);
    float g_a2[3];
    for (int k = 0; k < 3; ++k) {
        g_b1[k] += -t.d * g_v[k] + g_d * a2[k];
        g_a2[k] = g_v[k] + g_d * t.b1[k];
    }
    /* b1 = normalize(a1) */
    float g_a1[3];
    frb_normalize3_bwd(t.b1, t.n1, 1e-6f, g_b1, g_a1);
    for (int k = 0; k < 3; ++k) { g_r6[k] = g_a1[k]; g_r6[3 + k] = g_a2[k]; }
}

/* Fresnel zone snap, GM:833-838 with FresnelZones.get_zone_centers_for_depth (utils/fresnel_zones.py:96-139):
 * clamp to the depth range, zone index = number of INTERIOR boundaries below the value (torch.bucketize, right =
 * False), result = that zone's centre.  boundaries: num_zones + 1 values, centers: num_zones values. */
FRB_HEAD_HD float frb_zone_center(float d, const float* boundaries, const float* centers, int num_zones) {
    const float lo = boundaries[0], hi = boundaries[num_zones];
    d = d < lo ? lo : (d > hi ? hi : d);
    int idx = 0;
    for (int k = 1; k < num_zones; ++k) idx += (boundaries[k] < d) ? 1 : 0;
    return centers[idx];
}

/* rotate_positions_for_pose, GM:51-104: Ry(azimuth) then Rx(elevation); trig = (cos az, sin az, cos el, sin el). */
FRB_HEAD_HD void frb_pose_rotate(const float p[3], const float trig[4], float out[3]) {
    const float x_rot = p[0] * trig[0] + p[2] * trig[1];
    const float z_rot = -p[0] * trig[1] + p[2] * trig[0];
    out[0] = x_rot;
    out[1] = p[1] * trig[2] - z_rot * trig[3];
    out[2] = p[1] * trig[3] + z_rot * trig[2];
}

/* gp = J^T g of frb_pose_rotate (the rotation is orthogonal: its transpose) */
FRB_HEAD_HD void frb_pose_rotate_bwd(const float g[3], const float trig[4], float gp[3]) {
    const float g_zrot = -g[1] * trig[3] + g[2] * trig[2];
    gp[0] = g[0] * trig[0] - g_zrot * trig[1];
    gp[1] = g[1] * trig[2] + g[2] * trig[3];
    gp[2] = g[0] * trig[1] + g_zrot * trig[0];
}

/* base_x, base_y: the patch-grid coordinate in [-1, 1]; z_base = depth_offset - 2 * depth_grid (GM:840);
 * edge: Fresnel edge strength of the patch (0 when edge-aware placement is off, GM:881-895). */
FRB_HEAD_HD void frb_head_fwd_one(const float raw[FRB_HEAD_RAW], float base_x, float base_y, float z_base, float edge,
                                  float edge_scale_factor, float edge_opacity_boost, FrbHeadOut& o) {
    o.pos[0] = base_x + raw[0] * 0.25f;                                       /* GM:846-850 */
    o.pos[1] = base_y + raw[1] * 0.25f;
    o.pos[2] = z_base;
    const float smod = 1.0f - edge_scale_factor * edge;
    for (int k = 0; k < 3; ++k) {
        float c = raw[3 + k];
        c = c < -10.0f ? -10.0f : (c > 20.0f ? 20.0f : c);                    /* GM:864 */
        const float x = c + 1.0f;
        const float sp = (x > 20.0f) ? x : log1pf(expf(x));                   /* F.softplus, threshold 20 */
        float s = sp * 0.15f;                                                 /* GM:865 */
        s = s < 1e-6f ? 1e-6f : (s > 2.0f ? 2.0f : s);                        /* GM:867 */
        o.scl[k] = s * smod;                                                  /* GM:888-889 */
        o.col[k] = frb_sigmoidf(raw[12 + k]);                                 /* GM:873 */
    }
    FrbRot6dTmp t;
    frb_rot6d_to_quat(raw + 6, o.rot, t);                                     /* GM:870 */
    float op = frb_sigmoidf(raw[15]);                                         /* GM:876 */
    if (edge_opacity_boost != 0.0f || edge != 0.0f) {
        op = op + edge_opacity_boost * edge;                                  /* GM:892-893 */
        op = op < 0.0f ? 0.0f : (op > 1.0f ? 1.0f : op);
    }
    o.opa = op;
}

/* g_raw[16] and the z gradient (the caller sums it into depth_offset's gradient). */
FRB_HEAD_HD void frb_head_bwd_one(const float raw[FRB_HEAD_RAW], float edge, float edge_scale_factor,
                                  float edge_opacity_boost, const FrbHeadOut& g, float g_raw[FRB_HEAD_RAW],
                                  float& g_z, float& g_edge) {
    g_edge = 0.0f;
    g_raw[0] = 0.25f * g.pos[0];
    g_raw[1] = 0.25f * g.pos[1];
    g_raw[2] = 0.0f;
    g_z = g.pos[2];
    const float smod = 1.0f - edge_scale_factor * edge;
    for (int k = 0; k < 3; ++k) {
        const float r = raw[3 + k];
        const float c = r < -10.0f ? -10.0f : (r > 20.0f ? 20.0f : r);
        const float x = c + 1.0f;
        const float sp = (x > 20.0f) ? x : log1pf(expf(x));
        const float s = sp * 0.15f;
        float gs = g.scl[k] * smod;
        {
            const float sc = s < 1e-6f ? 1e-6f : (s > 2.0f ? 2.0f : s);
            g_edge -= g.scl[k] * edge_scale_factor * sc;                      /* scale = sc * (1 - f * edge) */
        }
        if (!(s >= 1e-6f && s <= 2.0f)) gs = 0.0f;                            /* clamp backward, inclusive */
        float gx = gs * 0.15f * ((x > 20.0f) ? 1.0f : frb_sigmoidf(x));
        if (!(r >= -10.0f && r <= 20.0f)) gx = 0.0f;
        g_raw[3 + k] = gx;
        const float y = frb_sigmoidf(raw[12 + k]);
        g_raw[12 + k] = g.col[k] * y * (1.0f - y);
    }
    frb_rot6d_to_quat_bwd(raw + 6, g.rot, g_raw + 6);
    const float y = frb_sigmoidf(raw[15]);
    float go = g.opa;
    if (edge_opacity_boost != 0.0f || edge != 0.0f) {
        const float pre = y + edge_opacity_boost * edge;
        if (!(pre >= 0.0f && pre <= 1.0f)) go = 0.0f;
        g_edge += go * edge_opacity_boost;
    }
    g_raw[15] = go * y * (1.0f - y);
}

#endif /* FRB_HEAD_H */
