/*
 * Per-Gaussian and per-pixel arithmetic of the renderer, written once as
 * __host__ __device__ inline functions.  The CUDA kernels are the only product
 * callers; tests/host_shim.cpp compiles the same header with g++ so that the
 * hand-derived backward chain can be checked against the oracle's autograd in a
 * container without a GPU (it is a checker of this file, not a fallback path).
 *
 * Reference lines (DR = scripts/models/differentiable_renderer.py):
 *   projection     compute_2d_covariance DR:123-195, quaternion_to_rotation_matrix DR:98-120
 *   radius         TileBasedRenderer._compute_radius DR:452-487
 *   visibility     DR:541-543      rectangles DR:594-597      conic DR:578-579, 618
 *
 * Bit-exactness.  The oracle (oracle/fresnel_oracle.py: project) evaluates the
 * projection as elementwise fp32 operations in a fixed left-to-right order with
 * no fused multiply-add.  The forward functions here follow that order with
 * FRB_MUL / FRB_ADD / ... which are the round-to-nearest intrinsics on the
 * device (never contracted into FMA) and plain operators on the host (compile
 * the shim with -ffp-contract=off).  The backward functions are free to use FMA.
 */
#ifndef FRB_MATH_H
#define FRB_MATH_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FRB_HD __host__ __device__ __forceinline__
#else
#define FRB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FRB_MUL(a, b) __fmul_rn((a), (b))
#define FRB_ADD(a, b) __fadd_rn((a), (b))
#define FRB_SUB(a, b) __fsub_rn((a), (b))
#define FRB_DIV(a, b) __fdiv_rn((a), (b))
#define FRB_SQRT(a) __fsqrt_rn((a))
#else
#define FRB_MUL(a, b) ((a) * (b))
#define FRB_ADD(a, b) ((a) + (b))
#define FRB_SUB(a, b) ((a) - (b))
#define FRB_DIV(a, b) ((a) / (b))
#define FRB_SQRT(a) sqrtf((a))
#endif

#define FRB_LOG2E 1.4426950408889634f
#define FRB_LN2 0.6931471805599453f
#define FRB_CONIC_SCALE (-0.5f * FRB_LOG2E) /* exp(-0.5 m) = exp2(FRB_CONIC_SCALE * m) */
#define FRB_COV_EPS 1e-4f                   /* DR:578 */
#define FRB_ALPHA_MAX 0.99f                 /* DR:647 */

/* Projection modes: which renderer of differentiable_renderer.py the records are for.
 *   TILE    TileBasedRenderer / WaveFieldRenderer / ASMWaveFieldRenderer (DR:452-487, 541-543, 594-600)
 *   DENSE   DifferentiableGaussianRenderer DR:245-409: every visible Gaussian touches every pixel.  The
 *           rectangle is the box where exp(-0.5 m) can exceed 2^-40 (FRB_DENSE_SIGMAS standard deviations of
 *           the regularised covariance), clipped to the image; what is dropped is below 1e-12 per pair.
 *   FOURIER FourierGaussianRenderer DR:1500-1774: isotropic Gaussian with sigma^2 = (a + d)/2 + 1e-8,
 *           exp(-r^2 / (2 sigma^2 + 1e-8)), same support truncation. */
#define FRB_MODE_TILE 0
#define FRB_MODE_DENSE 1
#define FRB_MODE_FOURIER 2
#define FRB_DENSE_SIGMAS 7.5f               /* exp(-0.5 * 7.5^2) = 6e-13 */
#define FRB_DENSE_MARGIN 100.0f             /* DR:316-318 */

/* One view: rows 0..2 of the 4x4 world->camera matrix, intrinsics, image size, clip planes. */
struct FrbCamera {
    float V[12];
    float fx, fy, cx, cy;
    float width, height;
    float near_, far_;
};

/* Everything the projection of one Gaussian produces. */
struct FrbProjected {
    float u, v, depth;
    float a, b, c, d;      /* 2D covariance (row-major 2x2, un-regularised) */
    float radius;
    int visible;
    int x0, x1, y0, y1;    /* pixel rectangle [x0,x1) x [y0,y1) */
    float A, B, C;         /* conic of (cov + 1e-4 I), B = inv01 + inv10, pre-scaled by FRB_CONIC_SCALE */
};

/* torch.clamp(x, min=lo): NaN propagates. */
FRB_HD float frb_clamp_min(float x, float lo) { return (x < lo) ? lo : x; }
FRB_HD float frb_clamp_max(float x, float hi) { return (x > hi) ? hi : x; }
FRB_HD float frb_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* Intermediate values shared by the forward and backward projection. */
struct FrbProjTmp {
    float pcx, pcy, pcz;   /* camera-space centre */
    float qw, qx, qy, qz;  /* normalised quaternion */
    float den, nrm;        /* max(||q||, 1e-12), ||q|| */
    float RC[3][3];        /* V_rot @ R(q) */
    float M[3][3];         /* RC @ diag(s) */
    float S3[3][3];        /* M M^T */
    float zs;              /* z_safe DR:175 */
    float j00, j02, j11, j12;
};

FRB_HD void frb_project_core(const float p[3], const float s[3], const float q[4],
                             const FrbCamera& cam, FrbProjTmp& t, FrbProjected& o) {
    const float* V = cam.V;
    /* DR:149-152 */
    t.pcx = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[0], p[0]), FRB_MUL(V[1], p[1])), FRB_MUL(V[2], p[2])), V[3]);
    t.pcy = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[4], p[0]), FRB_MUL(V[5], p[1])), FRB_MUL(V[6], p[2])), V[7]);
    t.pcz = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[8], p[0]), FRB_MUL(V[9], p[1])), FRB_MUL(V[10], p[2])), V[11]);
    o.depth = -t.pcz; /* DR:155 */

    /* F.normalize DR:109 */
    float n2 = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(q[0], q[0]), FRB_MUL(q[1], q[1])), FRB_MUL(q[2], q[2])),
                       FRB_MUL(q[3], q[3]));
    t.nrm = (n2 > 0.0f) ? FRB_SQRT(n2) : ((n2 == 0.0f) ? 0.0f : n2 /* NaN */);
    t.den = frb_clamp_min(t.nrm, 1e-12f);
    float w = FRB_DIV(q[0], t.den), x = FRB_DIV(q[1], t.den), y = FRB_DIV(q[2], t.den),
          z = FRB_DIV(q[3], t.den);
    t.qw = w; t.qx = x; t.qy = y; t.qz = z;

    /* DR:114-118, same expression trees: 2*a*b == (2a)*b, 1 - X - Y == (1 - X) - Y */
    float R[3][3];
    float x2 = FRB_MUL(2.0f, x), y2 = FRB_MUL(2.0f, y), w2 = FRB_MUL(2.0f, w);
    R[0][0] = FRB_SUB(FRB_SUB(1.0f, FRB_MUL(y2, y)), FRB_MUL(FRB_MUL(2.0f, z), z));
    R[0][1] = FRB_SUB(FRB_MUL(x2, y), FRB_MUL(w2, z));
    R[0][2] = FRB_ADD(FRB_MUL(x2, z), FRB_MUL(w2, y));
    R[1][0] = FRB_ADD(FRB_MUL(x2, y), FRB_MUL(w2, z));
    R[1][1] = FRB_SUB(FRB_SUB(1.0f, FRB_MUL(x2, x)), FRB_MUL(FRB_MUL(2.0f, z), z));
    R[1][2] = FRB_SUB(FRB_MUL(y2, z), FRB_MUL(w2, x));
    R[2][0] = FRB_SUB(FRB_MUL(x2, z), FRB_MUL(w2, y));
    R[2][1] = FRB_ADD(FRB_MUL(y2, z), FRB_MUL(w2, x));
    R[2][2] = FRB_SUB(FRB_SUB(1.0f, FRB_MUL(x2, x)), FRB_MUL(y2, y));

    /* M = V_rot @ R @ diag(s), Sigma3 = M M^T   DR:162-166 */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float rc = FRB_ADD(FRB_ADD(FRB_MUL(V[4 * i + 0], R[0][j]), FRB_MUL(V[4 * i + 1], R[1][j])),
                               FRB_MUL(V[4 * i + 2], R[2][j]));
            t.RC[i][j] = rc;
            t.M[i][j] = FRB_MUL(rc, s[j]);
        }
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            float v = FRB_ADD(FRB_ADD(FRB_MUL(t.M[i][0], t.M[j][0]), FRB_MUL(t.M[i][1], t.M[j][1])),
                              FRB_MUL(t.M[i][2], t.M[j][2]));
            t.S3[i][j] = v;
            t.S3[j][i] = v;
        }

    /* DR:175-185 (sign quirk of J[1,2] kept) */
    float az = fabsf(t.pcz);
    t.zs = FRB_MUL(frb_clamp_min(az, 0.01f), frb_sign(FRB_ADD(t.pcz, 1e-8f)));
    float z2 = FRB_MUL(t.zs, t.zs);
    t.j00 = FRB_DIV(cam.fx, -t.zs);
    t.j02 = FRB_DIV(FRB_MUL(cam.fx, t.pcx), z2);
    t.j11 = FRB_DIV(cam.fy, t.zs);
    t.j12 = FRB_DIV(FRB_MUL(cam.fy, t.pcy), z2);

    /* T = J Sigma3, cov = T J^T   DR:188 */
    float t00 = FRB_ADD(FRB_MUL(t.j00, t.S3[0][0]), FRB_MUL(t.j02, t.S3[2][0]));
    float t01 = FRB_ADD(FRB_MUL(t.j00, t.S3[0][1]), FRB_MUL(t.j02, t.S3[2][1]));
    float t02 = FRB_ADD(FRB_MUL(t.j00, t.S3[0][2]), FRB_MUL(t.j02, t.S3[2][2]));
    float t10 = FRB_ADD(FRB_MUL(t.j11, t.S3[1][0]), FRB_MUL(t.j12, t.S3[2][0]));
    float t11 = FRB_ADD(FRB_MUL(t.j11, t.S3[1][1]), FRB_MUL(t.j12, t.S3[2][1]));
    float t12 = FRB_ADD(FRB_MUL(t.j11, t.S3[1][2]), FRB_MUL(t.j12, t.S3[2][2]));
    o.a = FRB_ADD(FRB_MUL(t00, t.j00), FRB_MUL(t02, t.j02));
    o.b = FRB_ADD(FRB_MUL(t01, t.j11), FRB_MUL(t02, t.j12));
    o.c = FRB_ADD(FRB_MUL(t10, t.j00), FRB_MUL(t12, t.j02));
    o.d = FRB_ADD(FRB_MUL(t11, t.j11), FRB_MUL(t12, t.j12));

    /* DR:191-192 */
    o.u = FRB_ADD(FRB_DIV(FRB_MUL(cam.fx, t.pcx), -t.zs), cam.cx);
    o.v = FRB_ADD(FRB_DIV(FRB_MUL(cam.fy, -t.pcy), -t.zs), cam.cy);
}

FRB_HD void frb_project_finish(const FrbCamera& cam, float max_radius, FrbProjected& o,
                                int mode = FRB_MODE_TILE) {
    float r;
    int vis = (o.depth > cam.near_) && (o.depth < cam.far_);
    if (mode == FRB_MODE_TILE) {
        /* radius from the UN-regularised covariance DR:452-487 */
        float trace = FRB_ADD(o.a, o.d);
        float det = frb_clamp_min(FRB_SUB(FRB_MUL(o.a, o.d), FRB_MUL(o.b, o.c)), 1e-6f);
        float disc = frb_clamp_min(FRB_SUB(FRB_MUL(trace, trace), FRB_MUL(4.0f, det)), 0.0f);
        float lam = FRB_DIV(FRB_ADD(trace, FRB_SQRT(disc)), 2.0f);
        r = FRB_MUL(3.0f, FRB_SQRT(frb_clamp_min(lam, 1e-6f)));
        r = frb_clamp_max(r, max_radius);
        /* DR:541-543, strict inequalities; NaN compares false */
        vis = vis && (FRB_ADD(o.u, r) > 0.0f) && (FRB_SUB(o.u, r) < cam.width);
        vis = vis && (FRB_ADD(o.v, r) > 0.0f) && (FRB_SUB(o.v, r) < cam.height);
    } else if (mode == FRB_MODE_DENSE) {
        /* DR:315-318: frustum and a 100-pixel margin on the centre; support box from the largest
         * eigenvalue of cov + 1e-4 I (the quadratic form is >= r^2 / lambda_max) */
        vis = vis && (o.u > -FRB_DENSE_MARGIN) && (o.u < cam.width + FRB_DENSE_MARGIN);
        vis = vis && (o.v > -FRB_DENSE_MARGIN) && (o.v < cam.height + FRB_DENSE_MARGIN);
        float ar = o.a + FRB_COV_EPS, dr = o.d + FRB_COV_EPS, bs = 0.5f * (o.b + o.c);
        float half_tr = 0.5f * (ar + dr), half_df = 0.5f * (ar - dr);
        float lam = half_tr + sqrtf(half_df * half_df + bs * bs);
        r = FRB_DENSE_SIGMAS * sqrtf(frb_clamp_min(lam, 0.0f)) + 1.0f;
        if (!(r < 32000.0f)) r = 32000.0f;         /* also NaN: the rectangle becomes the whole image */
    } else {
        /* DR:1647-1649: frustum and one image size of margin on the centre */
        vis = vis && (o.u > -cam.width) && (o.u < 2.0f * cam.width);
        vis = vis && (o.v > -cam.height) && (o.v < 2.0f * cam.height);
        float var = FRB_ADD(FRB_DIV(FRB_ADD(o.a, o.d), 2.0f), 1e-8f);        /* sigma^2  DR:1677 */
        r = FRB_DENSE_SIGMAS * sqrtf(frb_clamp_min(var + 0.5e-8f, 0.0f)) + 1.0f;
        if (!(r < 32000.0f)) r = 32000.0f;
    }
    o.radius = r;
    o.visible = vis;

    /* DR:594-597: Python floats (fp64) from .item(), int() truncation */
    o.x0 = o.x1 = o.y0 = o.y1 = 0;
    if (vis) {
        double u = (double)o.u, v = (double)o.v, rr = (double)r;
        double W = (double)cam.width, H = (double)cam.height;
        double x0 = fmax(0.0, trunc(u - rr)), x1 = fmin(W, trunc(u + rr) + 1.0);
        double y0 = fmax(0.0, trunc(v - rr)), y1 = fmin(H, trunc(v + rr) + 1.0);
        o.x0 = (int)x0; o.x1 = (int)x1; o.y0 = (int)y0; o.y1 = (int)y1;
        if (o.x0 >= o.x1 || o.y0 >= o.y1) { o.x0 = o.x1 = o.y0 = o.y1 = 0; } /* DR:599: skipped */
    }

    if (mode == FRB_MODE_FOURIER) {
        /* exp(-r^2 / (2 sigma^2 + 1e-8)) = exp2(A r^2)   DR:1725 */
        float var = FRB_ADD(FRB_DIV(FRB_ADD(o.a, o.d), 2.0f), 1e-8f);
        float sig = FRB_SQRT(var);                                           /* the reference squares the root */
        float den = FRB_ADD(FRB_MUL(2.0f, FRB_MUL(sig, sig)), 1e-8f);
        o.A = -FRB_LOG2E / den;
        o.B = 0.0f;
        o.C = o.A;
        return;
    }
    /* pinv(cov + 1e-4 I) for a full-rank 2x2 is the inverse DR:578-579; cross term DR:618 */
    float ar = o.a + FRB_COV_EPS, dr = o.d + FRB_COV_EPS;
    float detr = ar * dr - o.b * o.c;
    float inv = 1.0f / detr;
    o.A = FRB_CONIC_SCALE * (dr * inv);
    o.B = FRB_CONIC_SCALE * (-(o.b + o.c) * inv);
    o.C = FRB_CONIC_SCALE * (ar * inv);
}

/*
 * Backward of the projection: the autograd tape through DR:123-195 and DR:578-579.
 * Inputs are the gradients with respect to the record fields the compositor reads:
 * g_u, g_v, g_A, g_B, g_C (pre-scaled conic), g_depth.  Radius, rectangle, order and
 * visibility carry no gradient in the reference (.item() / integer / boolean).
 */
FRB_HD void frb_project_bwd_one(const float p[3], const float s[3], const float q[4],
                                const FrbCamera& cam, float g_u, float g_v, float g_A, float g_B,
                                float g_C, float g_depth, float gp[3], float gs[3], float gq[4],
                                int mode = FRB_MODE_TILE) {
    FrbProjTmp t;
    FrbProjected o;
    frb_project_core(p, s, q, cam, t, o);
    const float* V = cam.V;

    float ga, gb, gc, gd;
    if (mode == FRB_MODE_FOURIER) {
        /* A = C = -log2e / (2 sigma^2 + 1e-8), sigma^2 = (a + d)/2 + 1e-8: dA/da = dA/dd = log2e / den^2 */
        float var = (o.a + o.d) * 0.5f + 1e-8f;
        float den = 2.0f * var + 1e-8f;
        float gvar = (g_A + g_C) * 2.0f * FRB_LOG2E / (den * den);
        ga = gd = 0.5f * gvar;
        gb = gc = 0.0f;
    } else {
        /* conic -> regularised covariance X = cov + eps I; Y = X^-1; dX = -Y^T G Y^T */
        float ar = o.a + FRB_COV_EPS, dr = o.d + FRB_COV_EPS;
        float inv = 1.0f / (ar * dr - o.b * o.c);
        float Y00 = dr * inv, Y01 = -o.b * inv, Y10 = -o.c * inv, Y11 = ar * inv;
        float G00 = FRB_CONIC_SCALE * g_A, G01 = FRB_CONIC_SCALE * g_B, G10 = G01, G11 = FRB_CONIC_SCALE * g_C;
        /* P = Y^T G */
        float P00 = Y00 * G00 + Y10 * G10, P01 = Y00 * G01 + Y10 * G11;
        float P10 = Y01 * G00 + Y11 * G10, P11 = Y01 * G01 + Y11 * G11;
        /* dX = -(P Y^T) */
        ga = -(P00 * Y00 + P01 * Y01); gb = -(P00 * Y10 + P01 * Y11);
        gc = -(P10 * Y00 + P11 * Y01); gd = -(P10 * Y10 + P11 * Y11);
    }

    /* cov = J S3 J^T with J = [[j00, 0, j02], [0, j11, j12]] */
    float J[2][3] = {{t.j00, 0.0f, t.j02}, {0.0f, t.j11, t.j12}};
    float G[2][2] = {{ga, gb}, {gc, gd}};
    /* dS3 = J^T G J  (3x3) */
    float dS3[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 2; ++k)
                for (int l = 0; l < 2; ++l) acc += J[k][i] * G[k][l] * J[l][j];
            dS3[i][j] = acc;
        }
    /* dJ = (G + G^T) J S3  (S3 symmetric) */
    float JS[2][3];
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 3; ++j)
            JS[k][j] = J[k][0] * t.S3[0][j] + J[k][1] * t.S3[1][j] + J[k][2] * t.S3[2][j];
    float Gs[2][2] = {{2.0f * ga, gb + gc}, {gb + gc, 2.0f * gd}};
    float dJ[2][3];
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < 3; ++j) dJ[k][j] = Gs[k][0] * JS[0][j] + Gs[k][1] * JS[1][j];

    /* J and (u, v) -> camera-space point */
    float zs = t.zs, iz = 1.0f / zs, iz2 = iz * iz, iz3 = iz2 * iz;
    float fx = cam.fx, fy = cam.fy;
    float g_pcx = dJ[0][2] * fx * iz2 - g_u * fx * iz;
    float g_pcy = dJ[1][2] * fy * iz2 + g_v * fy * iz;
    float g_zs = dJ[0][0] * fx * iz2 - 2.0f * dJ[0][2] * fx * t.pcx * iz3 - dJ[1][1] * fy * iz2 -
                 2.0f * dJ[1][2] * fy * t.pcy * iz3 + g_u * fx * t.pcx * iz2 - g_v * fy * t.pcy * iz2;
    /* z_safe = clamp(|z|, 0.01) * sign(z + 1e-8): clamp gate is inclusive, sign() has no gradient */
    float gate = (fabsf(t.pcz) >= 0.01f) ? 1.0f : 0.0f;
    float g_pcz = g_zs * gate * frb_sign(t.pcz) * frb_sign(t.pcz + 1e-8f) - g_depth;
    gp[0] = V[0] * g_pcx + V[4] * g_pcy + V[8] * g_pcz;
    gp[1] = V[1] * g_pcx + V[5] * g_pcy + V[9] * g_pcz;
    gp[2] = V[2] * g_pcx + V[6] * g_pcy + V[10] * g_pcz;

    /* S3 = M M^T -> dM = (dS3 + dS3^T) M ; M = RC diag(s) */
    float dM[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) acc += (dS3[i][k] + dS3[k][i]) * t.M[k][j];
            dM[i][j] = acc;
        }
    float dRC[3][3];
    for (int j = 0; j < 3; ++j) {
        gs[j] = dM[0][j] * t.RC[0][j] + dM[1][j] * t.RC[1][j] + dM[2][j] * t.RC[2][j];
        for (int i = 0; i < 3; ++i) dRC[i][j] = dM[i][j] * s[j];
    }
    /* RC = V_rot R -> dR = V_rot^T dRC */
    float dR[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            dR[i][j] = V[0 + i] * dRC[0][j] + V[4 + i] * dRC[1][j] + V[8 + i] * dRC[2][j];

    /* R(q^) DR:114-118 */
    float w = t.qw, x = t.qx, y = t.qy, z = t.qz;
    float gw = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
    float gx = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - w * dR[1][2] +
                       z * dR[2][0] + w * dR[2][1] - 2.0f * x * dR[2][2]);
    float gy = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + w * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                       w * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
    float gz = 2.0f * (-2.0f * z * dR[0][0] - w * dR[0][1] + x * dR[0][2] + w * dR[1][0] - 2.0f * z * dR[1][1] +
                       y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
    /* q^ = q / max(||q||, 1e-12) */
    float iden = 1.0f / t.den;
    if (t.nrm >= 1e-12f) {
        float dot = w * gw + x * gx + y * gy + z * gz;
        gq[0] = (gw - w * dot) * iden;
        gq[1] = (gx - x * dot) * iden;
        gq[2] = (gy - y * dot) * iden;
        gq[3] = (gz - z * dot) * iden;
    } else {
        gq[0] = gw * iden; gq[1] = gx * iden; gq[2] = gy * iden; gq[3] = gz * iden;
    }
}

#endif /* FRB_MATH_H */
