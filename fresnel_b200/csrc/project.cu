// EWA projection of 3D Gaussians to screen-space records, and its backward.
// Reference: compute_2d_covariance DR:123-195, _compute_radius DR:452-487, visibility and
// rectangles DR:541-543 / 594-600, pinv(cov + 1e-4 I) DR:578-579.  The arithmetic lives in
// frb_math.h (bit-exact against oracle/fresnel_oracle.py: project / compute_radius / rects).
//
// HBM-bound, one thread per Gaussian: 56 B in (+4 B opacity), 48 B record + 8 B sort inputs out.
#include "frb_common.cuh"

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
int frb_record_tensor_map(const float* records, int n, CUtensorMap* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return (int)e;
        if (!fn) return FRB_E_INVALID;
        encode = (EncodeFn)fn;
    }
    if (!records || n < 1 || frb_misaligned16(records)) return FRB_E_INVALID;
    cuuint64_t gdim[2] = {FRB_RECORD_FLOATS, (cuuint64_t)n};
    cuuint64_t gstride[1] = {FRB_RECORD_FLOATS * sizeof(float)};
    cuuint32_t box[2] = {16, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)records, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : FRB_E_INVALID;
}

void frb_note_launches(int k) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }
extern "C" unsigned long long frb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

#include <cstdlib>
bool frb_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("FRB_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

int frb_fill_views(int n, int n_views, const float* camera_host, FrbViewSet* vs) {
    if (n < 0 || n_views < 1 || n_views > FRB_MAX_VIEWS || camera_host == nullptr) return FRB_E_INVALID;
    if (n % n_views != 0) return FRB_E_INVALID;
    vs->n_views = n_views;
    vs->n_per_view = n / n_views;
    for (int k = 0; k < n_views; ++k) {
        const float* c = camera_host + k * FRB_CAMERA_FLOATS;
        FrbCamera& cam = vs->cam[k];
        for (int i = 0; i < 12; ++i) cam.V[i] = c[i];
        cam.fx = c[12]; cam.fy = c[13]; cam.cx = c[14]; cam.cy = c[15];
        cam.width = c[16]; cam.height = c[17]; cam.near_ = c[18]; cam.far_ = c[19];
        if (c[16] != camera_host[16] || c[17] != camera_host[17]) return FRB_E_INVALID;
        if (c[16] < 1.0f || c[17] < 1.0f) return FRB_E_INVALID;
        if (c[16] > (float)FRB_MAX_IMAGE_SIDE || c[17] > (float)FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    }
    return 0;
}

__global__ void __launch_bounds__(256)
frb_project_fwd_kernel(int n, const __grid_constant__ FrbViewSet vs, const float* __restrict__ positions,
                       const float* __restrict__ scales, const float* __restrict__ rotations,
                       const float* __restrict__ colors, const float* __restrict__ opacities,
                       float max_radius, int mode, float4* __restrict__ records, int4* __restrict__ rects,
                       uint32_t* __restrict__ depth_bits, uint32_t* __restrict__ tiles_touched,
                       float4* __restrict__ debug) {
    frb_pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FrbCamera& cam = vs.cam[i / vs.n_per_view];
    float p[3] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
    float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    float4 q4 = reinterpret_cast<const float4*>(rotations)[i];
    float q[4] = {q4.x, q4.y, q4.z, q4.w};

    FrbProjTmp t;
    FrbProjected o;
    frb_project_core(p, s, q, cam, t, o);
    frb_project_finish(cam, max_radius, o, mode);

    uint32_t touched = 0;
    if (o.visible && o.x1 > o.x0) {
        int tx0 = o.x0 / FRB_TILE, tx1 = (o.x1 - 1) / FRB_TILE;
        int ty0 = o.y0 / FRB_TILE, ty1 = (o.y1 - 1) / FRB_TILE;
        touched = (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1));
    }
    uint32_t rect_lo = (uint32_t)o.x0 | ((uint32_t)o.y0 << 16);
    uint32_t rect_hi = (uint32_t)o.x1 | ((uint32_t)o.y1 << 16) | 0x80008000u;

    records[3 * i + 0] = make_float4(o.u, o.v, o.A, o.B);
    records[3 * i + 1] = make_float4(o.C, opacities[i], __uint_as_float(rect_lo), __uint_as_float(rect_hi));
    records[3 * i + 2] = make_float4(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2], o.depth);
    depth_bits[i] = __float_as_uint(o.depth);
    tiles_touched[i] = touched;
    if (rects) rects[i] = make_int4(o.x0, o.x1, o.y0, o.y1);
    if (debug) {
        debug[2 * i + 0] = make_float4(o.a, o.b, o.c, o.d);
        debug[2 * i + 1] = make_float4(o.radius, o.visible ? 1.0f : 0.0f, 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(256)
frb_project_bwd_kernel(int n, const __grid_constant__ FrbViewSet vs, const float* __restrict__ positions,
                       const float* __restrict__ scales, const float* __restrict__ rotations,
                       const float4* __restrict__ grad2d, int mode, float* __restrict__ g_positions,
                       float* __restrict__ g_scales, float4* __restrict__ g_rotations,
                       float* __restrict__ g_colors, float* __restrict__ g_opacities) {
    frb_pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 g0 = grad2d[3 * i + 0], g1 = grad2d[3 * i + 1], g2 = grad2d[3 * i + 2];
    if (g_colors) {
        g_colors[3 * i] = g2.x; g_colors[3 * i + 1] = g2.y; g_colors[3 * i + 2] = g2.z;
    }
    if (g_opacities) g_opacities[i] = g1.y;

    float gp[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
    // A culled Gaussian never reaches the compositor: its gradient is exactly zero (DR:554-562).
    bool any = (g0.x != 0.f) | (g0.y != 0.f) | (g0.z != 0.f) | (g0.w != 0.f) | (g1.x != 0.f) | (g1.z != 0.f);
    if (any) {
        const FrbCamera& cam = vs.cam[i / vs.n_per_view];
        float p[3] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
        float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
        float4 q4 = reinterpret_cast<const float4*>(rotations)[i];
        float q[4] = {q4.x, q4.y, q4.z, q4.w};
        frb_project_bwd_one(p, s, q, cam, g0.x, g0.y, g0.z, g0.w, g1.x, g1.z, gp, gs, gq, mode);
    }
    g_positions[3 * i] = gp[0]; g_positions[3 * i + 1] = gp[1]; g_positions[3 * i + 2] = gp[2];
    g_scales[3 * i] = gs[0]; g_scales[3 * i + 1] = gs[1]; g_scales[3 * i + 2] = gs[2];
    g_rotations[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
}

extern "C" int frb_project_fwd(int n, int n_views, const float* positions, const float* scales,
                               const float* rotations, const float* colors, const float* opacities,
                               const float* camera_host, float max_radius, float* records,
                               int32_t* rects, uint32_t* depth_bits, uint32_t* tiles_touched,
                               float* debug, void* stream) {
    return frb_project_fwd_mode(n, n_views, positions, scales, rotations, colors, opacities, camera_host,
                                max_radius, FRB_MODE_TILE, records, rects, depth_bits, tiles_touched, debug, stream);
}

extern "C" int frb_project_fwd_mode(int n, int n_views, const float* positions, const float* scales,
                                    const float* rotations, const float* colors, const float* opacities,
                                    const float* camera_host, float max_radius, int mode, float* records,
                                    int32_t* rects, uint32_t* depth_bits, uint32_t* tiles_touched,
                                    float* debug, void* stream) {
    if (mode < FRB_MODE_TILE || mode > FRB_MODE_FOURIER) return FRB_E_INVALID;
    FrbViewSet vs;
    int rc = frb_fill_views(n, n_views, camera_host, &vs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!positions || !scales || !rotations || !colors || !opacities || !records || !depth_bits ||
        !tiles_touched)
        return FRB_E_INVALID;
    // read / written as 16-byte vectors: a misaligned pointer would fault on the device (sticky error)
    if (frb_misaligned16(rotations) || frb_misaligned16(records) || frb_misaligned16(rects) || frb_misaligned16(debug))
        return FRB_E_INVALID;
    frb_launch(frb_project_fwd_kernel, dim3(frb_div_up(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
        n, vs, positions, scales, rotations, colors, opacities, max_radius, mode, (float4*)records,
        (int4*)rects, depth_bits, tiles_touched, (float4*)debug);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_project_bwd(int n, int n_views, const float* positions, const float* scales,
                               const float* rotations, const float* camera_host, const float* grad2d,
                               float* g_positions, float* g_scales, float* g_rotations,
                               float* g_colors, float* g_opacities, void* stream) {
    return frb_project_bwd_mode(n, n_views, positions, scales, rotations, camera_host, grad2d, FRB_MODE_TILE,
                                g_positions, g_scales, g_rotations, g_colors, g_opacities, stream);
}

extern "C" int frb_project_bwd_mode(int n, int n_views, const float* positions, const float* scales,
                                    const float* rotations, const float* camera_host, const float* grad2d,
                                    int mode, float* g_positions, float* g_scales, float* g_rotations,
                                    float* g_colors, float* g_opacities, void* stream) {
    if (mode < FRB_MODE_TILE || mode > FRB_MODE_FOURIER) return FRB_E_INVALID;
    FrbViewSet vs;
    int rc = frb_fill_views(n, n_views, camera_host, &vs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!positions || !scales || !rotations || !grad2d || !g_positions || !g_scales || !g_rotations)
        return FRB_E_INVALID;
    if (frb_misaligned16(rotations) || frb_misaligned16(grad2d) || frb_misaligned16(g_rotations)) return FRB_E_INVALID;
    frb_launch(frb_project_bwd_kernel, dim3(frb_div_up(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
        n, vs, positions, scales, rotations, (const float4*)grad2d, mode, g_positions, g_scales,
        (float4*)g_rotations, g_colors, g_opacities);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_version(void) { return 100; }

extern "C" const char* frb_error_string(int code) {
    if (code == 0) return "ok";
    if (code == FRB_E_INVALID) return "fresnel_b200: invalid argument";
    if (code == FRB_E_TOO_LARGE) return "fresnel_b200: problem too large";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "fresnel_b200: unknown error";
}
