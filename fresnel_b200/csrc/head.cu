// Decoder output head as ONE kernel per direction: the raw MLP outputs go straight
// to the renderer's SoA parameter tensors, optionally only for the Gaussians kept by the stochastic
// subsampling (train_gaussian_decoder.py:1154-1187), and the backward scatters the 16 raw gradients back.
// Replaces about sixty elementwise PyTorch kernels forward and a hundred backward
// (DirectPatchDecoder.forward tail, scripts/models/gaussian_decoder_models.py:807-948).
// Arithmetic: frb_head.h.  HBM-bound: 64 B in / 56 B out per Gaussian forward.
#include "frb_common.cuh"
#include "frb_head.h"

namespace {

struct HeadShape {
    int B, H, W, K;      // batch, patch grid, Gaussians per patch
    int n_out;           // Gaussians per view that leave the head (n_sel or H*W*K)
};

constexpr int MAX_ZONES = 64;
struct HeadZones {       // Fresnel depth zones (0 = off), values as the reference's buffers hold them
    int num_zones;
    float boundaries[MAX_ZONES + 1];
    float centers[MAX_ZONES];
};

__device__ __forceinline__ void head_load_raw(const float4* __restrict__ raw4, long long row, float raw[16]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float4 v = raw4[row * 4 + k];
        raw[4 * k] = v.x; raw[4 * k + 1] = v.y; raw[4 * k + 2] = v.z; raw[4 * k + 3] = v.w;
    }
}

// base coordinates of patch-grid cell (h, w): torch.linspace(-1, 1, n)[i] (GM:807-811)
__device__ __forceinline__ float grid_coord(int i, int n) {
    return n > 1 ? -1.0f + (2.0f / (float)(n - 1)) * (float)i : -1.0f;
}

__global__ void __launch_bounds__(128)
head_fwd_kernel(HeadShape sh, const float4* __restrict__ raw4, const float* __restrict__ depth_grid,
                const float* __restrict__ depth_offset, const float* __restrict__ edge, float edge_scale_factor,
                float edge_opacity_boost, const __grid_constant__ HeadZones zones,
                const float4* __restrict__ pose_trig, const long long* __restrict__ idx,
                float* __restrict__ positions, float* __restrict__ scales, float4* __restrict__ rotations,
                float* __restrict__ colors, float* __restrict__ opacities) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)sh.B * sh.n_out) return;
    const int b = (int)(t / sh.n_out), j = (int)(t - (long long)b * sh.n_out);
    const int N = sh.H * sh.W * sh.K;
    const int i = idx ? (int)idx[j] : j;              // Gaussian index inside the view: (h, w, k) flattened
    const int cell = i / sh.K;
    const int h = cell / sh.W, w = cell - h * sh.W;
    float raw[16];
    head_load_raw(raw4, (long long)b * N + i, raw);
    float z = depth_offset[0];
    if (depth_grid) {
        float d = depth_grid[(long long)b * sh.H * sh.W + cell];
        if (zones.num_zones > 0) d = frb_zone_center(d, zones.boundaries, zones.centers, zones.num_zones);  // GM:833-838
        z += d * -2.0f;                                                                                    // GM:841
    }
    const float e = edge ? edge[(long long)b * sh.H * sh.W + cell] : 0.0f;
    FrbHeadOut o;
    frb_head_fwd_one(raw, grid_coord(w, sh.W), grid_coord(h, sh.H), z, e, edge ? edge_scale_factor : 0.0f,
                     edge ? edge_opacity_boost : 0.0f, o);
    if (pose_trig) {                                                                                       // GM:860
        const float4 tr = pose_trig[b];
        const float trig[4] = {tr.x, tr.y, tr.z, tr.w};
        float r[3];
        frb_pose_rotate(o.pos, trig, r);
        o.pos[0] = r[0]; o.pos[1] = r[1]; o.pos[2] = r[2];
    }
    positions[3 * t] = o.pos[0]; positions[3 * t + 1] = o.pos[1]; positions[3 * t + 2] = o.pos[2];
    scales[3 * t] = o.scl[0]; scales[3 * t + 1] = o.scl[1]; scales[3 * t + 2] = o.scl[2];
    rotations[t] = make_float4(o.rot[0], o.rot[1], o.rot[2], o.rot[3]);
    colors[3 * t] = o.col[0]; colors[3 * t + 1] = o.col[1]; colors[3 * t + 2] = o.col[2];
    opacities[t] = o.opa;
}

// g_raw is zero-filled by the caller when idx selects a subset; g_depth_offset accumulates (zeroed by the caller).
__global__ void __launch_bounds__(128)
head_bwd_kernel(HeadShape sh, const float4* __restrict__ raw4, const float* __restrict__ edge,
                float edge_scale_factor, float edge_opacity_boost, const float4* __restrict__ pose_trig,
                const long long* __restrict__ idx,
                const float* __restrict__ g_positions, const float* __restrict__ g_scales,
                const float4* __restrict__ g_rotations, const float* __restrict__ g_colors,
                const float* __restrict__ g_opacities, float4* __restrict__ g_raw4,
                float* __restrict__ g_depth_offset, float* __restrict__ g_edge) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float g_z = 0.0f;
    if (t < (long long)sh.B * sh.n_out) {
        const int b = (int)(t / sh.n_out), j = (int)(t - (long long)b * sh.n_out);
        const int N = sh.H * sh.W * sh.K;
        const int i = idx ? (int)idx[j] : j;
        const int cell = i / sh.K;
        float raw[16], g_raw[16];
        head_load_raw(raw4, (long long)b * N + i, raw);
        const float e = edge ? edge[(long long)b * sh.H * sh.W + cell] : 0.0f;
        FrbHeadOut g;
        g.pos[0] = g_positions ? g_positions[3 * t] : 0.f; g.pos[1] = g_positions ? g_positions[3 * t + 1] : 0.f;
        g.pos[2] = g_positions ? g_positions[3 * t + 2] : 0.f;
        g.scl[0] = g_scales ? g_scales[3 * t] : 0.f; g.scl[1] = g_scales ? g_scales[3 * t + 1] : 0.f;
        g.scl[2] = g_scales ? g_scales[3 * t + 2] : 0.f;
        float4 gr = g_rotations ? g_rotations[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        g.rot[0] = gr.x; g.rot[1] = gr.y; g.rot[2] = gr.z; g.rot[3] = gr.w;
        g.col[0] = g_colors ? g_colors[3 * t] : 0.f; g.col[1] = g_colors ? g_colors[3 * t + 1] : 0.f;
        g.col[2] = g_colors ? g_colors[3 * t + 2] : 0.f;
        g.opa = g_opacities ? g_opacities[t] : 0.f;
        if (pose_trig) {                          // positions left the head rotated: bring their gradient back
            const float4 tr = pose_trig[b];
            const float trig[4] = {tr.x, tr.y, tr.z, tr.w};
            float gp[3];
            frb_pose_rotate_bwd(g.pos, trig, gp);
            g.pos[0] = gp[0]; g.pos[1] = gp[1]; g.pos[2] = gp[2];
        }
        float g_e = 0.0f;
        frb_head_bwd_one(raw, e, edge ? edge_scale_factor : 0.0f, edge ? edge_opacity_boost : 0.0f, g, g_raw, g_z,
                         g_e);
        if (g_edge && edge && g_e != 0.0f) atomicAdd(g_edge + (long long)b * sh.H * sh.W + cell, g_e);
        const long long row = (long long)b * N + i;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            g_raw4[row * 4 + k] = make_float4(g_raw[4 * k], g_raw[4 * k + 1], g_raw[4 * k + 2], g_raw[4 * k + 3]);
    }
    g_z = frb_warp_sum(g_z);
    if ((threadIdx.x & 31) == 0 && g_z != 0.0f) atomicAdd(g_depth_offset, g_z);
}

}  // namespace

static int head_check(int B, int H, int W, int K, int n_sel) {
    if (B < 1 || H < 1 || W < 1 || K < 1 || n_sel < 0) return FRB_E_INVALID;
    if ((long long)B * H * W * K > 0x7fffffffLL || n_sel > H * W * K) return FRB_E_TOO_LARGE;
    return 0;
}

static int fill_zones(const FrbHeadExtras* ex, HeadZones* z) {
    z->num_zones = 0;
    if (!ex || ex->num_zones <= 0) return 0;
    if (ex->num_zones > MAX_ZONES || !ex->zone_boundaries_host || !ex->zone_centers_host) return FRB_E_INVALID;
    z->num_zones = ex->num_zones;
    for (int k = 0; k <= ex->num_zones; ++k) z->boundaries[k] = ex->zone_boundaries_host[k];
    for (int k = 0; k < ex->num_zones; ++k) z->centers[k] = ex->zone_centers_host[k];
    return 0;
}

extern "C" int frb_decode_head_fwd_ex(int B, int H, int W, int K, const float* raw, const float* depth_grid,
                                      const float* depth_offset, const FrbHeadExtras* ex, const long long* idx,
                                      int n_sel, float* positions, float* scales, float* rotations, float* colors,
                                      float* opacities, void* stream) {
    int rc = head_check(B, H, W, K, n_sel);
    if (rc) return rc;
    if (!raw || !depth_offset || !positions || !scales || !rotations || !colors || !opacities) return FRB_E_INVALID;
    if (frb_misaligned16(raw) || frb_misaligned16(rotations) || (ex && frb_misaligned16(ex->pose_trig)))
        return FRB_E_INVALID;
    HeadZones zones;
    if ((rc = fill_zones(ex, &zones))) return rc;
    HeadShape sh{B, H, W, K, idx ? n_sel : H * W * K};
    const long long total = (long long)B * sh.n_out;
    if (total == 0) return 0;
    head_fwd_kernel<<<frb_div_up(total, 128), 128, 0, (cudaStream_t)stream>>>(
        sh, (const float4*)raw, depth_grid, depth_offset, ex ? ex->edge : nullptr, ex ? ex->edge_scale_factor : 0.0f,
        ex ? ex->edge_opacity_boost : 0.0f, zones, ex ? (const float4*)ex->pose_trig : nullptr, idx, positions,
        scales, (float4*)rotations, colors, opacities);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_decode_head_bwd_ex(int B, int H, int W, int K, const float* raw, const FrbHeadExtras* ex,
                                      const long long* idx, int n_sel, const float* g_positions,
                                      const float* g_scales, const float* g_rotations, const float* g_colors,
                                      const float* g_opacities, float* g_raw, float* g_depth_offset, float* g_edge,
                                      void* stream) {
    int rc = head_check(B, H, W, K, n_sel);
    if (rc) return rc;
    if (!raw || !g_raw || !g_depth_offset) return FRB_E_INVALID;
    if (frb_misaligned16(raw) || frb_misaligned16(g_raw) || frb_misaligned16(g_rotations) ||
        (ex && frb_misaligned16(ex->pose_trig)))
        return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    HeadShape sh{B, H, W, K, idx ? n_sel : H * W * K};
    const long long total = (long long)B * sh.n_out;
    FRB_CUDA_OK(cudaMemsetAsync(g_depth_offset, 0, sizeof(float), st));
    if (g_edge) FRB_CUDA_OK(cudaMemsetAsync(g_edge, 0, sizeof(float) * (size_t)B * H * W, st));
    if (idx) FRB_CUDA_OK(cudaMemsetAsync(g_raw, 0, sizeof(float) * 16 * (size_t)B * H * W * K, st));
    if (total == 0) return 0;
    head_bwd_kernel<<<frb_div_up(total, 128), 128, 0, st>>>(
        sh, (const float4*)raw, ex ? ex->edge : nullptr, ex ? ex->edge_scale_factor : 0.0f,
        ex ? ex->edge_opacity_boost : 0.0f, ex ? (const float4*)ex->pose_trig : nullptr, idx, g_positions, g_scales,
        (const float4*)g_rotations, g_colors, g_opacities, (float4*)g_raw, g_depth_offset, g_edge);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_decode_head_fwd(int B, int H, int W, int K, const float* raw, const float* depth_grid,
                                   const float* depth_offset, const float* edge, float edge_scale_factor,
                                   float edge_opacity_boost, const long long* idx, int n_sel, float* positions,
                                   float* scales, float* rotations, float* colors, float* opacities, void* stream) {
    FrbHeadExtras ex = {edge, edge_scale_factor, edge_opacity_boost, 0, nullptr, nullptr, nullptr};
    return frb_decode_head_fwd_ex(B, H, W, K, raw, depth_grid, depth_offset, &ex, idx, n_sel, positions, scales,
                                  rotations, colors, opacities, stream);
}

extern "C" int frb_decode_head_bwd(int B, int H, int W, int K, const float* raw, const float* edge,
                                   float edge_scale_factor, float edge_opacity_boost, const long long* idx, int n_sel,
                                   const float* g_positions, const float* g_scales, const float* g_rotations,
                                   const float* g_colors, const float* g_opacities, float* g_raw,
                                   float* g_depth_offset, void* stream) {
    FrbHeadExtras ex = {edge, edge_scale_factor, edge_opacity_boost, 0, nullptr, nullptr, nullptr};
    return frb_decode_head_bwd_ex(B, H, W, K, raw, &ex, idx, n_sel, g_positions, g_scales, g_rotations, g_colors,
                                  g_opacities, g_raw, g_depth_offset, nullptr, stream);
}
