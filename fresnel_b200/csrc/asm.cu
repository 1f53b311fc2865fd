// Angular-spectrum propagation of the per-plane complex fields to the focal plane (cuFFT with a
// fused transfer-function multiply), the ASM renderer's epilogue, and their backward.
// Reference: AngularSpectrumPropagator DR:929-1065, ASMWaveFieldRenderer.forward DR:1136-1148
// (nearest depth plane), DR:1286-1332 (propagate, sum, intensity, normalise, background).
//
//   U_c = sum_p ifft2( fft2(f_{p,c}) * H_{p,c} ),   H = exp(i 2 pi z_p sqrt(max(1/l_c^2 - fx^2 - fy^2, 0)))
// The reference runs one fft2 + one ifft2 per (plane, channel).  ifft2 is linear, so this file
// runs ONE batched forward FFT over all planes and channels, multiplies by H and sums over planes
// in the frequency domain in one pass (H is generated on the fly, never stored), and runs one
// inverse FFT per channel: 16*3 + 3 transforms instead of 2*16*3.
// Backward: d f_{p,c} = ifft2( conj(H_{p,c}) * fft2(dU_c) )  (propagation by -z of the upstream field).
//
// Not reproduced: the reference skips planes whose field maximum is below 1e-8 (DR:1302); their
// contribution is below 1e-8 of the image scale.
#include <cufft.h>

#include <map>
#include <mutex>
#include <tuple>

#include "frb_common.cuh"

namespace {

constexpr int MAX_PLANES = 64;

struct AsmParams {
    int n_planes;
    float z[MAX_PLANES];        // focal_depth - plane_depth (fp32, as the reference computes it)
    float inv_wl[3];            // 1 / wavelength_c (fp32 division)
    float fscale_x, fscale_y;   // fftfreq: f = k * (float)(1 / (n * pitch))
};

struct PlaneSet {
    int n_planes;
    float depth[MAX_PLANES];
};

__global__ void __launch_bounds__(256)
asm_assign_planes_kernel(int n, const float4* __restrict__ records, const __grid_constant__ PlaneSet ps,
                         uint32_t* __restrict__ plane_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float depth = records[3 * (size_t)i + 2].w;
    float best = fabsf(__fsub_rn(depth, ps.depth[0]));      // DR:1147-1148: argmin |depth - plane|, first on ties
    uint32_t arg = 0;
    for (int p = 1; p < ps.n_planes; ++p) {
        float d = fabsf(__fsub_rn(depth, ps.depth[p]));
        if (d < best) { best = d; arg = p; }
    }
    plane_idx[i] = arg;
}

// phase of H at frequency bin (ky, kx) for plane p and channel c, in the reference's fp32 operation order
// (DR:990-999: kz_sq = (1/wl)**2 - FX**2 - FY**2; clamp; sqrt; exp(1j*2*pi*z*kz) = (2*pi*z)*kz).
__device__ __forceinline__ float asm_kz(const AsmParams& P, int c, int ky, int kx, int height, int width) {
    const int sx = (kx < (width + 1) / 2) ? kx : kx - width;      // torch.fft.fftfreq ordering
    const int sy = (ky < (height + 1) / 2) ? ky : ky - height;
    const float fx = __fmul_rn((float)sx, P.fscale_x), fy = __fmul_rn((float)sy, P.fscale_y);
    const float k0 = __fmul_rn(P.inv_wl[c], P.inv_wl[c]);
    float kz_sq = __fsub_rn(__fsub_rn(k0, __fmul_rn(fx, fx)), __fmul_rn(fy, fy));
    kz_sq = fmaxf(kz_sq, 0.0f);
    return __fsqrt_rn(kz_sq);
}

// out[view][c][f] = sum_p F[view][p][c][f] * H_{p,c}(f)
__global__ void __launch_bounds__(256)
asm_mul_sum_kernel(int n_views, int width, int height, const __grid_constant__ AsmParams P,
                   const float2* __restrict__ F, float2* __restrict__ out) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * 3 * hw) return;
    const size_t f = t % hw;
    const int c = (int)((t / hw) % 3);
    const int view = (int)(t / (3 * hw));
    const int ky = (int)(f / width), kx = (int)(f - (size_t)ky * width);
    const float kz = asm_kz(P, c, ky, kx, height, width);
    float2 acc = make_float2(0.f, 0.f);
    for (int p = 0; p < P.n_planes; ++p) {
        const float arg = __fmul_rn(__fmul_rn(6.2831855f, P.z[p]), kz);
        float sn, cn;
        sincosf(arg, &sn, &cn);
        const float2 v = F[(((size_t)view * P.n_planes + p) * 3 + c) * hw + f];
        acc.x += v.x * cn - v.y * sn;
        acc.y += v.x * sn + v.y * cn;
    }
    out[t] = acc;
}

// dF[view][p][c][f] = conj(H_{p,c}(f)) * G[view][c][f]
__global__ void __launch_bounds__(256)
asm_mul_conj_kernel(int n_views, int width, int height, const __grid_constant__ AsmParams P,
                    const float2* __restrict__ G, float2* __restrict__ dF) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * 3 * hw) return;
    const size_t f = t % hw;
    const int c = (int)((t / hw) % 3);
    const int view = (int)(t / (3 * hw));
    const int ky = (int)(f / width), kx = (int)(f - (size_t)ky * width);
    const float kz = asm_kz(P, c, ky, kx, height, width);
    const float2 g = G[t];
    for (int p = 0; p < P.n_planes; ++p) {
        const float arg = __fmul_rn(__fmul_rn(6.2831855f, P.z[p]), kz);
        float sn, cn;
        sincosf(arg, &sn, &cn);
        dF[(((size_t)view * P.n_planes + p) * 3 + c) * hw + f] = make_float2(g.x * cn + g.y * sn, g.y * cn - g.x * sn);
    }
}

// ---- epilogue DR:1316-1332.  total[view][3][H][W] holds the UNNORMALISED inverse FFT; inv_n = 1/(H*W).
__device__ __forceinline__ float asm_r(float re, float im) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)), 1e-8f));
}

__global__ void __launch_bounds__(256)
asm_max_kernel(int n_views, int width, int height, const float2* __restrict__ total, float inv_n,
               uint32_t* __restrict__ rmax_bits) {
    const size_t hw = (size_t)width * height;
    const int view = blockIdx.y;
    float rm = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * hw; i += (size_t)gridDim.x * blockDim.x) {
        float2 u = total[(size_t)view * 3 * hw + i];
        rm = fmaxf(rm, asm_r(u.x * inv_n, u.y * inv_n));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rmax_bits + view, __float_as_uint(rm));
}

__global__ void __launch_bounds__(256)
asm_finish_fwd_kernel(int n_views, int width, int height, const float2* __restrict__ total, float inv_n,
                      const uint32_t* __restrict__ rmax_bits, float3 bg, float* __restrict__ image) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float mx = fmaxf(__uint_as_float(rmax_bits[view]), 1.0f);
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float re[3], im[3], amp = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 u = total[((size_t)view * 3 + c) * hw + pix];
        re[c] = u.x * inv_n; im[c] = u.y * inv_n;
        amp += sqrtf(re[c] * re[c] + im[c] * im[c]);                 // |U_c|, DR:1327
    }
    float tot = fminf(fmaxf(amp, 0.f), 1.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float rend = fminf(fmaxf(asm_r(re[c], im[c]) / mx, 0.f), 1.f);
        image[(size_t)view * 3 * hw + c * hw + pix] = fminf(fmaxf(rend + bgc[c] * (1.0f - tot), 0.f), 1.f);
    }
}

__global__ void __launch_bounds__(256)
asm_finish_bwd_reduce_kernel(int n_views, int width, int height, const float2* __restrict__ total, float inv_n,
                             const uint32_t* __restrict__ rmax_bits, float3 bg, const float* __restrict__ g_image,
                             float* __restrict__ red) {
    const size_t hw = (size_t)width * height;
    const int view = blockIdx.y;
    const float rmax = __uint_as_float(rmax_bits[view]);
    const float mx = fmaxf(rmax, 1.0f);
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float S = 0.f, K = 0.f;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (size_t)gridDim.x * blockDim.x) {
        float re[3], im[3], amp = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 u = total[((size_t)view * 3 + c) * hw + pix];
            re[c] = u.x * inv_n; im[c] = u.y * inv_n;
            amp += sqrtf(re[c] * re[c] + im[c] * im[c]);
        }
        float tot = fminf(fmaxf(amp, 0.f), 1.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float r = asm_r(re[c], im[c]);
            float q = r / mx;
            float out = fminf(fmaxf(q, 0.f), 1.f) + bgc[c] * (1.0f - tot);
            float gi = (out >= 0.f && out <= 1.f) ? g_image[(size_t)view * 3 * hw + c * hw + pix] : 0.f;
            if (q >= 0.f && q <= 1.f) S += gi * r / (mx * mx);
            if (r == rmax) K += 1.f;
        }
    }
    S = frb_warp_sum(S);
    K = frb_warp_sum(K);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(red + 2 * view + 0, S);
        atomicAdd(red + 2 * view + 1, K);
    }
}

// g_total[view][3][H][W] float2 = dL/d(Re U), dL/d(Im U)
__global__ void __launch_bounds__(256)
asm_finish_bwd_kernel(int n_views, int width, int height, const float2* __restrict__ total, float inv_n,
                      const uint32_t* __restrict__ rmax_bits, float3 bg, const float* __restrict__ g_image,
                      const float* __restrict__ red, float2* __restrict__ g_total) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float rmax = __uint_as_float(rmax_bits[view]);
    const float mx = fmaxf(rmax, 1.0f);
    const float mxb = (rmax >= 1.0f) ? -red[2 * view] / fmaxf(red[2 * view + 1], 1.0f) : 0.f;
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float re[3], im[3], mag[3], amp = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 u = total[((size_t)view * 3 + c) * hw + pix];
        re[c] = u.x * inv_n; im[c] = u.y * inv_n;
        mag[c] = sqrtf(re[c] * re[c] + im[c] * im[c]);
        amp += mag[c];
    }
    float tot = fminf(fmaxf(amp, 0.f), 1.f);
    float totb = 0.f, rb[3], rr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float r = asm_r(re[c], im[c]);
        float q = r / mx;
        float out = fminf(fmaxf(q, 0.f), 1.f) + bgc[c] * (1.0f - tot);
        float gi = (out >= 0.f && out <= 1.f) ? g_image[(size_t)view * 3 * hw + c * hw + pix] : 0.f;
        totb -= gi * bgc[c];
        rb[c] = ((q >= 0.f && q <= 1.f) ? gi / mx : 0.f) + ((r == rmax) ? mxb : 0.f);
        rr[c] = r;
    }
    float ampb = (amp >= 0.f && amp <= 1.f) ? totb : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // r = sqrt(I + 1e-8): dI = rb / (2 r); |U| = sqrt(I): d(re, im) = ampb * (re, im) / |U| (0 at U = 0)
        float ib = rb[c] / (2.0f * rr[c]);
        float k = (mag[c] > 0.f) ? ampb / mag[c] : 0.f;
        g_total[((size_t)view * 3 + c) * hw + pix] = make_float2(re[c] * (2.0f * ib + k), im[c] * (2.0f * ib + k));
    }
}

// ---- cuFFT plan cache ------------------------------------------------------------------------
// A cuFFT handle owns ONE work area and ONE stream binding, so a handle shared by two streams (two HostRenderPipeline
// slots, two host threads) would race on both: the cache is keyed by the stream as well, and the stream is bound once,
// when the plan is made.  The cache is bounded: when it is full the least recently used plan is destroyed (cufftDestroy
// frees the work area, which synchronises the device - it only happens when a caller cycles through more than
// MAX_PLANS distinct (stream, size) combinations).
std::mutex g_plan_mutex;
struct PlanEntry { cufftHandle handle; unsigned long long last_use; };
std::map<std::tuple<int, cudaStream_t, int, int, int>, PlanEntry> g_plans;   // (device, stream, H, W, batch)
unsigned long long g_plan_clock = 0;
constexpr size_t MAX_PLANS = 64;

int get_plan(int height, int width, int batch, cudaStream_t st, cufftHandle* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto key = std::make_tuple(dev, st, height, width, batch);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        it->second.last_use = ++g_plan_clock;
        *out = it->second.handle;
        return 0;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap != cudaStreamCaptureStatusNone) {
        // no allocation inside a stream capture: re-bind a plan of the same shape made on another stream (the
        // warm-up before a capture usually ran on one), as the capture serialises its use anyway
        for (auto& kv : g_plans)
            if (std::get<0>(kv.first) == dev && std::get<2>(kv.first) == height && std::get<3>(kv.first) == width &&
                std::get<4>(kv.first) == batch) {
                if (cufftSetStream(kv.second.handle, st) != CUFFT_SUCCESS) return FRB_E_INVALID;
                PlanEntry e{kv.second.handle, ++g_plan_clock};
                g_plans.erase(kv.first);
                g_plans[key] = e;
                *out = e.handle;
                return 0;
            }
    }
    if (g_plans.size() >= MAX_PLANS && cap == cudaStreamCaptureStatusNone) {
        auto lru = g_plans.begin();
        for (auto j = g_plans.begin(); j != g_plans.end(); ++j)
            if (j->second.last_use < lru->second.last_use) lru = j;
        cufftDestroy(lru->second.handle);
        g_plans.erase(lru);
    }
    cufftHandle h;
    int n[2] = {height, width};
    if (cufftPlanMany(&h, 2, n, nullptr, 1, height * width, nullptr, 1, height * width, CUFFT_C2C, batch) !=
        CUFFT_SUCCESS)
        return FRB_E_INVALID;
    if (cufftSetStream(h, st) != CUFFT_SUCCESS) { cufftDestroy(h); return FRB_E_INVALID; }
    g_plans[key] = PlanEntry{h, ++g_plan_clock};
    *out = h;
    return 0;
}

int fill_params(int n_planes, const float* depth_planes_host, float focal_depth, float pixel_pitch,
                const float* wavelengths_host, int width, int height, AsmParams* P) {
    if (n_planes < 1 || n_planes > MAX_PLANES || !depth_planes_host || !wavelengths_host) return FRB_E_INVALID;
    P->n_planes = n_planes;
    for (int p = 0; p < n_planes; ++p) P->z[p] = focal_depth - depth_planes_host[p];       // DR:1293 (fp32)
    for (int c = 0; c < 3; ++c) P->inv_wl[c] = 1.0f / wavelengths_host[c];                  // DR:990
    P->fscale_x = (float)(1.0 / ((double)width * (double)pixel_pitch));                     // torch.fft.fftfreq
    P->fscale_y = (float)(1.0 / ((double)height * (double)pixel_pitch));
    return 0;
}

}  // namespace

extern "C" int frb_asm_assign_planes(int n, const float* records, int n_planes, const float* depth_planes_host,
                                     uint32_t* plane_idx, void* stream) {
    if (n < 0 || n_planes < 1 || n_planes > MAX_PLANES || !depth_planes_host) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!records || !plane_idx) return FRB_E_INVALID;
    PlaneSet ps;
    ps.n_planes = n_planes;
    for (int p = 0; p < n_planes; ++p) ps.depth[p] = depth_planes_host[p];
    asm_assign_planes_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, (const float4*)records, ps,
                                                                                 plane_idx);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// fields [view][plane][3][H][W] float2 (overwritten by its FFT); total [view][3][H][W] float2 (unnormalised
// inverse FFT of the summed spectrum); image [view][3][H][W].
extern "C" int frb_asm_propagate_fwd(int n_views, int width, int height, int n_planes,
                                     const float* depth_planes_host, float focal_depth, float pixel_pitch,
                                     const float* wavelengths_host, const float* background_host, float* fields,
                                     float* total, uint32_t* rmax_bits, float* image, void* stream) {
    if (n_views < 1 || width < 1 || height < 1 || !fields || !total || !rmax_bits || !image || !background_host)
        return FRB_E_INVALID;
    AsmParams P;
    int rc = fill_params(n_planes, depth_planes_host, focal_depth, pixel_pitch, wavelengths_host, width, height, &P);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cufftHandle big, small;
    if ((rc = get_plan(height, width, n_views * n_planes * 3, st, &big))) return rc;
    if ((rc = get_plan(height, width, n_views * 3, st, &small))) return rc;
    if (cufftExecC2C(big, (cufftComplex*)fields, (cufftComplex*)fields, CUFFT_FORWARD) != CUFFT_SUCCESS)
        return FRB_E_INVALID;
    const long long hw = (long long)width * height;
    asm_mul_sum_kernel<<<frb_div_up(hw * 3 * n_views, 256), 256, 0, st>>>(n_views, width, height, P,
                                                                           (const float2*)fields, (float2*)total);
    if (cufftExecC2C(small, (cufftComplex*)total, (cufftComplex*)total, CUFFT_INVERSE) != CUFFT_SUCCESS)
        return FRB_E_INVALID;
    const float inv_n = 1.0f / (float)hw;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    FRB_CUDA_OK(cudaMemsetAsync(rmax_bits, 0, sizeof(uint32_t) * n_views, st));
    dim3 grid(min(frb_div_up(3 * hw, 256), 592), n_views);
    asm_max_kernel<<<grid, 256, 0, st>>>(n_views, width, height, (const float2*)total, inv_n, rmax_bits);
    asm_finish_fwd_kernel<<<frb_div_up(hw * n_views, 256), 256, 0, st>>>(n_views, width, height,
                                                                          (const float2*)total, inv_n, rmax_bits, bg,
                                                                          image);
    frb_note_launches(3);
    FRB_LAUNCH_CHECK();
    return 0;
}

// g_total: scratch [view][3][H][W] float2; d_fields: [view][plane][3][H][W] float2, the UNNORMALISED inverse
// FFT of conj(H) fft2(dU) - multiply by 1/(H*W) when reading (frb_wave_splat_bwd's `scale`).
extern "C" int frb_asm_propagate_bwd(int n_views, int width, int height, int n_planes,
                                     const float* depth_planes_host, float focal_depth, float pixel_pitch,
                                     const float* wavelengths_host, const float* background_host,
                                     const float* total, const uint32_t* rmax_bits, const float* g_image, float* red,
                                     float* g_total, float* d_fields, void* stream) {
    if (n_views < 1 || width < 1 || height < 1 || !total || !rmax_bits || !g_image || !red || !g_total || !d_fields ||
        !background_host)
        return FRB_E_INVALID;
    AsmParams P;
    int rc = fill_params(n_planes, depth_planes_host, focal_depth, pixel_pitch, wavelengths_host, width, height, &P);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cufftHandle big, small;
    if ((rc = get_plan(height, width, n_views * n_planes * 3, st, &big))) return rc;
    if ((rc = get_plan(height, width, n_views * 3, st, &small))) return rc;
    const long long hw = (long long)width * height;
    const float inv_n = 1.0f / (float)hw;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    FRB_CUDA_OK(cudaMemsetAsync(red, 0, sizeof(float) * 2 * n_views, st));
    dim3 grid(min(frb_div_up(hw, 256), 592), n_views);
    asm_finish_bwd_reduce_kernel<<<grid, 256, 0, st>>>(n_views, width, height, (const float2*)total, inv_n, rmax_bits,
                                                       bg, g_image, red);
    asm_finish_bwd_kernel<<<frb_div_up(hw * n_views, 256), 256, 0, st>>>(
        n_views, width, height, (const float2*)total, inv_n, rmax_bits, bg, g_image, red, (float2*)g_total);
    if (cufftExecC2C(small, (cufftComplex*)g_total, (cufftComplex*)g_total, CUFFT_FORWARD) != CUFFT_SUCCESS)
        return FRB_E_INVALID;
    asm_mul_conj_kernel<<<frb_div_up(hw * 3 * n_views, 256), 256, 0, st>>>(n_views, width, height, P,
                                                                            (const float2*)g_total, (float2*)d_fields);
    if (cufftExecC2C(big, (cufftComplex*)d_fields, (cufftComplex*)d_fields, CUFFT_INVERSE) != CUFFT_SUCCESS)
        return FRB_E_INVALID;
    frb_note_launches(3);
    FRB_LAUNCH_CHECK();
    return 0;
}
