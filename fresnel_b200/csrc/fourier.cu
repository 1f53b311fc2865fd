// FourierGaussianRenderer epilogue and its backward.
// Reference: FourierGaussianRenderer.forward DR:1582-1766.  Despite its name the reference accumulates
// isotropic Gaussians in the SPATIAL domain (DR:1693-1738: "Use REAL spatial accumulation"):
//     image_c = sum_i colour_ic * opacity_i * exp(-r^2 / (2 sigma_i^2 + 1e-8)),  sigma_i^2 = (a + d)/2 + 1e-8
// which is an order-free additive splat: the accumulation itself runs through the wave splat kernels
// (wave.cu) with all phases zero (Re_c = the sum above, Im_c = 0) on records projected in
// FRB_MODE_FOURIER.  This file holds what follows the accumulation (DR:1740-1753):
//     mx = max(image); if mx > 1e-8: image /= mx
//     image += bg * clamp(1 - sum_c image_c, 0, 1);  image = clamp(image, 0, 1)
// and the backward of exactly that (torch.max() sends its gradient to the arg max, split evenly on ties).
#include "frb_common.cuh"

namespace {

__device__ __forceinline__ uint32_t ordered_key(float x) {          // monotone float -> uint32
    uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b ^ 0x80000000u);
}
__device__ __forceinline__ float ordered_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// accum: [view][8][H][W] (planes 0..2 = the three channel sums).  mx_key[view] zero-initialised.
__global__ void __launch_bounds__(256)
fourier_max_kernel(int width, int height, const float* __restrict__ accum, uint32_t* __restrict__ mx_key) {
    const size_t hw = (size_t)width * height;
    const int view = blockIdx.y;
    const float* a = accum + (size_t)view * 8 * hw;
    float m = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * hw; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(mx_key + view, ordered_key(m));
}

__global__ void __launch_bounds__(256)
fourier_finish_fwd_kernel(int n_views, int width, int height, const float* __restrict__ accum,
                          const uint32_t* __restrict__ mx_key, float3 bg, float* __restrict__ image) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float* a = accum + (size_t)view * 8 * hw + pix;
    const float mx = ordered_value(mx_key[view]);
    const bool norm = mx > 1e-8f;                                            // DR:1742
    float y[3] = {a[0], a[hw], a[2 * hw]};
    if (norm) { y[0] = y[0] / mx; y[1] = y[1] / mx; y[2] = y[2] / mx; }
    const float tot = (y[0] + y[1]) + y[2];                                  // DR:1746
    const float bgw = fminf(fmaxf(1.0f - tot, 0.f), 1.f);                    // DR:1747
    const float bgc[3] = {bg.x, bg.y, bg.z};
#pragma unroll
    for (int c = 0; c < 3; ++c)
        image[(size_t)view * 3 * hw + c * hw + pix] = fminf(fmaxf(y[c] + bgc[c] * bgw, 0.f), 1.f);   // DR:1748-1751
}

// dL/dy_c for one pixel (y = normalised sums), shared by the two backward kernels.
__device__ __forceinline__ void fourier_pixel_bwd(const float y[3], const float gi[3], const float bgc[3], float gy[3]) {
    const float tot = (y[0] + y[1]) + y[2];
    const float w0 = 1.0f - tot;
    const float bgw = fminf(fmaxf(w0, 0.f), 1.f);
    float g_bgw = 0.f, go[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float pre = y[c] + bgc[c] * bgw;
        go[c] = (pre >= 0.f && pre <= 1.f) ? gi[c] : 0.f;                    // clamp backward, inclusive
        g_bgw += go[c] * bgc[c];
    }
    const float g_tot = (w0 >= 0.f && w0 <= 1.f) ? -g_bgw : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) gy[c] = go[c] + g_tot;
}

// red[view] = { S = sum gy * x / mx^2 , K = number of elements equal to the maximum }
__global__ void __launch_bounds__(256)
fourier_bwd_reduce_kernel(int width, int height, const float* __restrict__ accum,
                          const uint32_t* __restrict__ mx_key, float3 bg, const float* __restrict__ g_image,
                          float* __restrict__ red) {
    const size_t hw = (size_t)width * height;
    const int view = blockIdx.y;
    const float mx = ordered_value(mx_key[view]);
    if (!(mx > 1e-8f)) return;
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float S = 0.f, K = 0.f;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (size_t)gridDim.x * blockDim.x) {
        const float* a = accum + (size_t)view * 8 * hw + pix;
        const float x[3] = {a[0], a[hw], a[2 * hw]};
        const float y[3] = {x[0] / mx, x[1] / mx, x[2] / mx};
        float gi[3], gy[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) gi[c] = g_image[(size_t)view * 3 * hw + c * hw + pix];
        fourier_pixel_bwd(y, gi, bgc, gy);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            S += gy[c] * x[c] / (mx * mx);
            if (x[c] == mx) K += 1.f;
        }
    }
    S = frb_warp_sum(S);
    K = frb_warp_sum(K);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(red + 2 * view + 0, S);
        atomicAdd(red + 2 * view + 1, K);
    }
}

// gpix: [view][8][H][W]; planes 0..2 = dL/d(channel sums), planes 3..7 = 0 (no imaginary part, depth or weight)
__global__ void __launch_bounds__(256)
fourier_finish_bwd_kernel(int n_views, int width, int height, const float* __restrict__ accum,
                          const uint32_t* __restrict__ mx_key, float3 bg, const float* __restrict__ g_image,
                          const float* __restrict__ red, float* __restrict__ gpix) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float* a = accum + (size_t)view * 8 * hw + pix;
    const float mx = ordered_value(mx_key[view]);
    const bool norm = mx > 1e-8f;
    const float bgc[3] = {bg.x, bg.y, bg.z};
    const float x[3] = {a[0], a[hw], a[2 * hw]};
    float y[3] = {x[0], x[1], x[2]};
    if (norm) { y[0] = x[0] / mx; y[1] = x[1] / mx; y[2] = x[2] / mx; }
    float gi[3], gy[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gi[c] = g_image[(size_t)view * 3 * hw + c * hw + pix];
    fourier_pixel_bwd(y, gi, bgc, gy);
    const float share = norm ? -red[2 * view] / fmaxf(red[2 * view + 1], 1.0f) : 0.f;   // to each arg max
    float* g = gpix + (size_t)view * 8 * hw + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c * hw] = (norm ? gy[c] / mx : gy[c]) + ((norm && x[c] == mx) ? share : 0.f);
#pragma unroll
    for (int c = 3; c < 8; ++c) g[c * hw] = 0.f;
}

}  // namespace

static int check_views(int n_views, int width, int height) {
    if (n_views < 1 || n_views > FRB_MAX_VIEWS || width < 1 || height < 1) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    return 0;
}

extern "C" int frb_fourier_finish_fwd(int n_views, int width, int height, const float* accum, uint32_t* mx_key,
                                      const float* background_host, float* image, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!accum || !mx_key || !background_host || !image) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    const long long hw = (long long)width * height;
    FRB_CUDA_OK(cudaMemsetAsync(mx_key, 0, sizeof(uint32_t) * n_views, st));
    dim3 grid(min(frb_div_up(3 * hw, 256), 592), n_views);
    fourier_max_kernel<<<grid, 256, 0, st>>>(width, height, accum, mx_key);
    fourier_finish_fwd_kernel<<<frb_div_up(hw * n_views, 256), 256, 0, st>>>(n_views, width, height, accum, mx_key,
                                                                             bg, image);
    frb_note_launches(2);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_fourier_finish_bwd(int n_views, int width, int height, const float* accum,
                                      const uint32_t* mx_key, const float* background_host, const float* g_image,
                                      float* red, float* gpix, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!accum || !mx_key || !background_host || !g_image || !red || !gpix) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    const long long hw = (long long)width * height;
    FRB_CUDA_OK(cudaMemsetAsync(red, 0, sizeof(float) * 2 * n_views, st));
    dim3 grid(min(frb_div_up(hw, 256), 592), n_views);
    fourier_bwd_reduce_kernel<<<grid, 256, 0, st>>>(width, height, accum, mx_key, bg, g_image, red);
    fourier_finish_bwd_kernel<<<frb_div_up(hw * n_views, 256), 256, 0, st>>>(n_views, width, height, accum, mx_key,
                                                                             bg, g_image, red, gpix);
    frb_note_launches(2);
    FRB_LAUNCH_CHECK();
    return 0;
}
