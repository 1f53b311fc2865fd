// Reconstruction loss front-end: what the training step computes on the renderer's
// outputs, as four kernels instead of about forty elementwise / reduction launches.
// Reference: compute_losses, scripts/training/train_gaussian_decoder.py:838-930, the terms that exist in this
// image (SSIM / LPIPS packages are absent and the reference then drops them):
//     rgb   = mean |rendered - target|                                              (:889, :892)
//     depth = mean |(rd - mean rd) / max(std rd, 1e-4) - (td - mean td) / max(std td, 1e-4)|   (:921-927)
//     bnd   = mean_{b,y,x} [ mean_c |rendered - target| * mask(target_depth) ]            (:941-953)
//             mask = sigmoid(10 / thr * (thr - min_k |d - boundary_k|))    (FresnelZones.compute_boundary_mask,
//             scripts/utils/fresnel_zones.py:141-180; hard variant: 1[min_k |d - boundary_k| < thr])
//     loss  = rgb_weight * rgb + depth_weight * depth + boundary_weight * bnd
//             (std is torch's default: unbiased, n - 1; the mask depends on the TARGET depth only: no gradient)
// Backward (hand-derived): with a_i = (rd_i - mu) / sigma, s_i = sign(a_i - b_i), n pixels,
//     dL/drd_i = depth_weight / n * [ s_i / sigma - S1 / (n sigma) - 1[std >= 1e-4] (rd_i - mu) S2 / ((n-1) sigma^3) ]
//     S1 = sum s_j,  S2 = sum s_j (rd_j - mu);
//     dL/drendered = sign(rendered - target) * (rgb_weight + boundary_weight * mask) / (3 n)
// HBM-bound: every tensor is read twice forward (statistics, then the normalised difference) and once backward.
#include "frb_common.cuh"

namespace {

// stats (double): 0 sum|r-t| 1 sum rd 2 sum rd^2 3 sum td 4 sum td^2 5 sum|a-b| 6 S1 7 S2 8 sum|r-t| mask
constexpr int LOSS_STATS = 12;
constexpr int MAX_ZONE_BOUNDARIES = 65;

// Zone boundaries (torch.linspace values, handed over by the host so that they are the reference's floats).
struct ZoneMask {
    int n;                // boundaries (0 = no boundary term)
    int soft;
    float threshold, sharpness;
    float b[MAX_ZONE_BOUNDARIES];
};

__device__ __forceinline__ float zone_mask(const ZoneMask& z, float d) {
    float md = fabsf(d - z.b[0]);
    for (int k = 1; k < z.n; ++k) md = fminf(md, fabsf(d - z.b[k]));
    if (!z.soft) return md < z.threshold ? 1.0f : 0.0f;
    return 1.0f / (1.0f + expf(-(z.sharpness * (z.threshold - md))));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int K>
__device__ __forceinline__ void block_accumulate(const double (&v)[K], double* __restrict__ out) {
    __shared__ double sh[K][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double w = warp_sum_d(v[k]);
        if (lane == 0) sh[k][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, t);
    }
}

__global__ void __launch_bounds__(256)
loss_stats1_kernel(long long n_rgb, long long n_pix, long long hw, const float* __restrict__ rendered,
                   const float* __restrict__ target, const float* __restrict__ rd, const float* __restrict__ td,
                   const __grid_constant__ ZoneMask zm, double* __restrict__ stats) {
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double vb[1] = {0.0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rgb; i += stride) {
        const float ad = fabsf(rendered[i] - target[i]);
        v[0] += (double)ad;
        if (zm.n) {       // element i = (view, channel, pixel): its mask is the target depth's at (view, pixel)
            const long long view = i / (3 * hw), pix = i % hw;
            vb[0] += (double)(ad * zone_mask(zm, td[view * hw + pix]));
        }
    }
    if (rd)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += stride) {
            const double a = rd[i], b = td[i];
            v[1] += a; v[2] += a * a; v[3] += b; v[4] += b * b;
        }
    block_accumulate<5>(v, stats);
    if (zm.n) {
        __syncthreads();
        block_accumulate<1>(vb, stats + 8);
    }
}

struct DepthNorm {
    float mu_r, inv_r, mu_t, inv_t;
    int gate_r;          // std(rd) >= 1e-4: the clamp passes its gradient
};

__device__ __forceinline__ DepthNorm depth_norm(const double* __restrict__ stats, long long n) {
    DepthNorm d;
    const double nn = (double)n, mr = stats[1] / nn, mt = stats[3] / nn;
    const double den = nn > 1.0 ? nn - 1.0 : 1.0;
    const double vr = fmax((stats[2] - nn * mr * mr) / den, 0.0), vt = fmax((stats[4] - nn * mt * mt) / den, 0.0);
    const float sr = (float)sqrt(vr), st = (float)sqrt(vt);
    d.mu_r = (float)mr; d.mu_t = (float)mt;
    d.gate_r = sr >= 1e-4f;
    d.inv_r = 1.0f / fmaxf(sr, 1e-4f);
    d.inv_t = 1.0f / fmaxf(st, 1e-4f);
    return d;
}

__global__ void __launch_bounds__(256)
loss_stats2_kernel(long long n_pix, const float* __restrict__ rd, const float* __restrict__ td,
                   double* __restrict__ stats) {
    const DepthNorm d = depth_norm(stats, n_pix);
    double v[3] = {0.0, 0.0, 0.0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += stride) {
        const float c = rd[i] - d.mu_r;
        const float diff = c * d.inv_r - (td[i] - d.mu_t) * d.inv_t;
        const float s = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
        v[0] += (double)fabsf(diff); v[1] += (double)s; v[2] += (double)(s * c);
    }
    block_accumulate<3>(v, stats + 5);
}

__global__ void loss_finish_kernel(long long n_rgb, long long n_pix, int has_depth, float rgb_weight,
                                   float depth_weight, float boundary_weight, const double* __restrict__ stats,
                                   float* __restrict__ loss) {
    double l = (double)rgb_weight * stats[0] / (double)n_rgb;
    if (has_depth) l += (double)depth_weight * stats[5] / (double)n_pix;
    if (boundary_weight != 0.0f) l += (double)boundary_weight * stats[8] / (double)n_rgb;   // mean_c then mean_pix
    loss[0] = (float)l;
}

__global__ void __launch_bounds__(256)
loss_bwd_kernel(long long n_rgb, long long n_pix, long long hw, const float* __restrict__ rendered,
                const float* __restrict__ target, const float* __restrict__ rd, const float* __restrict__ td,
                float rgb_weight, float depth_weight, float boundary_weight, const __grid_constant__ ZoneMask zm,
                const double* __restrict__ stats, const float* __restrict__ g_loss, float* __restrict__ g_rendered,
                float* __restrict__ g_rd) {
    const float g = g_loss[0];
    const float k_rgb = g * rgb_weight / (float)n_rgb;
    const float k_bnd = g * boundary_weight / (float)n_rgb;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rgb; i += stride) {
        const float diff = rendered[i] - target[i];
        float k = k_rgb;
        if (zm.n) {
            const long long view = i / (3 * hw), pix = i % hw;
            k += k_bnd * zone_mask(zm, td[view * hw + pix]);
        }
        g_rendered[i] = (diff > 0.f) ? k : ((diff < 0.f) ? -k : 0.f);
    }
    if (rd && g_rd) {
        const DepthNorm d = depth_norm(stats, n_pix);
        const float nn = (float)n_pix;
        const float k = g * depth_weight / nn;
        const float c1 = (float)stats[6] / nn * d.inv_r;                                   // S1 / (n sigma)
        const float c2 = d.gate_r ? (float)stats[7] * d.inv_r * d.inv_r * d.inv_r / fmaxf(nn - 1.0f, 1.0f) : 0.f;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += stride) {
            const float c = rd[i] - d.mu_r;
            const float diff = c * d.inv_r - (td[i] - d.mu_t) * d.inv_t;
            const float s = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
            g_rd[i] = k * (s * d.inv_r - c1 - c * c2);
        }
    }
}

}  // namespace

extern "C" size_t frb_recon_loss_workspace_bytes(void) { return sizeof(double) * LOSS_STATS; }

static int fill_zone_mask(float boundary_weight, int n_boundaries, const float* boundaries_host, float threshold,
                          int soft, long long n_pix, long long hw, const float* target_depth, ZoneMask* zm) {
    zm->n = 0;
    zm->soft = soft;
    zm->threshold = threshold;
    zm->sharpness = threshold != 0.0f ? 10.0f / threshold : 0.0f;         // fresnel_zones.py:172
    if (boundary_weight == 0.0f) return 0;
    if (n_boundaries < 1 || n_boundaries > MAX_ZONE_BOUNDARIES || !boundaries_host || !target_depth) return FRB_E_INVALID;
    if (hw < 1 || n_pix % hw != 0) return FRB_E_INVALID;
    zm->n = n_boundaries;
    for (int k = 0; k < n_boundaries; ++k) zm->b[k] = boundaries_host[k];
    return 0;
}

extern "C" int frb_recon_loss_fwd_ex(long long n_rgb, long long n_pix, long long hw, const float* rendered,
                                     const float* target, const float* rendered_depth, const float* target_depth,
                                     float rgb_weight, float depth_weight, float boundary_weight, int n_boundaries,
                                     const float* boundaries_host, float boundary_threshold, int soft_boundaries,
                                     void* stats, float* loss, void* stream) {
    if (n_rgb < 1 || n_pix < 0 || !rendered || !target || !stats || !loss) return FRB_E_INVALID;
    if (rendered_depth && !target_depth) return FRB_E_INVALID;
    ZoneMask zm;
    int rc = fill_zone_mask(boundary_weight, n_boundaries, boundaries_host, boundary_threshold, soft_boundaries,
                            n_pix, hw, target_depth, &zm);
    if (rc) return rc;
    if (zm.n && n_rgb != 3 * n_pix) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    double* s = (double*)stats;
    FRB_CUDA_OK(cudaMemsetAsync(s, 0, sizeof(double) * LOSS_STATS, st));
    const int grid = min(frb_div_up(n_rgb, 256 * 4), 148 * 8);
    loss_stats1_kernel<<<grid, 256, 0, st>>>(n_rgb, n_pix, hw, rendered, target, rendered_depth,
                                             target_depth, zm, s);
    int launches = 2;
    if (rendered_depth) {
        loss_stats2_kernel<<<min(frb_div_up(n_pix, 256 * 4), 148 * 8), 256, 0, st>>>(n_pix, rendered_depth,
                                                                                      target_depth, s);
        ++launches;
    }
    loss_finish_kernel<<<1, 1, 0, st>>>(n_rgb, n_pix, rendered_depth != nullptr, rgb_weight, depth_weight,
                                        zm.n ? boundary_weight : 0.0f, s, loss);
    frb_note_launches(launches);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_recon_loss_bwd_ex(long long n_rgb, long long n_pix, long long hw, const float* rendered,
                                     const float* target, const float* rendered_depth, const float* target_depth,
                                     float rgb_weight, float depth_weight, float boundary_weight, int n_boundaries,
                                     const float* boundaries_host, float boundary_threshold, int soft_boundaries,
                                     const void* stats, const float* g_loss, float* g_rendered,
                                     float* g_rendered_depth, void* stream) {
    if (n_rgb < 1 || n_pix < 0 || !rendered || !target || !stats || !g_loss || !g_rendered) return FRB_E_INVALID;
    if (rendered_depth && (!target_depth || !g_rendered_depth)) return FRB_E_INVALID;
    ZoneMask zm;
    int rc = fill_zone_mask(boundary_weight, n_boundaries, boundaries_host, boundary_threshold, soft_boundaries,
                            n_pix, hw, target_depth, &zm);
    if (rc) return rc;
    loss_bwd_kernel<<<min(frb_div_up(n_rgb, 256 * 4), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        n_rgb, n_pix, hw, rendered, target, rendered_depth, target_depth, rgb_weight, depth_weight,
        zm.n ? boundary_weight : 0.0f, zm, (const double*)stats, g_loss, g_rendered, g_rendered_depth);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_recon_loss_fwd(long long n_rgb, long long n_pix, const float* rendered, const float* target,
                                  const float* rendered_depth, const float* target_depth, float rgb_weight,
                                  float depth_weight, void* stats, float* loss, void* stream) {
    if ((rendered_depth == nullptr) != (target_depth == nullptr)) return FRB_E_INVALID;
    return frb_recon_loss_fwd_ex(n_rgb, n_pix, n_pix > 0 ? n_pix : 1, rendered, target, rendered_depth, target_depth,
                                 rgb_weight, depth_weight, 0.0f, 0, nullptr, 0.0f, 1, stats, loss, stream);
}

extern "C" int frb_recon_loss_bwd(long long n_rgb, long long n_pix, const float* rendered, const float* target,
                                  const float* rendered_depth, const float* target_depth, float rgb_weight,
                                  float depth_weight, const void* stats, const float* g_loss, float* g_rendered,
                                  float* g_rendered_depth, void* stream) {
    if ((rendered_depth == nullptr) != (target_depth == nullptr)) return FRB_E_INVALID;
    return frb_recon_loss_bwd_ex(n_rgb, n_pix, n_pix > 0 ? n_pix : 1, rendered, target, rendered_depth, target_depth,
                                 rgb_weight, depth_weight, 0.0f, 0, nullptr, 0.0f, 1, stats, g_loss, g_rendered,
                                 g_rendered_depth, stream);
}
