// On-disk Gaussian formats -> the renderer's SoA device tensors (and back), SURVEY.md section 8 f4.
//   .bin  14 floats per Gaussian: position 3, scale 3, rotation wxyz 4, colour 3, opacity 1
//         (load/save_gaussians_*_binary DR:1461-1497 == GaussianCloud::save/load_binary
//          src/core/renderer/renderer.cpp:557-647)
//   .ply  3DGS-style binary_little_endian vertex rows of the same 14 floats with the standard
//         parameterisation: log scale, f_dc = (colour - 0.5) / C0, logit opacity
//         (GaussianCloud::save_ply / load_ply renderer.cpp:649-793)
// The file body is copied to the device as it is (one H2D of 56 n bytes) and split there: HBM-bound,
// 56 B in / 56 B out per Gaussian, one thread per float of the row so that both sides are coalesced.
#include "frb_common.cuh"

namespace {

constexpr int ROW = 14;
constexpr float SH_C0 = 0.28209479177387814f;          // renderer.cpp:707, 744

// column of the 14-float row -> (tensor id, component): 0 pos, 1 scale, 2 rot, 3 colour, 4 opacity
__device__ __forceinline__ void column_target(int c, int& tensor, int& comp, int& width) {
    if (c < 3) { tensor = 0; comp = c; width = 3; }
    else if (c < 6) { tensor = 1; comp = c - 3; width = 3; }
    else if (c < 10) { tensor = 2; comp = c - 6; width = 4; }
    else if (c < 13) { tensor = 3; comp = c - 10; width = 3; }
    else { tensor = 4; comp = 0; width = 1; }
}

__global__ void __launch_bounds__(256)
unpack_rows_kernel(long long total, int ply, const float* __restrict__ rows, float* __restrict__ positions,
                   float* __restrict__ scales, float* __restrict__ rotations, float* __restrict__ colors,
                   float* __restrict__ opacities) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / ROW;
        const int c = (int)(t - i * ROW);
        float v = rows[t];
        int tensor, comp, width;
        column_target(c, tensor, comp, width);
        if (ply) {
            if (tensor == 1) v = expf(v);                                         // renderer.cpp:766-768
            else if (tensor == 3) v = fminf(fmaxf(v * SH_C0 + 0.5f, 0.0f), 1.0f); // renderer.cpp:777-779
            else if (tensor == 4) v = 1.0f / (1.0f + expf(-v));                   // renderer.cpp:782
        }
        float* dst = tensor == 0 ? positions : tensor == 1 ? scales : tensor == 2 ? rotations
                   : tensor == 3 ? colors : opacities;
        dst[i * width + comp] = v;
    }
}

__global__ void __launch_bounds__(256)
pack_rows_kernel(long long total, int ply, const float* __restrict__ positions, const float* __restrict__ scales,
                 const float* __restrict__ rotations, const float* __restrict__ colors,
                 const float* __restrict__ opacities, float* __restrict__ rows) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / ROW;
        const int c = (int)(t - i * ROW);
        int tensor, comp, width;
        column_target(c, tensor, comp, width);
        const float* src = tensor == 0 ? positions : tensor == 1 ? scales : tensor == 2 ? rotations
                         : tensor == 3 ? colors : opacities;
        float v = src[i * width + comp];
        if (ply) {
            if (tensor == 1) v = logf(fmaxf(v, 1e-7f));                           // renderer.cpp:687-689
            else if (tensor == 3) v = (v - 0.5f) / SH_C0;                         // renderer.cpp:708-710
            else if (tensor == 4) v = logf(v / fmaxf(1.0f - v, 1e-7f));           // renderer.cpp:716
        }
        rows[t] = v;
    }
}

}  // namespace

extern "C" int frb_unpack_gaussians(int n, int ply, const float* rows, float* positions, float* scales,
                                    float* rotations, float* colors, float* opacities, void* stream) {
    if (n < 0 || (ply != 0 && ply != 1)) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!rows || !positions || !scales || !rotations || !colors || !opacities) return FRB_E_INVALID;
    const long long total = (long long)n * ROW;
    unpack_rows_kernel<<<min(frb_div_up(total, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        total, ply, rows, positions, scales, rotations, colors, opacities);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_pack_gaussians(int n, int ply, const float* positions, const float* scales,
                                  const float* rotations, const float* colors, const float* opacities,
                                  float* rows, void* stream) {
    if (n < 0 || (ply != 0 && ply != 1)) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!rows || !positions || !scales || !rotations || !colors || !opacities) return FRB_E_INVALID;
    const long long total = (long long)n * ROW;
    pack_rows_kernel<<<min(frb_div_up(total, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        total, ply, positions, scales, rotations, colors, opacities, rows);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
