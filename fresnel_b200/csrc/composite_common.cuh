// Definitions shared by the plain and the phase-blending compositor kernels.
#pragma once
#include "frb_common.cuh"

namespace {

constexpr int TILE = FRB_TILE;
constexpr int CTA_THREADS = TILE * TILE;
constexpr int BATCH = 64;                       // records per stage
constexpr int STAGES = 4;
constexpr int RECORD_BYTES = FRB_RECORD_FLOATS * 4;
constexpr float T_FLOOR = 1e-20f;
constexpr int CHUNK = 4;                        // records between early-termination tests
constexpr uint32_t NULL_RECT_LO = 0x7fff7fffu;  // x0 = y0 = 32767: contains no pixel
constexpr uint32_t NULL_RECT_HI = 0x80008000u;  // x1 = y1 = 0
constexpr int STATE_GATE_SHIFT = 28;            // state_n = entries consumed | clamp gates << 28
constexpr int STATE_N_MASK = (1 << STATE_GATE_SHIFT) - 1;

// Pixel owned by a thread: a warp covers an 8x4 block of the 16x16 tile, not a 16x2 strip - fewer list entries have
// a rectangle that touches the squarer block, so fewer warps execute the per-pixel body for an entry.
constexpr int FOOT_W = 8, FOOT_H = 4;
__device__ __forceinline__ int foot_x(int tid) { return ((tid >> 5) % (TILE / FOOT_W)) * FOOT_W + ((tid & 31) % FOOT_W); }
__device__ __forceinline__ int foot_y(int tid) { return ((tid >> 5) / (TILE / FOOT_W)) * FOOT_H + ((tid & 31) / FOOT_W); }

struct __align__(16) StageBuf {
    float4 rec[BATCH * 3];
};

__device__ __forceinline__ bool rect_contains(uint32_t pxy_guard, uint32_t pxy_plus1, uint32_t lo, uint32_t hi) {
    // SWAR test of x0 <= px < x1 and y0 <= py < y1 on 15-bit halves with guard bits:
    // (px | G) - x0 keeps G iff px >= x0 ; (x1 | G) - (px + 1) keeps G iff px < x1.
    uint32_t a = pxy_guard - lo;
    uint32_t b = hi - pxy_plus1;
    return ((a & b) & 0x80008000u) == 0x80008000u;
}


// Sum of N per-lane values over the warp with a halving butterfly: 16 shuffles for N <= 16
// instead of 5 N.  On return lane l (l even) holds the warp total of value
// idx = bit4(l) << 3 | bit3(l) << 2 | bit2(l) << 1 | bit1(l); odd lanes hold the same as l - 1.
template <int N>
__device__ __forceinline__ float warp_reduce_multi(const float (&in)[N], int lane) {
    static_assert(N <= 16, "at most 16 values");
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (k < N) ? in[k] : 0.0f;
#pragma unroll
    for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            float keep = upper ? v[k + half] : v[k];
            float send = upper ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// The same over the 16 lanes of a half-warp (both halves at once): 15 shuffles; lane l of the half (hl = lane & 15)
// ends with the half's total of value hl.
template <int N>
__device__ __forceinline__ float halfwarp_reduce_multi(const float (&in)[N], int hl) {
    static_assert(N <= 16, "at most 16 values");
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (k < N) ? in[k] : 0.0f;
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool upper = (hl & half) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            float keep = upper ? v[k + half] : v[k];
            float send = upper ? v[k] : v[k + half];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

__device__ __forceinline__ int warp_reduce_multi_index(int lane) {
    return ((lane >> 4) & 1) << 3 | ((lane >> 3) & 1) << 2 | ((lane >> 2) & 1) << 1 | ((lane >> 1) & 1);
}

}  // namespace
