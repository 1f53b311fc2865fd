// Gradient exchange fused with the optimiser step over NVLink peer memory.
//
// The multi-view trainer keeps ONE cloud replicated on every rank; each rank renders its own view and the
// per-Gaussian gradients (15 floats x 1M = 60 MB) have to be summed before every rank applies the same Adam
// update (the loop being sharded is train_gaussian_decoder.py:1209-1266: render, loss.backward(),
// optimizer.step()).  The library form of that is all-reduce + replicated Adam: 2 x 7/8 x 60 MB over the
// links, then 7 x 60 MB of HBM traffic on every rank for an update all ranks compute identically.
//
// Here it is ONE kernel per rank over peer-mapped buffers (reduce-scatter + sharded Adam + all-gather):
//   entry barrier   rank r tells every peer "my gradients are complete" and waits for the same from all;
//   main loop       rank r owns the float4s [lo_r, hi_r): it LOADS that slice of the gradient from every
//                   peer (fixed rank order: every rank ends up with bit-identical parameters), updates its
//                   shard of the Adam moments (stored only here: 1/world of the optimiser state), and
//                   STORES the new parameters into every peer's parameter buffer;
//   exit barrier    the last CTA to finish tells every peer "my stores are visible" and waits for the same,
//                   so the kernel that follows on the stream reads a complete parameter buffer and may
//                   overwrite the gradient buffer (every peer has finished reading it).
// The step counter lives on the device (state[1]), so the launch arguments never change and the kernel
// replays from a CUDA graph.  world = 1 degenerates to a plain fused Adam over local memory.
#include "frb_common.cuh"

namespace {

struct PeerAdamArgs {
    int world, rank;
    long long n_floats;
    const unsigned long long* grad_ptrs;     // [world] device array: every rank's gradient buffer
    const unsigned long long* param_ptrs;    // [world] every rank's parameter buffer
    const unsigned long long* signal_ptrs;   // [world] every rank's signal pad: 2 * world uint32, zero at start
    float* exp_avg;                          // this rank's shard of the first / second moments
    float* exp_avg_sq;
    uint32_t* state;                         // local: [0] CTA ticket counter, [1] steps taken so far
    double lr, beta1, beta2;                 // double: the bias corrections are formed as torch forms them
    float eps, grad_scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Peer data is read exactly once: bypass L1 (a stale line from the previous step must not be hit).
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float adam_one(float g, float p, float& m, float& v, float b1, float b2, float step_size,
                                          float sqrt_bc2, float eps) {
    // torch.optim.Adam (no weight decay, no amsgrad): lerp form of the first moment, addcmul of the second
    m = m + (g - m) * (1.0f - b1);
    v = v * b2 + (1.0f - b2) * g * g;
    const float denom = sqrtf(v) / sqrt_bc2 + eps;
    return p - step_size * (m / denom);
}

template <int MAXW>
__global__ void __launch_bounds__(256) peer_adam_kernel(PeerAdamArgs a) {
    __shared__ uint32_t s_last;
    __shared__ float s_step_size, s_sqrt_bc2;
    const int W = a.world, R = a.rank;
    const uint32_t epoch = a.state[1] + 1u;          // every CTA reads it before the last one bumps it
    uint32_t* my_pad = (uint32_t*)a.signal_ptrs[R];

    // ---- entry barrier: gradients of every rank are complete ------------------------------------
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys((uint32_t*)a.signal_ptrs[threadIdx.x] + R, epoch);
    }
    if (threadIdx.x < W) {
        while ((int32_t)(ld_acquire_sys(my_pad + threadIdx.x) - epoch) < 0) __nanosleep(64);
    }
    if (threadIdx.x == 32) {                         // bias corrections in fp64, once per CTA
        const double bc1 = 1.0 - pow(a.beta1, (double)epoch), bc2 = 1.0 - pow(a.beta2, (double)epoch);
        s_step_size = (float)(a.lr / bc1);
        s_sqrt_bc2 = (float)sqrt(bc2);
    }
    __syncthreads();

    // ---- shard owned by this rank, in float4 units (the < 4 float tail goes to the last rank) ----
    const long long n4 = a.n_floats >> 2;
    const long long per = (n4 + W - 1) / W;
    const long long lo = min(n4, per * R), hi = min(n4, lo + per);
    const float step_size = s_step_size, sqrt_bc2 = s_sqrt_bc2;
    const float b1 = (float)a.beta1, b2 = (float)a.beta2;

    const float4* gp[MAXW];
    float4* pp[MAXW];
#pragma unroll
    for (int k = 0; k < MAXW; ++k) {
        gp[k] = (const float4*)a.grad_ptrs[k < W ? k : 0];
        pp[k] = (float4*)a.param_ptrs[k < W ? k : 0];
    }
    const float4* own = (const float4*)a.param_ptrs[R];
    float4* m4 = (float4*)a.exp_avg;
    float4* v4 = (float4*)a.exp_avg_sq;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        float4 g[MAXW];
#pragma unroll
        for (int k = 0; k < MAXW; ++k)
            if (k < W) g[k] = ld_peer_f4(gp[k] + i);          // all loads in flight before the first use
        float4 p = own[i];                                    // own copy (only this rank writes this slice)
        float4 m = m4[i - lo], v = v4[i - lo];
        float4 s = g[0];
#pragma unroll
        for (int k = 1; k < MAXW; ++k)
            if (k < W) { s.x += g[k].x; s.y += g[k].y; s.z += g[k].z; s.w += g[k].w; }
        s.x *= a.grad_scale; s.y *= a.grad_scale; s.z *= a.grad_scale; s.w *= a.grad_scale;
        p.x = adam_one(s.x, p.x, m.x, v.x, b1, b2, step_size, sqrt_bc2, a.eps);
        p.y = adam_one(s.y, p.y, m.y, v.y, b1, b2, step_size, sqrt_bc2, a.eps);
        p.z = adam_one(s.z, p.z, m.z, v.z, b1, b2, step_size, sqrt_bc2, a.eps);
        p.w = adam_one(s.w, p.w, m.w, v.w, b1, b2, step_size, sqrt_bc2, a.eps);
        m4[i - lo] = m;
        v4[i - lo] = v;
#pragma unroll
        for (int k = 0; k < MAXW; ++k)
            if (k < W) pp[k][i] = p;                          // all-gather: the new slice lands on every rank
    }
    // scalar tail (n_floats not a multiple of 4): last rank, first CTA
    if (R == W - 1 && blockIdx.x == 0) {
        const long long tail0 = n4 << 2;
        const long long i = tail0 + threadIdx.x;
        if (i < a.n_floats) {
            float s = 0.0f;
            for (int k = 0; k < W; ++k) s += ld_peer_f1((const float*)a.grad_ptrs[k] + i);
            s *= a.grad_scale;
            const long long j = ((hi - lo) << 2) + threadIdx.x;    // moments of the tail follow the shard
            float m = a.exp_avg[j], v = a.exp_avg_sq[j];
            float p = ((const float*)a.param_ptrs[R])[i];
            p = adam_one(s, p, m, v, b1, b2, step_size, sqrt_bc2, a.eps);
            a.exp_avg[j] = m;
            a.exp_avg_sq[j] = v;
            for (int k = 0; k < W; ++k) ((float*)a.param_ptrs[k])[i] = p;
        }
    }

    // ---- exit barrier: every rank's stores are visible everywhere, every rank has finished reading ----
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.state, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < W) {
        st_release_sys((uint32_t*)a.signal_ptrs[threadIdx.x] + W + R, epoch);
        while ((int32_t)(ld_acquire_sys(my_pad + W + threadIdx.x) - epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.state[0] = 0u;
        a.state[1] = epoch;
    }
}

}  // namespace

extern "C" long long frb_peer_shard_floats(int world, int rank, long long n_floats) {
    if (world < 1 || rank < 0 || rank >= world || n_floats < 0) return -1;
    const long long n4 = n_floats >> 2, per = (n4 + world - 1) / world;
    const long long lo = per * rank < n4 ? per * rank : n4;
    const long long hi = lo + per < n4 ? lo + per : n4;
    return ((hi - lo) << 2) + (rank == world - 1 ? (n_floats & 3) : 0);
}

extern "C" int frb_peer_adam_step(int world, int rank, long long n_floats, const unsigned long long* grad_ptrs,
                                  const unsigned long long* param_ptrs, const unsigned long long* signal_ptrs,
                                  float* exp_avg, float* exp_avg_sq, uint32_t* state, double lr, double beta1,
                                  double beta2, float eps, float grad_scale, void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || n_floats < 0) return FRB_E_INVALID;
    if (!grad_ptrs || !param_ptrs || !signal_ptrs || !exp_avg || !exp_avg_sq || !state) return FRB_E_INVALID;
    PeerAdamArgs a{world, rank, n_floats, grad_ptrs, param_ptrs, signal_ptrs, exp_avg, exp_avg_sq, state,
                   lr,    beta1, beta2,   eps,       grad_scale};
    int dev = 0, sms = 148;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    FRB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one resident wave of CTAs walking the shard with a grid stride (no CTA waits for another local CTA:
    // the entry barrier is fed by block 0 of every rank, the exit barrier is run by whichever CTA ends last)
    void (*kern)(PeerAdamArgs) = world <= 2 ? peer_adam_kernel<2> : world <= 4 ? peer_adam_kernel<4>
                                 : world <= 8 ? peer_adam_kernel<8> : peer_adam_kernel<16>;
    int per_sm = 1;
    FRB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) per_sm = 1;
    const long long shard4 = (frb_peer_shard_floats(world, rank, n_floats) + 3) >> 2;
    int grid = (int)((shard4 + 255) / 256);
    if (grid > per_sm * sms) grid = per_sm * sms;
    if (grid < 1) grid = 1;
    kern<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
