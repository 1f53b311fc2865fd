// Shared device helpers: error plumbing, mbarrier + 1D TMA bulk copy (sm_100a PTX), warp utilities.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fresnel_b200.h"
#include "frb_math.h"

#define FRB_MAX_VIEWS 32

struct FrbViewSet {
    int n_views;
    int n_per_view;
    FrbCamera cam[FRB_MAX_VIEWS];
};

// Fills `vs` from n_views * FRB_CAMERA_FLOATS host floats.  Returns 0 or FRB_E_INVALID.
int frb_fill_views(int n, int n_views, const float* camera_host, FrbViewSet* vs);

#define FRB_CUDA_OK(expr)                          \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

#define FRB_LAUNCH_CHECK()                         \
    do {                                           \
        cudaError_t _e = cudaGetLastError();       \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

// 2D tensor map over an array of n 48-byte records (12 floats per row) for frb_tma_gather4: box {16, 1}, no swizzle.
// Encoded through the driver entry point cuTensorMapEncodeTiled (no link-time dependency on libcuda).
int frb_record_tensor_map(const float* records, int n, CUtensorMap* out);

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
void frb_note_launches(int k);

static inline int frb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool frb_misaligned16(const void* p) { return ((uintptr_t)p & 15u) != 0; }   // NULL counts as aligned

// Programmatic dependent launch (sm_90+): kernels of the render chain are launched with the programmatic stream
// serialization attribute, so that the NEXT kernel's CTAs are scheduled (and run their prologue up to
// frb_pdl_prologue()) while the previous kernel drains, instead of after it has drained.  Every chain kernel calls
// frb_pdl_prologue() before its first global-memory access; without the attribute it is a no-op.
// FRB_PDL=0 in the environment launches the plain way.
bool frb_pdl_enabled();

#if defined(__CUDACC__)

// Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute is per DEVICE (one process may render
// on several GPUs), so the "already done" state is a bitmask indexed by the current device, not a process-wide flag;
// the mask is atomic because callers may launch from several host threads.
template <typename K>
static inline cudaError_t frb_opt_in_smem(K kernel, int bytes, unsigned long long* done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (dev < 64 && (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
    return e;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t frb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = frb_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Let the dependent grid be scheduled now; then wait until the grid this one depends on has completed and its
// writes are visible.  (A dependent grid waits at the same point for THIS grid, so starting it early is safe.)
__device__ __forceinline__ void frb_pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Index of the highest set bit (x != 0): bfind -> one FLO.U32; 31 - __clz(x) costs two more instructions.
__device__ __forceinline__ int frb_bfind(uint32_t x) {
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

__device__ __forceinline__ uint32_t frb_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier (shared::cta) ------------------------------------------------
__device__ __forceinline__ void frb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(frb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void frb_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void frb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(frb_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void frb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(frb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- TMA: 1D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void frb_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            frb_smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(frb_smem_u32(bar))
        : "memory");
}

// ---- TMA tile::gather4: four rows of a 2D tensor map (row = one record) by row index (SASS: UTMALDG.2D.GATHER4) ----
// dst 128-byte aligned; the tensor map's box is {row floats, 1}; completes 4 x row bytes on the mbarrier.
__device__ __forceinline__ void frb_tma_gather4(void* smem_dst, const CUtensorMap* map, uint4 rows, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(frb_smem_u32(smem_dst)),
        "l"(map), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(frb_smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float frb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float frb_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32 pairs (sm_100a FFMA2 / FMUL2 / FADD2: one issue slot for two IEEE fp32 operations) --------
// The compositor and splat kernels are issue-bound with the FMA pipe at ~40 %, so pairing independent FMAs that
// share a multiplicand (SASS takes the scalar as a broadcast operand, "R4.F32") removes issue slots without
// changing a single result bit: each half is a correctly rounded fp32 fma, exactly like fmaf.
__device__ __forceinline__ unsigned long long frb_pack2(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 frb_unpack2(unsigned long long r) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
    return d;
}
__device__ __forceinline__ float2 frb_fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(frb_pack2(a.x, a.y)), "l"(frb_pack2(b.x, b.y)),
        "l"(frb_pack2(c.x, c.y)));
    return frb_unpack2(d);
}
// (s * b.x + c.x, s * b.y + c.y)
__device__ __forceinline__ float2 frb_fma2s(float s, float2 b, float2 c) { return frb_fma2(make_float2(s, s), b, c); }
__device__ __forceinline__ float2 frb_mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(frb_pack2(a.x, a.y)), "l"(frb_pack2(b.x, b.y)));
    return frb_unpack2(d);
}
__device__ __forceinline__ float2 frb_mul2s(float s, float2 b) { return frb_mul2(make_float2(s, s), b); }
__device__ __forceinline__ float2 frb_add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(frb_pack2(a.x, a.y)), "l"(frb_pack2(b.x, b.y)));
    return frb_unpack2(d);
}

// 16-byte vector reduction into global memory (sm_90+): four fp32 adds in one L2 atomic transaction, no return
// value.  addr must be 16-byte aligned.
__device__ __forceinline__ void frb_red_add_f4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Orders generic-proxy shared-memory writes before later async-proxy (TMA) accesses.
__device__ __forceinline__ void frb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float frb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
