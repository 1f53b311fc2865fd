// Probe (not product code): does TMA tile::gather4 (UTMALDG.2D.GATHER4, sm_100a) fetch 48-byte records by row index
// into shared memory the way a compositor would need it?  Variant A: box {12, 1} (48-byte rows packed, group
// destination only 16-byte aligned when the group index is odd).  Variant B: box {16, 1} (64-byte rows, 256-byte
// groups: 128-byte aligned destinations; columns 12..15 are out of bounds and zero-filled).
// Every wait is bounded, so a wrong transaction count reports instead of hanging.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/_gather4_probe tools/probes/tma_gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROW_FLOATS>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ gids, int m,
                             float* __restrict__ out, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* buf = reinterpret_cast<float*>(smem);
    __shared__ __align__(8) uint64_t bar;
    const int groups = m / 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)),
                     "r"(groups * 4 * ROW_FLOATS * 4) : "memory");
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
        const int q = threadIdx.x;
        const int r0 = gids[4 * q], r1 = gids[4 * q + 1], r2 = gids[4 * q + 2], r3 = gids[4 * q + 3];
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(s32(buf + q * 4 * ROW_FLOATS)),
            "l"(&tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(&bar))
            : "memory");
    }
    // bounded wait
    bool ok = false;
    for (int it = 0; it < (1 << 22) && !ok; ++it) {
        uint32_t p;
        asm volatile(
            "{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0, 1, 0, q;\n}\n"
            : "=r"(p) : "r"(s32(&bar)) : "memory");
        ok = p != 0;
    }
    if (threadIdx.x == 0) status[blockIdx.x] = ok ? 1 : -1;
    __syncthreads();
    if (ok)
        for (int i = threadIdx.x; i < m * ROW_FLOATS; i += blockDim.x) out[(size_t)blockIdx.x * m * ROW_FLOATS + i] = buf[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROW_FLOATS>
int run(EncodeFn encode, const float* d_rec, int n, const int* d_gid, int m, const std::vector<float>& rec,
        const std::vector<int>& gid) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {12, (cuuint64_t)n};
    cuuint64_t gstride[1] = {48};
    cuuint32_t box[2] = {(cuuint32_t)ROW_FLOATS, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d_rec, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %d: encode rc=%d\n", ROW_FLOATS, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int blocks = 4;
    float* d_out;
    int* d_status;
    cudaMalloc(&d_out, sizeof(float) * blocks * m * ROW_FLOATS);
    cudaMalloc(&d_status, sizeof(int) * blocks);
    cudaMemset(d_status, 0, sizeof(int) * blocks);
    probe_kernel<ROW_FLOATS><<<blocks, 128, m * ROW_FLOATS * 4>>>(tm, d_gid, m, d_out, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box %d: kernel %s\n", ROW_FLOATS, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> out(blocks * m * ROW_FLOATS);
    int status[4];
    cudaMemcpy(out.data(), d_out, sizeof(float) * out.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(status, d_status, sizeof(status), cudaMemcpyDeviceToHost);
    printf("box %d: barrier status %d %d %d %d\n", ROW_FLOATS, status[0], status[1], status[2], status[3]);
    int bad = 0;
    for (int b = 0; b < blocks && status[b] == 1; ++b)
        for (int i = 0; i < m; ++i)
            for (int k = 0; k < 12; ++k)
                if (out[(size_t)b * m * ROW_FLOATS + i * ROW_FLOATS + k] != rec[(size_t)gid[i] * 12 + k]) ++bad;
    printf("box %d: mismatches %d of %d\n", ROW_FLOATS, bad, blocks * m * 12);
    return bad != 0 || status[0] != 1;
}

int main() {
    const int n = 100003, m = 64;
    std::vector<float> rec((size_t)n * 12);
    for (size_t i = 0; i < rec.size(); ++i) rec[i] = (float)(i % 9973) + 0.25f;
    std::vector<int> gid(m);
    srand(5);
    for (int i = 0; i < m; ++i) gid[i] = rand() % n;
    gid[3] = n - 1;   // last row: columns 12..15 of variant B run past the end of the array
    gid[7] = 0;
    float* d_rec;
    int* d_gid;
    cudaMalloc(&d_rec, sizeof(float) * rec.size());
    cudaMalloc(&d_gid, sizeof(int) * m);
    cudaMemcpy(d_rec, rec.data(), sizeof(float) * rec.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_gid, gid.data(), sizeof(int) * m, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s (%d) %p\n", cudaGetErrorString(e), (int)q, fn);
    if (e != cudaSuccess || !fn) return 3;
    int rc = 0;
    rc |= run<16>((EncodeFn)fn, d_rec, n, d_gid, m, rec, gid) << 4;
    rc |= run<12>((EncodeFn)fn, d_rec, n, d_gid, m, rec, gid);
    printf("probe rc=%d\n", rc);
    return 0;
}
