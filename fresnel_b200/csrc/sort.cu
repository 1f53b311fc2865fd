// Tile binning: depth order, instance offsets, (tile | depth) key emission, stable LSD radix
// sort, per-tile ranges and the record gather.
//
// Restates torch.argsort(depths) + mask compaction (DR:527-562) as the 64-bit key sort the
// tile compositor needs: key = (global_tile << 32) | fp32 depth bits, ties by ascending Gaussian
// index (stable).  The pipeline exploits that an LSD radix sort only has to touch the digits
// that are not already ordered: Gaussians are first put in stable depth order (32-bit keys, N
// elements), instances are emitted in that order, and the 64-bit keys are then sorted on the
// tile bits only.  The result is bit-identical to a full 64-bit stable sort of instances emitted
// in index order (tests/test_binning.py checks both against oracle tile_keys()).
// Nearest native code in the reference: src/core/compute/radix_sort.cpp (32-bit, unstable).
//
// All kernels here are HBM/L2-bound integer work: coalesced loads, shared-memory digit counters,
// match.any ranking for stability.
#include "frb_common.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_IPT = 8;                                // items per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;         // 2048 items per block
constexpr int SORT_WARP_ITEMS = 32 * SORT_IPT;

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, uint32_t mask) {
    return (uint32_t)(k >> shift) & mask;
}

// Per-block digit histogram, stored digit-major: hist[d * n_blocks + block].
template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(int m, const KeyT* __restrict__ keys, int shift, uint32_t mask, uint32_t* __restrict__ hist,
                  int n_blocks) {
    __shared__ uint32_t cnt[RADIX];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
        long long i = base + r * SORT_THREADS + threadIdx.x;
        if (i < m) atomicAdd(&cnt[digit_of(keys[i], shift, mask)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = cnt[threadIdx.x];
}

// Exclusive scan of `count` uint32 values in place, single block (count = 256 * n_blocks).
__global__ void __launch_bounds__(1024) scan_single_block_kernel(uint32_t* __restrict__ data, int count) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PER = 4;
    for (int base = 0; base < count; base += 1024 * PER) {
        int i0 = base + threadIdx.x * PER;
        uint32_t v[PER];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            v[k] = (i0 + k < count) ? data[i0 + k] : 0u;
            sum += v[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[warp] + incl - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (i0 + k < count) data[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;
        __syncthreads();
    }
}

// Stable scatter.  Warp w of a block owns the contiguous items [w*256, (w+1)*256) of the block's
// tile, visited in 8 rounds of 32 consecutive items, so (warp, round, lane) order is input order.
template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(int m, const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int shift, uint32_t mask,
                     const uint32_t* __restrict__ hist_scanned, int n_blocks) {
    __shared__ uint32_t cnt[SORT_WARPS][RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = lane; d < RADIX; d += 32) cnt[warp][d] = 0;
    __syncwarp();

    long long base = (long long)blockIdx.x * SORT_TILE + warp * SORT_WARP_ITEMS;
    KeyT key[SORT_IPT];
    uint32_t val[SORT_IPT];
    uint32_t rank[SORT_IPT];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
        long long i = base + r * 32 + lane;
        bool valid = i < m;
        key[r] = valid ? keys_in[i] : (KeyT)0;
        val[r] = valid ? vals_in[i] : 0u;
        uint32_t d = valid ? digit_of(key[r], shift, mask) : RADIX;  // invalid lanes match each other only
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d: turn per-warp counts into starting positions (global base + earlier warps)
        int d = threadIdx.x;
        uint32_t run = hist_scanned[(size_t)d * n_blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_IPT; ++r) {
        long long i = base + r * 32 + lane;
        if (i < m) {
            uint32_t d = digit_of(key[r], shift, mask);
            uint32_t pos = cnt[warp][d] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

template <typename KeyT>
int radix_sort_impl(int m, KeyT* keys, uint32_t* vals, KeyT* keys_tmp, uint32_t* vals_tmp, int begin_bit,
                    int end_bit, uint32_t* hist, cudaStream_t st) {
    if (m <= 0 || end_bit <= begin_bit) return 0;
    const int n_blocks = frb_div_up(m, SORT_TILE);
    KeyT* kin = keys; uint32_t* vin = vals; KeyT* kout = keys_tmp; uint32_t* vout = vals_tmp;
    for (int bit = begin_bit; bit < end_bit; bit += RADIX_BITS) {
        int nb = min(RADIX_BITS, end_bit - bit);
        uint32_t mask = (1u << nb) - 1u;
        radix_hist_kernel<KeyT><<<n_blocks, SORT_THREADS, 0, st>>>(m, kin, bit, mask, hist, n_blocks);
        scan_single_block_kernel<<<1, 1024, 0, st>>>(hist, RADIX * n_blocks);
        radix_scatter_kernel<KeyT><<<n_blocks, SORT_THREADS, 0, st>>>(m, kin, vin, kout, vout, bit, mask, hist,
                                                                      n_blocks);
        frb_note_launches(3);
        FRB_LAUNCH_CHECK();
        KeyT* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        FRB_CUDA_OK(cudaMemcpyAsync(keys, kin, sizeof(KeyT) * (size_t)m, cudaMemcpyDeviceToDevice, st));
        FRB_CUDA_OK(cudaMemcpyAsync(vals, vin, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

__global__ void iota_copy_kernel(int n, const uint32_t* __restrict__ src, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = src[i];
        vals[i] = (uint32_t)i;
    }
}

// ---- exclusive scan of tiles_touched[order[k]] : three phases ---------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total) {
    __shared__ uint32_t ws[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t s = ws[w];
        if (w < warp) off += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return off + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
offsets_reduce_kernel(int n, const uint32_t* __restrict__ touched, const uint32_t* __restrict__ order,
                      uint32_t* __restrict__ block_sums) {
    long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        long long i = base + k;
        if (i < n) s += touched[order ? order[i] : (uint32_t)i];
    }
    uint32_t tot;
    block_exclusive_scan_256(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS)
offsets_write_kernel(int n, const uint32_t* __restrict__ touched, const uint32_t* __restrict__ order,
                     const uint32_t* __restrict__ block_sums_scanned, uint32_t* __restrict__ offsets) {
    long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        long long i = base + k;
        v[k] = (i < n) ? touched[order ? order[i] : (uint32_t)i] : 0u;
        s += v[k];
    }
    uint32_t tot;
    uint32_t excl = block_exclusive_scan_256(s, &tot) + block_sums_scanned[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        long long i = base + k;
        if (i < n) offsets[i] = excl;
        excl += v[k];
        if (i == n - 1) offsets[n] = excl;
    }
}

__global__ void __launch_bounds__(256)
bin_emit_kernel(int n, int n_per_view, int tiles_x, int tiles_per_view, const float4* __restrict__ records,
                const uint32_t* __restrict__ depth_bits, const uint32_t* __restrict__ order,
                const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys, uint32_t* __restrict__ gids) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t off = offsets[k], end = offsets[k + 1];
    if (end == off) return;
    uint32_t g = order ? order[k] : (uint32_t)k;
    uint32_t lo = __float_as_uint(records[3 * (size_t)g + 1].w);
    uint32_t hi = __float_as_uint(records[3 * (size_t)g + 2].w) & 0x7fff7fffu;
    int x0 = lo & 0xffff, y0 = lo >> 16, x1 = hi & 0xffff, y1 = hi >> 16;
    int tx0 = x0 / FRB_TILE, tx1 = (x1 - 1) / FRB_TILE, ty0 = y0 / FRB_TILE, ty1 = (y1 - 1) / FRB_TILE;
    uint64_t view_base = (uint64_t)(g / (uint32_t)n_per_view) * (uint64_t)tiles_per_view;
    uint64_t db = depth_bits[g];
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) {
            uint64_t tile = view_base + (uint64_t)(ty * tiles_x + tx);
            keys[off] = (tile << 32) | db;
            gids[off] = g;
            ++off;
        }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(int m, const uint64_t* __restrict__ keys, int2* __restrict__ ranges) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0) {
        ranges[t].x = 0;
    } else {
        uint32_t p = (uint32_t)(keys[i - 1] >> 32);
        if (p != t) {
            ranges[p].y = i;
            ranges[t].x = i;
        }
    }
    if (i == m - 1) ranges[t].y = m;
}

// 3 threads per instance, one float4 each: 48-byte records land contiguous and coalesced.
__global__ void __launch_bounds__(256)
gather_records_kernel(int m, const uint32_t* __restrict__ gids, const float4* __restrict__ records,
                      float4* __restrict__ sorted_records, const float* __restrict__ phases,
                      float* __restrict__ sorted_phases) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3ll * m) return;
    int i = (int)(t / 3), part = (int)(t - 3ll * i);
    uint32_t g = gids[i];
    sorted_records[t] = records[3 * (size_t)g + part];
    if (sorted_phases && part == 0) sorted_phases[i] = phases[g];
}

}  // namespace

extern "C" size_t frb_sort_workspace_bytes(int m) {
    if (m < 0) m = 0;
    return sizeof(uint32_t) * (size_t)RADIX * (size_t)(frb_div_up(m, SORT_TILE) + 1);
}

extern "C" int frb_radix_sort_pairs(int m, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                                    uint32_t* vals_tmp, int begin_bit, int end_bit, void* workspace,
                                    void* stream) {
    if (m < 0 || begin_bit < 0 || end_bit > 64) return FRB_E_INVALID;
    if (m == 0 || end_bit <= begin_bit) return 0;
    if (!keys || !vals || !keys_tmp || !vals_tmp || !workspace) return FRB_E_INVALID;
    return radix_sort_impl<uint64_t>(m, keys, vals, keys_tmp, vals_tmp, begin_bit, end_bit, (uint32_t*)workspace,
                                     (cudaStream_t)stream);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t frb_depth_order_workspace_bytes(int n) {
    if (n < 0) n = 0;
    return 3 * align256(sizeof(uint32_t) * (size_t)n) + align256(frb_sort_workspace_bytes(n));
}

extern "C" int frb_depth_order(int n, const uint32_t* depth_bits, uint32_t* order, void* workspace,
                               void* stream) {
    if (n < 0) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!depth_bits || !order || !workspace) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    size_t a = align256(sizeof(uint32_t) * (size_t)n);
    uint32_t* keys = (uint32_t*)w;
    uint32_t* keys_tmp = (uint32_t*)(w + a);
    uint32_t* vals_tmp = (uint32_t*)(w + 2 * a);
    uint32_t* hist = (uint32_t*)(w + 3 * a);
    iota_copy_kernel<<<frb_div_up(n, 256), 256, 0, st>>>(n, depth_bits, keys, order);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return radix_sort_impl<uint32_t>(n, keys, order, keys_tmp, vals_tmp, 0, 32, hist, st);
}

extern "C" size_t frb_scan_workspace_bytes(int n) {
    if (n < 0) n = 0;
    return sizeof(uint32_t) * (size_t)(frb_div_up(n, SCAN_TILE) + 1);
}

extern "C" int frb_tile_offsets(int n, const uint32_t* tiles_touched, const uint32_t* order,
                                uint32_t* offsets, void* workspace, void* stream) {
    if (n < 0 || !offsets) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        FRB_CUDA_OK(cudaMemsetAsync(offsets, 0, sizeof(uint32_t), st));
        return 0;
    }
    if (!tiles_touched || !workspace) return FRB_E_INVALID;
    int nb = frb_div_up(n, SCAN_TILE);
    uint32_t* sums = (uint32_t*)workspace;
    offsets_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(n, tiles_touched, order, sums);
    scan_single_block_kernel<<<1, 1024, 0, st>>>(sums, nb);
    offsets_write_kernel<<<nb, SCAN_THREADS, 0, st>>>(n, tiles_touched, order, sums, offsets);
    frb_note_launches(3);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_bin_emit(int n, int n_views, int width, int height, const float* records,
                            const uint32_t* depth_bits, const uint32_t* order, const uint32_t* offsets,
                            uint64_t* keys, uint32_t* gids, void* stream) {
    if (n < 0 || n_views < 1 || n_views > FRB_MAX_VIEWS || n % n_views != 0) return FRB_E_INVALID;
    if (width < 1 || height < 1) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    if (n == 0) return 0;
    if (!records || !depth_bits || !offsets || !keys || !gids) return FRB_E_INVALID;
    int tiles_x = frb_div_up(width, FRB_TILE), tiles_y = frb_div_up(height, FRB_TILE);
    bin_emit_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, n / n_views, tiles_x, tiles_x * tiles_y, (const float4*)records, depth_bits, order, offsets, keys,
        gids);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_tile_ranges(int m, const uint64_t* keys, int n_tiles, int32_t* ranges, void* stream) {
    if (m < 0 || n_tiles < 0 || !ranges) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    FRB_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)n_tiles, st));
    if (m == 0) return 0;
    if (!keys) return FRB_E_INVALID;
    tile_ranges_kernel<<<frb_div_up(m, 256), 256, 0, st>>>(m, keys, (int2*)ranges);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_gather_records(int m, const uint32_t* gids, const float* records, float* sorted_records,
                                  const float* phases, float* sorted_phases, void* stream) {
    if (m < 0) return FRB_E_INVALID;
    if (m == 0) return 0;
    if (!gids || !records || !sorted_records) return FRB_E_INVALID;
    if ((phases == nullptr) != (sorted_phases == nullptr)) return FRB_E_INVALID;
    gather_records_kernel<<<frb_div_up(3ll * m, 256), 256, 0, (cudaStream_t)stream>>>(
        m, gids, (const float4*)records, (float4*)sorted_records, phases, sorted_phases);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
