// Per-tile sorted instance lists WITHOUT sorting the M (tile | depth) keys: count, scan, emit, rank.
//
// The 64-bit key order the compositor needs - ascending tile, then ascending fp32 depth bits, ties by ascending
// Gaussian index (DR:527-562 made stable; oracle tile_keys()) - factorises: the Gaussians are ranked ONCE by
// (depth bits, index) (frb_depth_order: N keys, not M), and inside one tile the order of its entries is the order
// of their ranks.  Ranks are unique integers below N, so a tile's list is put in order by a BITMAP, not by a sort:
//
//   frb_tile_count        the Gaussians (in index order) are cut into G chunks; CTA c histograms the tiles its chunk
//                         touches in SHARED memory and writes the histogram row hist[c][*]    (no global atomics:
//                         1024 tile counters share 32 cache lines, and L2 serialises atomics per line - measured
//                         114 us for 770k RED.ADDs at config 2, against 5 us this way)
//   frb_tile_scan         column prefix over the chunks (hist[c][t] <- instances of tile t in chunks < c), then the
//                         exclusive scan of the tile totals -> ranges, M, launch order
//   frb_tile_emit         CTA c again: slot = start[t] + hist[c][t] + (shared-memory cursor)++ ;
//                         inst_rank[slot] = depth rank
//   frb_tile_rank_gather  one CTA per tile: set bit `rank` in a shared-memory bitmap of N bits, popcount-scan the
//                         bitmap, position of an entry = number of set bits below its own; write the Gaussian id
//                         there and gather its 48-byte record into the tile's contiguous span
//
// Against "emit M keys in depth order + two one-sweep radix passes over M (key, id) pairs + ranges + gather"
// (sort.cu, still used when N exceeds the bitmap) this reads and writes 4 bytes per instance where the sort moved
// 2 x 24, needs no look-back chains over M, and produces the same lists bit for bit: the order inside a tile does not
// depend on the (atomic, arbitrary) emission order because the bitmap forgets it.
// HBM-shaped integer work: coalesced 4-byte streams, shared-memory bitmap, one CTA per tile in longest-first order.
#include "frb_common.cuh"

#include <algorithm>

namespace {

constexpr int RANK_THREADS = 512;
constexpr int RANK_WARPS = RANK_THREADS / 32;
constexpr int MAX_RANKED = 1 << 20;          // bitmap of at most 2^20 bits (128 KB) + prefixes fits one SM
constexpr int SCAN_THREADS = 1024;
constexpr int SCHED_BUCKETS = 1024;

__device__ __forceinline__ void rect_tiles(const float4* __restrict__ records, uint32_t g, int& tx0, int& tx1,
                                           int& ty0, int& ty1) {
    const uint32_t lo = __float_as_uint(records[3 * (size_t)g + 1].z);
    const uint32_t hi = __float_as_uint(records[3 * (size_t)g + 1].w) & 0x7fff7fffu;
    const int x0 = lo & 0xffff, y0 = lo >> 16, x1 = hi & 0xffff, y1 = hi >> 16;
    if (x1 <= x0 || y1 <= y0) {              // culled or empty rectangle: no tiles
        tx0 = ty0 = 0;
        tx1 = ty1 = -1;
        return;
    }
    tx0 = x0 / FRB_TILE; tx1 = (x1 - 1) / FRB_TILE; ty0 = y0 / FRB_TILE; ty1 = (y1 - 1) / FRB_TILE;
}

constexpr int CHUNK_THREADS = 1024;
constexpr int MAX_CHUNKS = 148;              // one chunk per SM at most
constexpr int CHUNK_MIN = 1024;              // Gaussians per chunk at least
constexpr int MAX_SMEM_TILES = 48 * 1024;    // tile counters per CTA in shared memory (192 KB)

// CTA c: shared-memory histogram of the tiles touched by the Gaussians [c * chunk, (c + 1) * chunk).
__global__ void __launch_bounds__(CHUNK_THREADS)
tile_count_kernel(int n, int chunk, int n_tiles, int n_per_view, int tiles_x, int tiles_per_view,
                  const float4* __restrict__ records, uint32_t* __restrict__ chunk_hist) {
    frb_pdl_prologue();
    extern __shared__ uint32_t hist_s[];
    for (int t = threadIdx.x; t < n_tiles; t += CHUNK_THREADS) hist_s[t] = 0;
    __syncthreads();
    const int k0 = blockIdx.x * chunk, k1 = min(n, k0 + chunk);
    for (int k = k0 + threadIdx.x; k < k1; k += CHUNK_THREADS) {
        const uint32_t g = (uint32_t)k;
        int tx0, tx1, ty0, ty1;
        rect_tiles(records, g, tx0, tx1, ty0, ty1);
        const uint32_t view_base = (g / (uint32_t)n_per_view) * (uint32_t)tiles_per_view;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&hist_s[view_base + (uint32_t)(ty * tiles_x + tx)], 1u);
    }
    __syncthreads();
    uint32_t* row = chunk_hist + (size_t)blockIdx.x * n_tiles;
    for (int t = threadIdx.x; t < n_tiles; t += CHUNK_THREADS) row[t] = hist_s[t];
}

// Column prefix over the chunks: hist[c][t] <- sum of hist[c'][t] for c' < c; total[t] = column sum.
// A CTA owns 32 tiles; thread (cy, tx): cy = one of 8 contiguous chunk ranges, tx = tile.  Two walks over the own
// range (sum, then write) with a prefix over the 8 ranges in between: 2 * G / 8 coalesced loads per thread.
__global__ void __launch_bounds__(256)
tile_column_prefix_kernel(int n_chunks, int n_tiles, uint32_t* __restrict__ chunk_hist, uint32_t* __restrict__ total) {
    frb_pdl_prologue();
    __shared__ uint32_t part[8][32];
    const int tx = threadIdx.x & 31, cy = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + tx;
    const int per = (n_chunks + 7) / 8;
    const int c0 = min(n_chunks, cy * per), c1 = min(n_chunks, c0 + per);
    uint32_t sum = 0;
    if (t < n_tiles)
        for (int c = c0; c < c1; ++c) sum += chunk_hist[(size_t)c * n_tiles + t];
    part[cy][tx] = sum;
    __syncthreads();
    uint32_t run = 0, all = 0;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        const uint32_t v = part[y][tx];
        if (y < cy) run += v;
        all += v;
    }
    if (t < n_tiles) {
        for (int c = c0; c < c1; ++c) {
            uint32_t* p = chunk_hist + (size_t)c * n_tiles + t;
            const uint32_t v = *p;
            *p = run;
            run += v;
        }
        if (cy == 0) total[t] = all;
    }
}

// One CTA (1024 threads): exclusive scan of the tile counts (ranges, cursors, total) and the launch order of the
// tiles (descending list length in buckets of 8 entries: the long centre tiles start first, the short ones fill the
// tail).  Entries past m_capacity do not exist: ranges are clamped to it and the status says so.
// tile_count may live in shared or global memory.
struct ScanSmem {
    uint32_t warp_tot[32];
    uint32_t carry;
    int hist[SCHED_BUCKETS];
};

__device__ __forceinline__ void tile_scan_body(ScanSmem& sc, int n_tiles, const uint32_t* tile_count,
                                               uint32_t m_capacity, int2* __restrict__ ranges,
                                               int* __restrict__ tile_order, uint32_t* __restrict__ tile_start,
                                               uint32_t* __restrict__ m_out, uint32_t* __restrict__ flags,
                                               const uint32_t* __restrict__ upstream_error) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sc.carry = 0;
    sc.hist[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [](uint32_t c) { return SCHED_BUCKETS - 1 - (int)min(c >> 3, (uint32_t)(SCHED_BUCKETS - 1)); };
    for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
        const int t = base + threadIdx.x;
        const uint32_t c = (t < n_tiles) ? tile_count[t] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) sc.warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sc.warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += x;
            }
            sc.warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = sc.carry;
        const uint32_t start = carry + sc.warp_tot[warp] + incl - c;
        if (t < n_tiles) {
            const uint32_t s = min(start, m_capacity), e = min(start + c, m_capacity);
            ranges[t] = make_int2((int)s, (int)e);
            tile_start[t] = start;               // unclamped: frb_tile_emit drops slots past the capacity
            atomicAdd(&sc.hist[bucket(e - s)], 1);
        }
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) sc.carry = start + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        m_out[0] = sc.carry;
        // m_out[1]: status the host reads with the count - 1 = lists truncated to the capacity, 2 = the depth sort's
        // look-back gave up (its order is not trustworthy)
        const uint32_t st = (sc.carry > m_capacity ? 1u : 0u) | ((upstream_error && *upstream_error) ? 2u : 0u);
        m_out[1] = st;
        flags[0] = st;
    }
    // exclusive scan of the 1024 schedule buckets (bucket 0 = longest lists), then place the tiles
    {
        int v = sc.hist[threadIdx.x], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) sc.warp_tot[warp] = (uint32_t)incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = sc.warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += x;
            }
            sc.warp_tot[lane] = wi - w;
        }
        __syncthreads();
        sc.hist[threadIdx.x] = (int)sc.warp_tot[warp] + incl - v;
        __syncthreads();
    }
    for (int t = threadIdx.x; t < n_tiles; t += SCAN_THREADS) {
        const uint32_t c = tile_count[t];
        // same length as ranges[t] (read from the counts: the ranges written above by other threads need no fence)
        const uint32_t s = min(tile_start[t], m_capacity), e = min(tile_start[t] + c, m_capacity);
        tile_order[atomicAdd(&sc.hist[bucket(e - s)], 1)] = t;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(int n_tiles, const uint32_t* __restrict__ tile_count, uint32_t m_capacity, int2* __restrict__ ranges,
                 int* __restrict__ tile_order, uint32_t* __restrict__ tile_start, uint32_t* __restrict__ m_out,
                 uint32_t* __restrict__ flags, const uint32_t* __restrict__ upstream_error) {
    frb_pdl_prologue();
    __shared__ ScanSmem sc;
    tile_scan_body(sc, n_tiles, tile_count, m_capacity, ranges, tile_order, tile_start, m_out, flags, upstream_error);
}

// Count, column prefix and tile scan in ONE kernel.  CTA with ticket c histograms the tiles of its chunk in shared
// memory, publishes the row (status[c][t] = count | AGG), obtains "instances of tile t in chunks before c" by decoupled
// look-back over the predecessors' rows (thread = tile, LOOK_WINDOW rows per round trip), publishes the inclusive
// value (| PREFIX) and keeps the exclusive one in base[c][t] for frb_tile_emit.  The CTA with the LAST ticket ends
// up with the column totals of every tile in its shared memory and runs the tile scan on them.  status must be zero.
constexpr uint32_t TFLAG_AGG = 1u << 30, TFLAG_PREFIX = 2u << 30, TFLAG_MASK = 3u << 30;
constexpr int T_LOOK_WINDOW = 8;
constexpr int T_SPIN_LIMIT = 1 << 24;

__device__ __forceinline__ uint32_t t_ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void t_st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(SCAN_THREADS)
tile_count_scan_kernel(int n, int chunk, int n_chunks, int n_tiles, int n_per_view, int tiles_x, int tiles_per_view,
                       const float4* __restrict__ records, uint32_t* __restrict__ status,
                       uint32_t* __restrict__ chunk_base, uint32_t* __restrict__ ticket, uint32_t m_capacity,
                       int2* __restrict__ ranges, int* __restrict__ tile_order, uint32_t* __restrict__ tile_start,
                       uint32_t* __restrict__ m_out, uint32_t* __restrict__ flags,
                       const uint32_t* __restrict__ upstream_error) {
    frb_pdl_prologue();
    extern __shared__ uint32_t hist_s[];
    __shared__ ScanSmem sc;
    __shared__ uint32_t chunk_s;
    if (threadIdx.x == 0) chunk_s = atomicAdd(ticket, 1u);     // chunks are numbered in start order
    for (int t = threadIdx.x; t < n_tiles; t += SCAN_THREADS) hist_s[t] = 0;
    __syncthreads();
    const int c = (int)chunk_s;
    const int k0 = c * chunk, k1 = min(n, k0 + chunk);
    for (int k = k0 + threadIdx.x; k < k1; k += SCAN_THREADS) {
        int tx0, tx1, ty0, ty1;
        rect_tiles(records, (uint32_t)k, tx0, tx1, ty0, ty1);
        const uint32_t view_base = ((uint32_t)k / (uint32_t)n_per_view) * (uint32_t)tiles_per_view;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&hist_s[view_base + (uint32_t)(ty * tiles_x + tx)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += SCAN_THREADS) {
        const uint32_t v = hist_s[t];
        uint32_t excl = 0;
        uint32_t* mine = status + (size_t)c * n_tiles + t;
        if (c == 0) {
            t_st_volatile(mine, v | TFLAG_PREFIX);
        } else {
            t_st_volatile(mine, v | TFLAG_AGG);
            int look = c - 1, spins = 0;
            bool done = false;
            while (!done) {
                uint32_t w[T_LOOK_WINDOW];
#pragma unroll
                for (int q = 0; q < T_LOOK_WINDOW; ++q)
                    w[q] = (look - q >= 0) ? t_ld_volatile(status + (size_t)(look - q) * n_tiles + t) : TFLAG_PREFIX;
#pragma unroll
                for (int q = 0; q < T_LOOK_WINDOW; ++q) {
                    if (done) break;
                    const uint32_t f = w[q] & TFLAG_MASK;
                    if (f == 0) {                       // not published yet: poll again from this row
                        look -= q;
                        if (++spins > T_SPIN_LIMIT) { flags[1] = 1; done = true; }
                        break;
                    }
                    excl += w[q] & ~TFLAG_MASK;
                    if (f == TFLAG_PREFIX) done = true;
                    else if (q == T_LOOK_WINDOW - 1) look -= T_LOOK_WINDOW;
                }
            }
            t_st_volatile(mine, (excl + v) | TFLAG_PREFIX);
        }
        chunk_base[(size_t)c * n_tiles + t] = excl;
        hist_s[t] = excl + v;                           // the last chunk: the column totals
    }
    if (c != n_chunks - 1) return;
    __syncthreads();
    tile_scan_body(sc, n_tiles, hist_s, m_capacity, ranges, tile_order, tile_start, m_out, flags, upstream_error);
}

// CTA c of frb_tile_count again: the slot of an instance = start of its tile's span + instances of the tile in
// earlier chunks + its arrival number inside this chunk (shared-memory cursor).  The arrival order inside a chunk is
// arbitrary; frb_tile_rank_gather does not depend on it.
__global__ void __launch_bounds__(CHUNK_THREADS)
tile_emit_kernel(int n, int chunk, int n_tiles, int n_per_view, int tiles_x, int tiles_per_view,
                 const float4* __restrict__ records, const uint32_t* __restrict__ rank,
                 const uint32_t* __restrict__ chunk_base, const uint32_t* __restrict__ tile_start,
                 uint32_t m_capacity, uint32_t* __restrict__ inst_rank) {
    frb_pdl_prologue();
    extern __shared__ uint32_t cursor_s[];
    const uint32_t* row = chunk_base + (size_t)blockIdx.x * n_tiles;
    for (int t = threadIdx.x; t < n_tiles; t += CHUNK_THREADS) cursor_s[t] = tile_start[t] + row[t];
    __syncthreads();
    const int k0 = blockIdx.x * chunk, k1 = min(n, k0 + chunk);
    for (int k = k0 + threadIdx.x; k < k1; k += CHUNK_THREADS) {
        const uint32_t g = (uint32_t)k;
        int tx0, tx1, ty0, ty1;
        rect_tiles(records, g, tx0, tx1, ty0, ty1);
        if (ty1 < ty0) continue;
        const uint32_t r = rank[g];                         // depth rank: what the bitmap of the tile kernel is over
        const uint32_t view_base = (g / (uint32_t)n_per_view) * (uint32_t)tiles_per_view;
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const uint32_t slot = atomicAdd(&cursor_s[view_base + (uint32_t)(ty * tiles_x + tx)], 1u);
                if (slot < m_capacity) inst_rank[slot] = r;
            }
    }
}

// One CTA per tile.  Shared memory: bitmap[words] | fine[words] (u16: set bits before the word inside its 32-word
// chunk) | coarse[words / 32] (set bits before the chunk).  position(rank) = coarse + fine + popc(word below bit).
template <bool WITH_KEYS>
__global__ void __launch_bounds__(RANK_THREADS)
tile_rank_gather_kernel(int words, const int* __restrict__ tile_order, const int2* __restrict__ ranges,
                        const uint32_t* __restrict__ inst_rank, const uint32_t* __restrict__ order,
                        const float4* __restrict__ records, const uint32_t* __restrict__ depth_bits,
                        const float* __restrict__ phases, uint32_t* __restrict__ sorted_gids,
                        float4* __restrict__ sorted_records, float* __restrict__ sorted_phases,
                        uint64_t* __restrict__ sorted_keys) {
    frb_pdl_prologue();
    extern __shared__ __align__(16) uint32_t smem_words[];
    uint32_t* bitmap = smem_words;
    uint16_t* fine = reinterpret_cast<uint16_t*>(bitmap + words);
    uint32_t* coarse = reinterpret_cast<uint32_t*>(fine + words);
    __shared__ uint32_t coarse_carry;

    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;
    const int2 range = ranges[tile];
    const int count = range.y - range.x;
    if (count <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = words >> 5;
    const uint32_t* my_rank = inst_rank + range.x;

    for (int w = threadIdx.x; w < (words >> 2); w += RANK_THREADS)
        reinterpret_cast<uint4*>(bitmap)[w] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) coarse_carry = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += RANK_THREADS) {
        const uint32_t k = my_rank[i];
        atomicOr(&bitmap[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    // set bits before every word inside its chunk (warp scan of the popcounts), chunk totals
    for (int c = warp; c < chunks; c += RANK_WARPS) {
        const int p = __popc(bitmap[c * 32 + lane]);
        int incl = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        fine[c * 32 + lane] = (uint16_t)(incl - p);
        if (lane == 31) coarse[c] = (uint32_t)incl;
    }
    __syncthreads();
    // exclusive scan of the chunk totals (at most 1024 of them): warp 0, 32 at a time with a carry
    if (warp == 0) {
        uint32_t carry = 0;
        for (int base = 0; base < chunks; base += 32) {
            const int c = base + lane;
            const uint32_t v = (c < chunks) ? coarse[c] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            if (c < chunks) coarse[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    // two entries per thread and round: the dependent loads (rank -> id) of both are in flight together
    for (int i0 = threadIdx.x; i0 < count; i0 += 2 * RANK_THREADS) {
        uint32_t k[2], g[2], pos[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = i0 + q * RANK_THREADS;
            k[q] = (i < count) ? my_rank[i] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            g[q] = order[k[q]];
            const uint32_t w = k[q] >> 5;
            pos[q] = coarse[w >> 5] + fine[w] + __popc(bitmap[w] & ((1u << (k[q] & 31)) - 1u));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (i0 + q * RANK_THREADS < count) {
                sorted_gids[range.x + pos[q]] = g[q];
                if (WITH_KEYS) sorted_keys[range.x + pos[q]] = ((uint64_t)(uint32_t)tile << 32) | depth_bits[g[q]];
            }
        }
    }
    if (!sorted_records) return;     // id lists only: the compositor fetches the records itself (TMA gather4)
    __syncthreads();        // the ids written above are read back by other threads of this CTA
    // 3 threads per instance, one float4 each: the tile's records land contiguous and coalesced; four rounds are
    // issued together so that four id -> record load chains overlap (a 3,000-entry tile is 18 rounds, not 70)
    const int total = 3 * count;
    for (int t0 = threadIdx.x; t0 < total; t0 += 4 * RANK_THREADS) {
        uint32_t g[4];
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q * RANK_THREADS;
            g[q] = (t < total) ? sorted_gids[range.x + t / 3] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q * RANK_THREADS;
            v[q] = records[3 * (size_t)g[q] + (t - 3 * (t / 3))];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + q * RANK_THREADS;
            if (t < total) {
                sorted_records[3 * (size_t)range.x + t] = v[q];
                if (sorted_phases && t % 3 == 0) sorted_phases[range.x + t / 3] = phases[g[q]];
            }
        }
    }
}

inline int bitmap_words(int n) { return ((n + 31) / 32 + 31) & ~31; }          // whole 32-word chunks
inline size_t rank_smem_bytes(int n) {
    const size_t w = (size_t)bitmap_words(n);
    return 4 * w + 2 * w + 4 * (w / 32);
}

int check_views(int n, int n_views, int width, int height) {
    if (n < 0 || n_views < 1 || n_views > FRB_MAX_VIEWS || n % n_views != 0 || width < 1 || height < 1)
        return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    return 0;
}

}  // namespace

extern "C" int frb_tile_lists_max_gaussians(void) { return MAX_RANKED; }

namespace {
struct ChunkPlan {
    int chunks, chunk;
};
ChunkPlan chunk_plan(int n) {
    ChunkPlan p;
    p.chunks = std::max(1, std::min(MAX_CHUNKS, frb_div_up(n, CHUNK_MIN)));
    p.chunk = frb_div_up(std::max(n, 1), p.chunks);
    p.chunks = frb_div_up(std::max(n, 1), p.chunk);
    return p;
}
// workspace words: [chunk_hist / chunk_base: chunks * n_tiles][tile_total: n_tiles][tile_start: n_tiles][flags: 4]
// [ticket: 4][status: chunks * n_tiles] (flags .. status are zeroed by frb_tile_count_scan)
struct WsView {
    uint32_t *hist, *total, *start, *flags, *ticket, *status;
};
WsView ws_view(void* workspace, int n, int n_tiles) {
    const ChunkPlan p = chunk_plan(n);
    WsView v;
    v.hist = (uint32_t*)workspace;
    v.total = v.hist + (size_t)p.chunks * n_tiles;
    v.start = v.total + n_tiles;
    v.flags = v.start + n_tiles;
    v.ticket = v.flags + 4;
    v.status = v.ticket + 4;
    return v;
}
unsigned long long g_count_opt_in = 0, g_emit_opt_in = 0, g_count_scan_opt_in = 0;
}  // namespace

extern "C" int frb_tile_lists_max_tiles(void) { return MAX_SMEM_TILES; }

extern "C" size_t frb_tile_lists_workspace_bytes(int n, int n_tiles) {
    if (n < 0) n = 0;
    if (n_tiles < 0) n_tiles = 0;
    return sizeof(uint32_t) * ((size_t)(2 * chunk_plan(n).chunks + 2) * n_tiles + 8);
}

extern "C" int frb_tile_count(int n, int n_views, int width, int height, const float* records,
                              void* workspace, void* stream) {
    int rc = check_views(n, n_views, width, height);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!records || !workspace || frb_misaligned16(records)) return FRB_E_INVALID;
    const int tiles_x = frb_div_up(width, FRB_TILE), tiles_y = frb_div_up(height, FRB_TILE);
    const int n_tiles = n_views * tiles_x * tiles_y;
    if (n_tiles > MAX_SMEM_TILES) return FRB_E_TOO_LARGE;
    const ChunkPlan p = chunk_plan(n);
    const WsView w = ws_view(workspace, n, n_tiles);
    FRB_CUDA_OK(frb_opt_in_smem(tile_count_kernel, 4 * MAX_SMEM_TILES, &g_count_opt_in));
    frb_launch(tile_count_kernel, dim3(p.chunks), dim3(CHUNK_THREADS), sizeof(uint32_t) * (size_t)n_tiles,
               (cudaStream_t)stream, n, p.chunk, n_tiles, n / n_views, tiles_x, tiles_x * tiles_y,
               (const float4*)records, w.hist);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_tile_scan(int n, int n_tiles, int m_capacity, int32_t* ranges, int32_t* tile_order,
                             uint32_t* m_out, const uint32_t* upstream_error, void* workspace, void* stream) {
    if (n < 0 || n_tiles < 0 || m_capacity < 0) return FRB_E_INVALID;
    if (n_tiles == 0) return 0;
    if (!ranges || !tile_order || !m_out || !workspace) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const ChunkPlan p = chunk_plan(n);
    const WsView w = ws_view(workspace, n, n_tiles);
    if (n == 0) {
        FRB_CUDA_OK(cudaMemsetAsync(w.total, 0, sizeof(uint32_t) * (size_t)n_tiles, st));
    } else {
        frb_launch(tile_column_prefix_kernel, dim3(frb_div_up(n_tiles, 32)), dim3(256), 0, st, p.chunks, n_tiles,
                   w.hist, w.total);
        frb_note_launches(1);
    }
    frb_launch(tile_scan_kernel, dim3(1), dim3(SCAN_THREADS), 0, st, n_tiles, (const uint32_t*)w.total,
               (uint32_t)m_capacity, (int2*)ranges, tile_order, w.start, m_out, w.flags, upstream_error);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// frb_tile_count + frb_tile_scan in one kernel launch (what the renderers call): the chunk CTAs obtain their column
// prefixes by decoupled look-back and the last one runs the tile scan.  Same outputs as the pair.
extern "C" int frb_tile_count_scan(int n, int n_views, int width, int height, const float* records, int m_capacity,
                                   int32_t* ranges, int32_t* tile_order, uint32_t* m_out,
                                   const uint32_t* upstream_error, void* workspace, void* stream) {
    int rc = check_views(n, n_views, width, height);
    if (rc) return rc;
    if (m_capacity < 0) return FRB_E_INVALID;
    const int tiles_x = frb_div_up(width, FRB_TILE), tiles_y = frb_div_up(height, FRB_TILE);
    const int n_tiles = n_views * tiles_x * tiles_y;
    if (n_tiles > MAX_SMEM_TILES) return FRB_E_TOO_LARGE;
    if (!ranges || !tile_order || !m_out || !workspace) return FRB_E_INVALID;
    if (n > 0 && (!records || frb_misaligned16(records))) return FRB_E_INVALID;
    if (n == 0) {
        rc = frb_tile_count(n, n_views, width, height, records, workspace, stream);
        return rc ? rc : frb_tile_scan(n, n_tiles, m_capacity, ranges, tile_order, m_out, upstream_error, workspace, stream);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const ChunkPlan p = chunk_plan(n);
    const WsView w = ws_view(workspace, n, n_tiles);
    FRB_CUDA_OK(cudaMemsetAsync(w.flags, 0, sizeof(uint32_t) * (8 + (size_t)p.chunks * n_tiles), st));
    FRB_CUDA_OK(frb_opt_in_smem(tile_count_scan_kernel, 4 * MAX_SMEM_TILES, &g_count_scan_opt_in));
    frb_launch(tile_count_scan_kernel, dim3(p.chunks), dim3(SCAN_THREADS), sizeof(uint32_t) * (size_t)n_tiles, st, n,
               p.chunk, p.chunks, n_tiles, n / n_views, tiles_x, tiles_x * tiles_y, (const float4*)records, w.status,
               w.hist, w.ticket, (uint32_t)m_capacity, (int2*)ranges, tile_order, w.start, m_out, w.flags,
               upstream_error);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_tile_emit(int n, int n_views, int width, int height, const float* records,
                             const uint32_t* rank, int m_capacity, void* workspace, uint32_t* inst_rank,
                             void* stream) {
    int rc = check_views(n, n_views, width, height);
    if (rc) return rc;
    if (n == 0 || m_capacity <= 0) return 0;
    if (!records || !rank || !workspace || !inst_rank || frb_misaligned16(records)) return FRB_E_INVALID;
    const int tiles_x = frb_div_up(width, FRB_TILE), tiles_y = frb_div_up(height, FRB_TILE);
    const int n_tiles = n_views * tiles_x * tiles_y;
    if (n_tiles > MAX_SMEM_TILES) return FRB_E_TOO_LARGE;
    const ChunkPlan p = chunk_plan(n);
    const WsView w = ws_view(workspace, n, n_tiles);
    FRB_CUDA_OK(frb_opt_in_smem(tile_emit_kernel, 4 * MAX_SMEM_TILES, &g_emit_opt_in));
    frb_launch(tile_emit_kernel, dim3(p.chunks), dim3(CHUNK_THREADS), sizeof(uint32_t) * (size_t)n_tiles,
               (cudaStream_t)stream, n, p.chunk, n_tiles, n / n_views, tiles_x, tiles_x * tiles_y,
               (const float4*)records, rank, (const uint32_t*)w.hist, (const uint32_t*)w.start,
               (uint32_t)m_capacity, inst_rank);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_tile_rank_gather(int n, int n_tiles, const int32_t* tile_order, const int32_t* ranges,
                                    const uint32_t* inst_rank, const uint32_t* order, const float* records,
                                    const uint32_t* depth_bits, const float* phases, uint32_t* sorted_gids,
                                    float* sorted_records, float* sorted_phases, uint64_t* sorted_keys,
                                    void* stream) {
    if (n < 0 || n_tiles < 0) return FRB_E_INVALID;
    if (n > MAX_RANKED) return FRB_E_TOO_LARGE;
    if (n == 0 || n_tiles == 0) return 0;
    if (!ranges || !inst_rank || !order || !sorted_gids) return FRB_E_INVALID;
    if (sorted_records && !records) return FRB_E_INVALID;        // sorted_records NULL: id lists only
    if ((phases == nullptr) != (sorted_phases == nullptr) || (sorted_phases && !sorted_records)) return FRB_E_INVALID;
    if (sorted_keys && !depth_bits) return FRB_E_INVALID;
    if (frb_misaligned16(records) || frb_misaligned16(sorted_records)) return FRB_E_INVALID;
    const int words = bitmap_words(n);
    const size_t smem = rank_smem_bytes(n);
    static unsigned long long opted_in[2] = {0, 0};       // per-device bitmasks, one per instantiation
    if (sorted_keys) {
        FRB_CUDA_OK(frb_opt_in_smem(tile_rank_gather_kernel<true>, (int)rank_smem_bytes(MAX_RANKED), &opted_in[0]));
        frb_launch(tile_rank_gather_kernel<true>, dim3(n_tiles), dim3(RANK_THREADS), smem, (cudaStream_t)stream,
                   words, tile_order, (const int2*)ranges, inst_rank, order, (const float4*)records, depth_bits,
                   phases, sorted_gids, (float4*)sorted_records, sorted_phases, sorted_keys);
    } else {
        FRB_CUDA_OK(frb_opt_in_smem(tile_rank_gather_kernel<false>, (int)rank_smem_bytes(MAX_RANKED), &opted_in[1]));
        frb_launch(tile_rank_gather_kernel<false>, dim3(n_tiles), dim3(RANK_THREADS), smem, (cudaStream_t)stream,
                   words, tile_order, (const int2*)ranges, inst_rank, order, (const float4*)records, depth_bits,
                   phases, sorted_gids, (float4*)sorted_records, sorted_phases, sorted_keys);
    }
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
