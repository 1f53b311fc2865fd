// Tile binning: depth order, instance offsets, (tile | depth) key emission, stable LSD radix
// sort, per-tile ranges and the record gather.
//
// Restates torch.argsort(depths) + mask compaction (DR:527-562) as the 64-bit key sort the
// tile compositor needs: key = (global_tile << 32) | fp32 depth bits, ties by ascending Gaussian
// index (stable).  The pipeline exploits that an LSD radix sort only has to touch the digits
// that are not already ordered: Gaussians are first put in stable depth order (32-bit keys, N
// elements), instances are emitted in that order, and the 64-bit keys are then sorted on the
// tile bits only.  The result is bit-identical to a full 64-bit stable sort of instances emitted
// in index order (tests/test_gpu_parity.py checks both against oracle tile_keys()).
// Nearest native code in the reference: src/core/compute/radix_sort.cpp (32-bit, unstable,
// three dispatches and a host round trip per pass).
//
// Sort structure ("one sweep"): ONE kernel builds the global digit histograms of every pass,
// then ONE kernel per 8-bit digit ranks its 2048-key tile with match.any (stable), obtains the
// tile's per-digit offsets by decoupled look-back over the preceding tiles' published counts,
// and scatters.  Keys and values are read once and written once per pass: HBM/L2-bound integer
// work with coalesced loads and shared-memory digit counters.
#include "frb_common.cuh"

#include <cooperative_groups.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int RADIX_BITS_MAX = 9;                          // frb_depth_order_range: 27 key bits in three 9-bit passes
constexpr int RADIX_MAX = 1 << RADIX_BITS_MAX;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_IPT = 8;                                // items per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;         // 2048 items per block
constexpr int SORT_WARP_ITEMS = 32 * SORT_IPT;
constexpr int MAX_PASSES = 8;
constexpr int SORT_GRID_MAX = 148 * 6;                     // persistent grid: 6 resident blocks per SM

constexpr uint32_t FLAG_AGG = 1u << 30;      // tile count published
constexpr uint32_t FLAG_PREFIX = 2u << 30;   // inclusive prefix over tiles 0..this published
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;
constexpr int LOOK_WINDOW = 8;               // predecessors inspected per look-back round trip.  Wider windows for the
                                             // small-tile sort (196 resident tiles at 100k keys) measured SLOWER on one
                                             // box: depth order 41.7 us (8), 43.6 us (16), 46.5 us (32)
                                             // (profiles/r2_w_look_window*.json).  tools/probes/sort_trace.cu stamps the
                                             // phases of every tile (profiles/r2_y_sort_trace_100k.txt): a pass is ~1 us
                                             // of load + rank, ~3 us of look-back (5 rounds on average, 9 at most, of
                                             // ~0.37 us), ~1.3 us of scatter and ~1.5 us of kernel boundary
constexpr int SPIN_LIMIT = 1 << 24;          // a look-back that spins this long reports an error instead of hanging

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift, uint32_t mask) {
    return (uint32_t)(k >> shift) & mask;
}

struct PassPlan {
    int n_passes;
    int shift[MAX_PASSES];
    uint32_t mask[MAX_PASSES];
};

// Key transform of the first pass (and of the histogram kernel) for frb_depth_order_range: depth bits clamped to
// [lo, hi] (the bit patterns of the near and far planes; positive floats order like their bits) minus lo.  Visible
// Gaussians (near < depth < far, DR:541) keep their exact relative order in hi - lo + 1 values; culled ones (behind
// the camera: sign bit set, i.e. negative as int32; beyond far; NaN) collapse to the ends, where their mutual order
// does not matter - they never reach a tile.  on = 0: identity.
struct KeyRange {
    uint32_t lo, hi;
    int on;
    int ballot;     // this pass finds equal digits in a warp with ballots instead of match.any (radix_onesweep_kernel)
};
template <typename KeyT>
__device__ __forceinline__ KeyT key_xform(KeyT k, const KeyRange& kr) { return k; }
template <>
__device__ __forceinline__ uint32_t key_xform<uint32_t>(uint32_t k, const KeyRange& kr) {
    if (!kr.on) return k;
    if ((int32_t)k <= (int32_t)kr.lo) return 0u;
    return (k >= kr.hi ? kr.hi : k) - kr.lo;
}

// Sort workspace layout (uint32 words): [hist: MAX_PASSES*RADIX_MAX][tickets: MAX_PASSES][error: 1]
// [pad to 16 words][status: n_passes * n_blocks * radix]
constexpr int WS_HIST = 0;
constexpr int WS_TICKET = MAX_PASSES * RADIX_MAX;
constexpr int WS_ERROR = WS_TICKET + MAX_PASSES;
constexpr int WS_STATUS = WS_ERROR + 8;

// Global digit histograms of all passes in one read of the keys.  RDX: digits per pass (256 or 512).
template <typename KeyT, int RDX>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_all_kernel(int m, const uint32_t* __restrict__ m_dev, const KeyT* __restrict__ keys,
                      const __grid_constant__ PassPlan plan, const __grid_constant__ KeyRange kr,
                      uint32_t* __restrict__ hist) {
    frb_pdl_prologue();
    __shared__ uint32_t cnt[MAX_PASSES / (RDX / RADIX)][RDX];      // 8 passes of 256 digits or 4 of 512: 8 KB
    if (m_dev) m = min(m, (int)*m_dev);
    for (int p = 0; p < plan.n_passes; ++p)
        for (int d = threadIdx.x; d < RDX; d += SORT_THREADS) cnt[p][d] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * SORT_THREADS + threadIdx.x; i < m;
         i += (long long)gridDim.x * SORT_THREADS) {
        KeyT k = key_xform<KeyT>(keys[i], kr);
        for (int p = 0; p < plan.n_passes; ++p) atomicAdd(&cnt[p][digit_of(k, plan.shift[p], plan.mask[p])], 1u);
    }
    __syncthreads();
    for (int p = 0; p < plan.n_passes; ++p)
        for (int d = threadIdx.x; d < RDX; d += SORT_THREADS) {
            uint32_t c = cnt[p][d];
            if (c) atomicAdd(&hist[p * RDX + d], c);
        }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pass: rank (stable), look back, scatter.  Warp w of a block owns the contiguous items
// [w*32*IPT, (w+1)*32*IPT) of the block's tile, visited in IPT rounds of 32 consecutive items, so
// (tile, warp, round, lane) order is input order.  GEN_VALS: values are the input indices.
// RDX digits per pass (256 or 512); thread t owns the digits [t * DPT, (t + 1) * DPT), DPT = RDX / 256.
// XFORM: the keys read are raw depth bits and are transformed by key_xform (first pass of frb_depth_order_range).
// FRB_SORT_TRACE (tools/probes/sort_trace.cu only): thread 0 of every tile stamps %globaltimer at the phase boundaries.
#ifdef FRB_SORT_TRACE
__device__ unsigned long long frb_sort_trace[4][1024][8];
__device__ __forceinline__ unsigned long long frb_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FRB_TRACE(slot) do { if (threadIdx.x == 0 && tile_tr < 1024) frb_sort_trace[(shift >> 3) & 3][tile_tr][slot] = frb_gtime(); } while (0)
#else
#define FRB_TRACE(slot) do { } while (0)
#endif

template <typename KeyT, bool GEN_VALS, int IPT, int RDX, bool XFORM>
__global__ void __launch_bounds__(SORT_THREADS)
radix_onesweep_kernel(int m, const uint32_t* __restrict__ m_dev, const KeyT* __restrict__ keys_in,
                      const uint32_t* __restrict__ vals_in,
                      KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int shift, uint32_t mask,
                      const uint32_t* __restrict__ hist_pass, uint32_t* __restrict__ status,
                      uint32_t* __restrict__ ticket, uint32_t* __restrict__ error_flag,
                      uint32_t* __restrict__ rank_out, const __grid_constant__ KeyRange kr) {
#ifdef FRB_SORT_TRACE
    const unsigned long long t_entry = frb_gtime();
    uint32_t tile_tr = 0xffffffffu;
    int rounds_tr = 0;
#endif
    frb_pdl_prologue();
    constexpr int DPT = RDX / SORT_THREADS;
    __shared__ uint32_t cnt[SORT_WARPS][RDX];
    __shared__ uint32_t digit_base[RDX];
    __shared__ uint32_t scan_ws[SORT_WARPS];
    __shared__ uint32_t tile_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (m_dev) m = min(m, (int)*m_dev);                     // capacity mode: the true count lives on the device
    // persistent blocks: keep drawing tiles (numbered in start order, so every predecessor of a tile has
    // started) until the keys run out
    while (true) {
    __syncthreads();                                        // shared arrays of the previous tile are free
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    for (int d = lane; d < RDX; d += 32) cnt[warp][d] = 0;
    __syncthreads();
    const uint32_t tile = tile_s;
    if ((long long)tile * (SORT_THREADS * IPT) >= m) return;           // past the end: nobody looks back at this tile
#ifdef FRB_SORT_TRACE
    tile_tr = tile;
    if (threadIdx.x == 0 && tile_tr < 1024) frb_sort_trace[(shift >> 3) & 3][tile_tr][0] = t_entry;
#endif
    FRB_TRACE(1);       // ticket drawn, counters cleared

    long long base = (long long)tile * (SORT_THREADS * IPT) + warp * (32 * IPT);
    KeyT key[IPT];
    uint32_t rank[IPT];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        long long i = base + r * 32 + lane;
        key[r] = (i < m) ? keys_in[i] : (KeyT)0;
        if (XFORM) key[r] = key_xform<KeyT>(key[r], kr);
    }
    FRB_TRACE(2);       // keys requested (loads in flight)
    // the match.any operations are independent: issue them back to back, then update the counters
    uint32_t peers[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        long long i = base + r * 32 + lane;
        uint32_t d = (i < m) ? digit_of(key[r], shift, mask) : RDX;     // invalid lanes match each other only
        // lanes with my digit from one ballot per digit bit: MATCH.ANY iterates over the distinct values of the warp
        // (up to 32 for an 8-bit digit), the ballots are a fixed log2(RDX) + 1
        // - faster for the spread-out low digits of depth keys (depth order of 100k keys 42.8 -> 41.2 us), slower
        // where a warp holds few distinct digits (plane indices of the ASM renderer: 65 -> 88 us for 1M keys), so the
        // caller chooses per pass (kr.ballot)
        if (kr.ballot) {
            uint32_t pm = 0xffffffffu;
#pragma unroll
            for (int bb = 0; (1 << bb) <= RDX; ++bb) {
                const bool bit = (d >> bb) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                pm &= bit ? bal : ~bal;
            }
            peers[r] = pm;
        } else {
            peers[r] = __match_any_sync(0xffffffffu, d);
        }
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        long long i = base + r * 32 + lane;
        bool valid = i < m;
        uint32_t d = digit_of(key[r], shift, mask);
        int leader = __ffs(peers[r]) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = old + __popc(peers[r]);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers[r] & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    FRB_TRACE(3);       // ranked within the warps

    // thread t owns DPT consecutive digits: per-warp counts -> warp-exclusive offsets; tile count -> look-back
    uint32_t excl[DPT], run[DPT], h[DPT];
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
        const int d = threadIdx.x * DPT + q;
        uint32_t r_ = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = r_;
            r_ += c;
        }
        run[q] = r_;
        excl[q] = 0;
        h[q] = hist_pass[d];
    }
    if (tile == 0) {
#pragma unroll
        for (int q = 0; q < DPT; ++q) st_volatile_u32(status + threadIdx.x * DPT + q, run[q] | FLAG_PREFIX);
    } else {
#pragma unroll
        for (int q = 0; q < DPT; ++q)
            st_volatile_u32(status + (size_t)tile * RDX + threadIdx.x * DPT + q, run[q] | FLAG_AGG);
        // Look back LOOK_WINDOW predecessors per round trip: the loads are independent, so a walk over k
        // published counts costs k / LOOK_WINDOW L2 latencies instead of k.  The DPT digits of a thread walk
        // together (all their loads of a round are issued before the first is used).
        FRB_TRACE(4);   // count published
        long long look[DPT];
        bool done[DPT];
        int spins = 0;
#pragma unroll
        for (int q = 0; q < DPT; ++q) { look[q] = (long long)tile - 1; done[q] = false; }
        bool all_done = false;
        while (!all_done) {
#ifdef FRB_SORT_TRACE
            ++rounds_tr;
#endif
            uint32_t v[DPT][LOOK_WINDOW];
#pragma unroll
            for (int q = 0; q < DPT; ++q) {
                const int d = threadIdx.x * DPT + q;
#pragma unroll
                for (int k = 0; k < LOOK_WINDOW; ++k)
                    v[q][k] = (!done[q] && look[q] - k >= 0)
                                  ? ld_volatile_u32(status + (size_t)(look[q] - k) * RDX + d) : FLAG_PREFIX;
            }
            all_done = true;
#pragma unroll
            for (int q = 0; q < DPT; ++q) {
                if (done[q]) continue;
#pragma unroll
                for (int k = 0; k < LOOK_WINDOW; ++k) {
                    if (done[q]) break;
                    const uint32_t f = v[q][k] & FLAG_MASK;
                    if (f == 0) {                      // not published yet: poll again from this tile
                        look[q] -= k;
                        if (++spins > SPIN_LIMIT) { *error_flag = 1; done[q] = true; }
                        break;
                    }
                    excl[q] += v[q][k] & VALUE_MASK;
                    if (f == FLAG_PREFIX) done[q] = true;
                    else if (k == LOOK_WINDOW - 1) look[q] -= LOOK_WINDOW;
                }
                all_done = all_done && done[q];
            }
        }
#pragma unroll
        for (int q = 0; q < DPT; ++q)
            st_volatile_u32(status + (size_t)tile * RDX + threadIdx.x * DPT + q, (excl[q] + run[q]) | FLAG_PREFIX);
    }
    FRB_TRACE(5);       // look-back done, prefix published
    // exclusive scan of the global digit histogram (RDX values, DPT consecutive ones per thread)
    uint32_t tsum = 0;
#pragma unroll
    for (int q = 0; q < DPT; ++q) tsum += h[q];
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) scan_ws[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w)
        if (w < warp) woff += scan_ws[w];
    uint32_t before = woff + incl - tsum;
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
        digit_base[threadIdx.x * DPT + q] = before + excl[q];
        before += h[q];
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        long long i = base + r * 32 + lane;
        if (i < m) {
            uint32_t dg = digit_of(key[r], shift, mask);
            uint32_t pos = digit_base[dg] + cnt[warp][dg] + rank[r];
            keys_out[pos] = key[r];
            const uint32_t val = GEN_VALS ? (uint32_t)i : vals_in[i];
            vals_out[pos] = val;
            if (rank_out) rank_out[val] = pos;            // last pass of frb_depth_order: the inverse permutation
        }
    }
    FRB_TRACE(6);       // scattered
#ifdef FRB_SORT_TRACE
    if (threadIdx.x == 0 && tile_tr < 1024) frb_sort_trace[(shift >> 3) & 3][tile_tr][7] = (unsigned long long)rounds_tr;
#endif
    }   // while: next tile
}

// Small sorts (the depth order of one view's Gaussians) use 512-key tiles: a 100k-key pass is a chain of L2 round
// trips, not bandwidth, and four times as many, four times shorter tiles put every SM to work on it.
// (1024-key tiles instead: 42.0 / 43.3 us against 43.2 / 46.5 us for the depth order of 100k keys on one box,
// profiles/r2_v_sort_tile*.json - no difference: a pass is launch + a fixed chain of dependent L2 round trips.)
constexpr int SORT_IPT_SMALL = 2;
constexpr int SORT_TILE_SMALL = SORT_THREADS * SORT_IPT_SMALL;
constexpr int SORT_SMALL_MAX = 1 << 19;                    // below this many keys: small tiles

inline int sort_tile_of(int m) { return m <= SORT_SMALL_MAX ? SORT_TILE_SMALL : SORT_TILE; }

size_t sort_ws_words(int m, int n_passes, int radix = RADIX) {
    return (size_t)WS_STATUS + (size_t)n_passes * (size_t)frb_div_up(m, sort_tile_of(m)) * radix;
}

// Sorts on bits [begin_bit, end_bit).  Pass p reads buffer (p even ? A : B) and writes the other;
// when first_in is given, pass 0 reads keys from there (and generates values if vals_first is null).
// Returns in *result_in_b whether the sorted data ended in the B buffers.
// RB: digit width (8, or 9 for frb_depth_order_range); kr: key transform applied to first_keys (kr.on).
template <typename KeyT, int RB>
int radix_sort_impl(int m, const uint32_t* m_dev, const KeyT* first_keys, const uint32_t* first_vals, KeyT* keys_a, uint32_t* vals_a,
                    KeyT* keys_b, uint32_t* vals_b, int begin_bit, int end_bit, uint32_t* ws, cudaStream_t st,
                    bool* result_in_b, bool hist_ready = false, uint32_t* rank_out = nullptr,
                    KeyRange kr = KeyRange{0u, 0u, 0, 0}) {
    constexpr int RDX = 1 << RB;
    PassPlan plan;
    plan.n_passes = 0;
    for (int bit = begin_bit; bit < end_bit; bit += RB) {
        if (plan.n_passes == MAX_PASSES / (RDX / RADIX)) return FRB_E_INVALID;
        int nb = min(RB, end_bit - bit);
        plan.shift[plan.n_passes] = bit;
        plan.mask[plan.n_passes] = (1u << nb) - 1u;
        ++plan.n_passes;
    }
    const bool small = sort_tile_of(m) == SORT_TILE_SMALL;
    const int n_blocks = frb_div_up(m, sort_tile_of(m));
    if (!hist_ready) {      // otherwise the caller zeroed ws and the producer of the keys filled the histograms
        FRB_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(uint32_t) * sort_ws_words(m, plan.n_passes, RDX), st));
        frb_launch(radix_hist_all_kernel<KeyT, RDX>, dim3(min(n_blocks, 592)), dim3(SORT_THREADS), 0, st, m, m_dev,
                   first_keys, plan, kr, ws + WS_HIST);
        frb_note_launches(1);
    }
    const KeyT* kin = first_keys;
    const uint32_t* vin = first_vals;
    bool to_b = (first_keys == keys_a);     // in-place start: A -> B; external start: -> A first
    for (int p = 0; p < plan.n_passes; ++p) {
        KeyT* kout = to_b ? keys_b : keys_a;
        uint32_t* vout = to_b ? vals_b : vals_a;
        uint32_t* status = ws + WS_STATUS + (size_t)p * n_blocks * RDX;
        const int grid = min(n_blocks, SORT_GRID_MAX);
        const bool xf = kr.on && p == 0;
        KeyRange kr_pass = kr;
        kr_pass.ballot = (kr.on && plan.mask[p] == (uint32_t)(RDX - 1) && p + 1 < plan.n_passes) ? 1 : 0;
#define FRB_ONESWEEP(GEN, IPT_, XF)                                                                                 \
    frb_launch(radix_onesweep_kernel<KeyT, GEN, IPT_, RDX, XF>, dim3(grid), dim3(SORT_THREADS), 0, st,              \
        m, m_dev, kin, vin, kout, vout, plan.shift[p], plan.mask[p], ws + WS_HIST + p * RDX, status,                \
        ws + WS_TICKET + p, ws + WS_ERROR, (p == plan.n_passes - 1) ? rank_out : (uint32_t*)nullptr, kr_pass)
        if (vin == nullptr) {
            if (xf) { if (small) FRB_ONESWEEP(true, SORT_IPT_SMALL, true); else FRB_ONESWEEP(true, SORT_IPT, true); }
            else { if (small) FRB_ONESWEEP(true, SORT_IPT_SMALL, false); else FRB_ONESWEEP(true, SORT_IPT, false); }
        } else {
            if (small) FRB_ONESWEEP(false, SORT_IPT_SMALL, false); else FRB_ONESWEEP(false, SORT_IPT, false);
        }
#undef FRB_ONESWEEP
        frb_note_launches(1);
        FRB_LAUNCH_CHECK();
        kin = kout; vin = vout;
        *result_in_b = to_b;
        to_b = !to_b;
    }
    return 0;
}

// ---- exclusive scan of tiles_touched[order[k]] : single pass, decoupled look-back ----------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;
constexpr int EMIT_TILE = 256;               // Gaussians per block of scan_emit_kernel (one per thread)
constexpr unsigned long long SFLAG_AGG = 1ull << 62, SFLAG_PREFIX = 2ull << 62, SFLAG_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// workspace: [ticket u64][error u64][status u64 x n_blocks]
__global__ void __launch_bounds__(SCAN_THREADS)
offsets_scan_kernel(int n, const uint32_t* __restrict__ touched, const uint32_t* __restrict__ order,
                    uint32_t* __restrict__ offsets, unsigned long long* __restrict__ ws) {
    frb_pdl_prologue();
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t tile_s;
    __shared__ uint32_t prefix_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = (uint32_t)atomicAdd(ws, 1ull);
    __syncthreads();
    const uint32_t tile = tile_s;
    long long base = (long long)tile * SCAN_TILE + threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        long long i = base + k;
        v[k] = (i < n) ? touched[order ? order[i] : (uint32_t)i] : 0u;
        s += v[k];
    }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t x = wsum[w];
        if (w < warp) woff += x;
        total += x;
    }
    unsigned long long* status = ws + 2;
    if (warp == 0) {
        // warp-wide look-back: 32 predecessors per L2 round trip (lane l inspects tile - 1 - l)
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(status, (unsigned long long)total | SFLAG_PREFIX);
        } else {
            if (lane == 0) st_volatile_u64(status + tile, (unsigned long long)total | SFLAG_AGG);
            long long look = (long long)tile - 1;
            int spins = 0;
            while (true) {
                const long long idx = look - lane;
                const unsigned long long x = (idx >= 0) ? ld_volatile_u64(status + idx) : SFLAG_PREFIX;
                const unsigned long long f = x & SFLAG_MASK;
                const uint32_t unpublished = __ballot_sync(0xffffffffu, f == 0);
                const uint32_t prefixes = __ballot_sync(0xffffffffu, f == SFLAG_PREFIX);
                const int stop = prefixes ? (__ffs(prefixes) - 1) : 31;       // last lane that contributes
                const uint32_t need = (stop == 31) ? 0xffffffffu : ((2u << stop) - 1u);
                if (unpublished & need) {                                     // poll the same window again
                    if (++spins > SPIN_LIMIT) { if (lane == 0) ws[1] = 1; break; }
                    continue;
                }
                uint32_t c = (lane <= stop) ? (uint32_t)(x & ~SFLAG_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (prefixes) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(status + tile, (unsigned long long)(excl + total) | SFLAG_PREFIX);
        }
        if (lane == 0) prefix_s = excl;
    }
    __syncthreads();
    uint32_t excl = prefix_s + woff + incl - s;
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
    frb_pdl_prologue();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t off = offsets[k], end = offsets[k + 1];
    if (end == off) return;
    uint32_t g = order ? order[k] : (uint32_t)k;
    uint32_t lo = __float_as_uint(records[3 * (size_t)g + 1].z);
    uint32_t hi = __float_as_uint(records[3 * (size_t)g + 1].w) & 0x7fff7fffu;
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

// Scan + emit + histogram in ONE kernel (the whole-pass forward, pipeline.cu): block b takes 256 Gaussians in depth
// order, scans their tile counts, obtains its offset by warp-wide decoupled look-back over the preceding blocks,
// emits the (tile | depth) keys, and counts the tile digits of the two sort passes on the way (shared-memory
// counters, flushed once per block) - the offsets array, the offsets kernel and the histogram pass over the M keys
// disappear.  ws: [ticket u64][error u64][status u64 x blocks], zeroed by the caller; hist: the sort workspace's
// histogram words (pass p at hist + 256 p), zeroed by the caller; m_out: total number of instances.
__global__ void __launch_bounds__(EMIT_TILE)
scan_emit_kernel(int n, int n_per_view, int tiles_x, int tiles_per_view, const float4* __restrict__ records,
                 const uint32_t* __restrict__ depth_bits, const uint32_t* __restrict__ touched,
                 const uint32_t* __restrict__ order, uint64_t* __restrict__ keys, uint32_t* __restrict__ gids,
                 unsigned long long* __restrict__ ws, uint32_t* __restrict__ hist, uint32_t mask1,
                 uint32_t* __restrict__ m_out) {
    frb_pdl_prologue();
    __shared__ uint32_t h[2][RADIX];
    __shared__ uint32_t wsum[EMIT_TILE / 32];
    __shared__ uint32_t tile_s, prefix_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = (uint32_t)atomicAdd(ws, 1ull);     // blocks are numbered in start order
    h[0][threadIdx.x] = 0;
    h[1][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t blk = tile_s;
    const long long k = (long long)blk * EMIT_TILE + threadIdx.x;
    uint32_t g = 0, cnt = 0;
    if (k < n) {
        g = order ? order[k] : (uint32_t)k;
        cnt = touched[g];
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < EMIT_TILE / 32; ++w) {
        uint32_t x = wsum[w];
        if (w < warp) woff += x;
        total += x;
    }
    unsigned long long* status = ws + 2;
    if (warp == 0) {
        uint32_t excl = 0;
        if (blk == 0) {
            if (lane == 0) st_volatile_u64(status, (unsigned long long)total | SFLAG_PREFIX);
        } else {
            if (lane == 0) st_volatile_u64(status + blk, (unsigned long long)total | SFLAG_AGG);
            long long look = (long long)blk - 1;
            int spins = 0;
            while (true) {
                const long long idx = look - lane;
                const unsigned long long x = (idx >= 0) ? ld_volatile_u64(status + idx) : SFLAG_PREFIX;
                const unsigned long long f = x & SFLAG_MASK;
                const uint32_t unpublished = __ballot_sync(0xffffffffu, f == 0);
                const uint32_t prefixes = __ballot_sync(0xffffffffu, f == SFLAG_PREFIX);
                const int stop = prefixes ? (__ffs(prefixes) - 1) : 31;
                const uint32_t need = (stop == 31) ? 0xffffffffu : ((2u << stop) - 1u);
                if (unpublished & need) {
                    if (++spins > SPIN_LIMIT) { if (lane == 0) ws[1] = 1; break; }
                    continue;
                }
                uint32_t c = (lane <= stop) ? (uint32_t)(x & ~SFLAG_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (prefixes) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(status + blk, (unsigned long long)(excl + total) | SFLAG_PREFIX);
        }
        if (lane == 0) prefix_s = excl;
    }
    __syncthreads();
    uint32_t off = prefix_s + woff + incl - cnt;
    if (k == n - 1) *m_out = off + cnt;
    if (cnt) {
        uint32_t lo = __float_as_uint(records[3 * (size_t)g + 1].z);
        uint32_t hi = __float_as_uint(records[3 * (size_t)g + 1].w) & 0x7fff7fffu;
        int x0 = lo & 0xffff, y0 = lo >> 16, x1 = hi & 0xffff, y1 = hi >> 16;
        int tx0 = x0 / FRB_TILE, tx1 = (x1 - 1) / FRB_TILE, ty0 = y0 / FRB_TILE, ty1 = (y1 - 1) / FRB_TILE;
        uint32_t view_base = (g / (uint32_t)n_per_view) * (uint32_t)tiles_per_view;
        uint64_t db = depth_bits[g];
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                uint32_t tile = view_base + (uint32_t)(ty * tiles_x + tx);
                keys[off] = ((uint64_t)tile << 32) | db;
                gids[off] = g;
                ++off;
                atomicAdd(&h[0][tile & (RADIX - 1)], 1u);
                atomicAdd(&h[1][(tile >> RADIX_BITS) & mask1], 1u);
            }
    }
    __syncthreads();
    uint32_t c0 = h[0][threadIdx.x], c1 = h[1][threadIdx.x];
    if (c0) atomicAdd(&hist[threadIdx.x], c0);
    if (c1) atomicAdd(&hist[RADIX + threadIdx.x], c1);
}

__device__ __forceinline__ void range_boundary(int i, int m, const uint64_t* __restrict__ keys,
                                               int2* __restrict__ ranges) {
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

__global__ void __launch_bounds__(256)
tile_ranges_kernel(int m, const uint64_t* __restrict__ keys, int2* __restrict__ ranges) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) range_boundary(i, m, keys, ranges);
}

// 3 threads per instance, one float4 each: 48-byte records land contiguous and coalesced.
// keys != NULL also writes the tile ranges (fused frb_tile_ranges).
__global__ void __launch_bounds__(256)
gather_records_kernel(int m, const uint32_t* __restrict__ m_dev, const uint32_t* __restrict__ gids,
                      const float4* __restrict__ records,
                      float4* __restrict__ sorted_records, const float* __restrict__ phases,
                      float* __restrict__ sorted_phases, const uint64_t* __restrict__ keys,
                      int2* __restrict__ ranges) {
    frb_pdl_prologue();
    if (m_dev) m = min(m, (int)*m_dev);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 3ll * m;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t / 3), part = (int)(t - 3ll * i);
        uint32_t g = gids[i];
        sorted_records[t] = records[3 * (size_t)g + part];
        if (part == 0 && sorted_phases) sorted_phases[i] = phases[g];
        if (part == 1 && keys) range_boundary(i, m, keys, ranges);
    }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------------------------
// Cluster-resident depth sort (sm_100a: thread-block clusters + distributed shared memory).
//
// The one-sweep chain above spends a 100k-key pass on round trips, not data: ~1 us of load + rank inside ~7.5 us of
// look-back rounds through L2, scatter and kernel boundary (profiles/r2_y_sort_trace_100k.txt), times four passes, plus
// a memset and a histogram kernel.  A view's depth keys fit on chip: ONE cluster of 16 CTAs holds all (key, index)
// pairs in its shared memory (two buffers of 8 bytes per element, <= 12.4k elements per CTA), and all passes run inside
// ONE kernel.  CTA c owns positions [c S, (c + 1) S) of the array.  Per 8-bit pass:
//   1. rank: warp w walks its contiguous chunk in order, 32 elements per round (match.any + per-warp digit counters,
//      as in the one-sweep kernel), so (CTA, warp, round, lane) order is array order - the pass is stable;
//   2. thread d scans digit d over the 32 warps and stores the CTA's count of d into EVERY CTA's table (st.shared::cluster);
//   3. barrier.cluster; thread d adds up the column (all CTAs' counts of d), the counts of the CTAs before this one,
//      and the 256 column totals are scanned: base[d] = position of this CTA's first element with digit d;
//   4. local scatter into a staging buffer in (digit, array order) - plain shared-memory stores; then thread i sends
//      staged element i to position p = base[d] + (i - first staged index of d): CTA p / S, slot p % S of that CTA's
//      third buffer (st.shared::cluster.u64).  Consecutive threads hold consecutive positions of a digit run, so the
//      remote stores of a warp are a few contiguous spans.  (Scattering straight from the rank order - 32 unrelated
//      8-byte remote stores per warp instruction - was packet-bound: 14 us per pass, profiles/r2_z_cluster_sort_v1_*.)
//   5. barrier.cluster; rotate the three buffers.
// No global memory between the load of the depth bits and the store of order / rank, no look-back, no histogram
// kernel, no workspace.  Same result bit for bit (a stable LSD sort of the same keys).
// ------------------------------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 1024;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_MAX_CTAS = 16;
constexpr int CS_FIXED_SMEM = CS_MAX_CTAS * 256 * 4 + 2 * 256 * 4 + 2 * CS_WARPS * 4 + CS_WARPS * 256 * 2;
constexpr int CS_MAX_SMEM = 227 * 1024;
constexpr int CS_MAX_W = (CS_MAX_SMEM - CS_FIXED_SMEM) / 24 / 32;      // elements per warp: 256 -> 8192 per CTA

inline size_t cluster_sort_smem(int W) { return (size_t)24 * 32 * W + CS_FIXED_SMEM; }

#ifdef FRB_SORT_TRACE
__device__ unsigned long long frb_cluster_trace[CS_MAX_CTAS][6][10];     // [CTA][0 = load, 1.. = pass, 5 = store][stamp]
#define FRB_CTRACE(row, slot) do { if (threadIdx.x == 0) frb_cluster_trace[me][row][slot] = frb_gtime(); } while (0)
#else
#define FRB_CTRACE(row, slot) do { } while (0)
#endif

template <int IPT>      // rounds of 32 elements per warp: IPT >= ceil(W / 32)
__global__ void __launch_bounds__(CS_THREADS, 1)
cluster_sort_kernel(int n, const uint32_t* __restrict__ depth_bits, const __grid_constant__ KeyRange kr, int n_passes,
                    uint32_t last_mask, int W, uint32_t div_magic, uint32_t* __restrict__ order,
                    uint32_t* __restrict__ rank_out, uint32_t* __restrict__ error_word) {
    frb_pdl_prologue();
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned me = cluster.block_rank(), n_ctas = cluster.num_blocks();
    extern __shared__ __align__(16) unsigned char cs_smem[];
    const int S = 32 * W;
    uint64_t* buf0 = reinterpret_cast<uint64_t*>(cs_smem);
    uint64_t* buf1 = buf0 + S;
    uint64_t* buf2 = buf1 + S;
    uint32_t* all_tot = reinterpret_cast<uint32_t*>(buf2 + S);      // [CS_MAX_CTAS][256], row c written by CTA c
    uint32_t* base = all_tot + CS_MAX_CTAS * 256;                   // [256] array position of my first element of d
    uint32_t* lstart = base + 256;                                  // [256] staged index of my first element of d
    uint32_t* wsum = lstart + 256;                                  // [2][CS_WARPS]
    uint16_t* cnt = reinterpret_cast<uint16_t*>(wsum + 2 * CS_WARPS);   // [CS_WARPS][256]: a chunk is < 2^16 elements
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const long long g0 = (long long)me * S;

    FRB_CTRACE(0, 0);
    for (int j = threadIdx.x; j < S; j += CS_THREADS) {
        const long long g = g0 + j;
        if (g < n) buf0[j] = ((uint64_t)key_xform<uint32_t>(depth_bits[g], kr) << 32) | (uint32_t)g;
    }
    if (me == 0 && threadIdx.x == 0 && error_word) *error_word = 0;     // what the one-sweep path's memset leaves there
    FRB_CTRACE(0, 1);
    cluster.sync();         // every CTA of the cluster is running (its shared memory may be written) and loaded
    FRB_CTRACE(0, 2);

    uint64_t *cur = buf0, *stg = buf1, *nxt = buf2;
    const int my_count = (int)max(0ll, min((long long)S, (long long)n - g0));
    const int wbase = warp * W;
    const int w_valid = (int)max(0ll, min((long long)W, (long long)n - g0 - wbase));     // elements of my warp's chunk
    for (int p = 0; p < n_passes; ++p) {
        const int shift = 32 + 8 * p;
        const uint32_t mask = (p == n_passes - 1) ? last_mask : 255u;
        FRB_CTRACE(1 + p, 0);
        for (int i = threadIdx.x; i < CS_WARPS * 256 / 2; i += CS_THREADS) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
        __syncthreads();
        // 1. rank inside the warp's chunk
        uint32_t rnk[IPT];
        uint16_t* my_cnt = cnt + warp * 256;
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
            const int j = r * 32 + lane;
            const bool valid = j < w_valid;
            const uint32_t d = valid ? ((uint32_t)(cur[wbase + j] >> shift) & mask) : 256u;
            // lanes holding my digit, from nine ballots: MATCH.ANY iterates over the distinct values of the warp, and
            // with 32 different digits per round the ranking took 6.5 us of a pass (1.3 us in the 3-bit last pass)
            uint32_t peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 9; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t m = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? m : ~m;
            }
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = my_cnt[d];
                my_cnt[d] = (uint16_t)(old + __popc(peers));
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rnk[r] = old + __popc(peers & lt_mask);
            __syncwarp();
        }
        __syncthreads();
        FRB_CTRACE(1 + p, 1);       // ranked
        // 2. digit d over the warps; the CTA's count goes to every CTA's table
        uint32_t mine = 0;
        if (threadIdx.x < 256) {
            const int d = threadIdx.x;
#pragma unroll 8
            for (int w = 0; w < CS_WARPS; ++w) {
                const uint32_t c = cnt[w * 256 + d];
                cnt[w * 256 + d] = (uint16_t)mine;          // < S <= 12.4k
                mine += c;
            }
            for (unsigned c = 0; c < n_ctas; ++c) cluster.map_shared_rank(all_tot, c)[me * 256 + d] = mine;
        }
        FRB_CTRACE(1 + p, 2);       // counts published
        cluster.sync();     // 3. all counts are here
        FRB_CTRACE(1 + p, 3);
        uint32_t col = 0, before = 0, incl = 0, incl_l = 0;
        if (threadIdx.x < 256) {
            const int d = threadIdx.x;
            for (unsigned c = 0; c < n_ctas; ++c) {
                const uint32_t v = all_tot[c * 256 + d];
                col += v;
                if (c < me) before += v;
            }
            incl = col;
            incl_l = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                const uint32_t tl = __shfl_up_sync(0xffffffffu, incl_l, o);
                if (lane >= o) { incl += t; incl_l += tl; }
            }
            if (lane == 31) { wsum[warp] = incl; wsum[CS_WARPS + warp] = incl_l; }
        }
        __syncthreads();
        if (threadIdx.x < 256) {
            uint32_t woff = 0, woff_l = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (w < warp) { woff += wsum[w]; woff_l += wsum[CS_WARPS + w]; }
            base[threadIdx.x] = woff + incl - col + before;
            lstart[threadIdx.x] = woff_l + incl_l - mine;
        }
        __syncthreads();
        FRB_CTRACE(1 + p, 4);       // bases known
        // 4a. my elements into the staging buffer, ordered by (digit, array order)
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
            const int j = r * 32 + lane;
            if (j < w_valid) {
                const uint64_t e = cur[wbase + j];
                const uint32_t d = (uint32_t)(e >> shift) & mask;
                stg[lstart[d] + my_cnt[d] + rnk[r]] = e;
            }
        }
        __syncthreads();
        FRB_CTRACE(1 + p, 5);       // staged
        // 4b. staged element i -> its position in the array: the third buffer of the CTA that owns it
        for (int i = threadIdx.x; i < my_count; i += CS_THREADS) {
            const uint64_t e = stg[i];
            const uint32_t d = (uint32_t)(e >> shift) & mask;
            const uint32_t pos = base[d] + ((uint32_t)i - lstart[d]);
            uint32_t t = __umulhi(pos, div_magic);          // pos / S (div_magic = ceil(2^32 / S); pos * S < 2^32)
            uint32_t off = pos - t * (uint32_t)S;
            if (off >= (uint32_t)S) { off -= (uint32_t)S; ++t; }
            cluster.map_shared_rank(nxt, t)[off] = e;
        }
        FRB_CTRACE(1 + p, 6);       // sent
        cluster.sync();     // 5. every element is in place
        FRB_CTRACE(1 + p, 7);
        uint64_t* tmp = cur; cur = nxt; nxt = stg; stg = tmp;
    }
    for (int j = threadIdx.x; j < S; j += CS_THREADS) {
        const long long g = g0 + j;
        if (g < n) {
            const uint32_t id = (uint32_t)cur[j];
            order[g] = id;
            if (rank_out) rank_out[id] = (uint32_t)g;
        }
    }
    FRB_CTRACE(5, 0);
}

// 0: launched.  1: not applicable here (size, device, FRB_CLUSTER_SORT=0) - the caller runs the one-sweep chain.
template <int IPT>
int cluster_sort_launch_ipt(int n_ctas, int W, int n, const uint32_t* depth_bits, KeyRange kr, int nbits, uint32_t* order,
                            uint32_t* rank, uint32_t* error_word, cudaStream_t st) {
    const size_t smem = cluster_sort_smem(W);
    int dev = 0;
    FRB_CUDA_OK(cudaGetDevice(&dev));
    // per device and instantiation: opt in to the cluster size and the shared memory once, and ask whether such a
    // cluster can be resident at all
    static int usable[64][2] = {};          // [device][16 / 8 CTAs]: 0 unknown, 1 yes, -1 no
    static size_t opted_smem[64] = {};
    const int slot = n_ctas == 16 ? 0 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas);
    cfg.blockDim = dim3(CS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = n_ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    if (dev < 64 && (usable[dev][slot] == 0 || opted_smem[dev] < smem)) {
        cudaError_t e = cudaFuncSetAttribute(cluster_sort_kernel<IPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(cluster_sort_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_MAX_SMEM);
        int clusters = 0;
        cfg.numAttrs = 1;
        cfg.dynamicSmemBytes = CS_MAX_SMEM;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&clusters, cluster_sort_kernel<IPT>, &cfg);
        cfg.dynamicSmemBytes = smem;
        if (e != cudaSuccess) (void)cudaGetLastError();
        usable[dev][slot] = (e == cudaSuccess && clusters >= 1) ? 1 : -1;
        opted_smem[dev] = CS_MAX_SMEM;
    }
    if (dev >= 64 || usable[dev][slot] != 1) return 1;
    cfg.numAttrs = frb_pdl_enabled() ? 2 : 1;
    const int n_passes = (nbits + 7) / 8;
    const uint32_t last_mask = (1u << (nbits - 8 * (n_passes - 1))) - 1u;
    const uint32_t S = 32u * (uint32_t)W;
    const uint32_t div_magic = (uint32_t)((0x100000000ull + S - 1) / S);
    FRB_CUDA_OK(cudaLaunchKernelEx(&cfg, cluster_sort_kernel<IPT>, n, depth_bits, kr, n_passes, last_mask, W, div_magic,
                                   order, rank, error_word));
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// Which of the two sorts frb_depth_order* runs for n <= 131,072 (frb_depth_sort_in_cluster): the cluster keeps 16 SMs
// busy for ~56 us where the chain keeps all of them waiting on round trips for ~42 us.  One frame at a time the chain is
// faster (2614 against 2510 frames/s); with several frames in flight on other streams the cluster leaves 132 SMs to
// their compositor kernels (host-to-host 2893 against 2845 frames/s, profiles/r3_d_cluster_sort_*).  Default: the chain;
// fresnel_b200.host.HostRenderPipeline captures its graphs with the cluster sort.
int g_cluster_sort_mode = -1;       // -1: environment FRB_CLUSTER_SORT (unset = off), 0: off, 1: on

int cluster_sort_try(int n, const uint32_t* depth_bits, KeyRange kr, int nbits, uint32_t* order, uint32_t* rank,
                     uint32_t* error_word, cudaStream_t st) {
    int mode = __atomic_load_n(&g_cluster_sort_mode, __ATOMIC_RELAXED);
    if (mode < 0) {
        const char* sw = getenv("FRB_CLUSTER_SORT");    // read per call: the tests compare both paths in one process
        mode = (sw && sw[0] == '1') ? 1 : 0;
    }
    if (!mode || nbits < 1 || nbits > 32) return 1;
    for (int n_ctas = CS_MAX_CTAS; n_ctas >= 8; n_ctas /= 2) {
        const int per_cta = frb_div_up(n, n_ctas);
        const int W = max(1, frb_div_up(per_cta, 32));
        if (W > CS_MAX_W) return 1;
        const int rounds = frb_div_up(W, 32);
        int rc;
        if (rounds <= 1) rc = cluster_sort_launch_ipt<1>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        else if (rounds <= 2) rc = cluster_sort_launch_ipt<2>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        else if (rounds <= 4) rc = cluster_sort_launch_ipt<4>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        else if (rounds <= 7) rc = cluster_sort_launch_ipt<7>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        else if (rounds <= 10) rc = cluster_sort_launch_ipt<10>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        else rc = cluster_sort_launch_ipt<13>(n_ctas, W, n, depth_bits, kr, nbits, order, rank, error_word, st);
        if (rc != 1) return rc;
    }
    return 1;
}

}  // namespace

extern "C" size_t frb_sort_workspace_bytes(int m) {
    if (m < 0) m = 0;
    return sizeof(uint32_t) * sort_ws_words(m, MAX_PASSES);
}

static int sort_pairs(int m, const uint32_t* m_dev, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                      uint32_t* vals_tmp, int begin_bit, int end_bit, void* workspace, void* stream) {
    if (m < 0 || begin_bit < 0 || end_bit > 64) return FRB_E_INVALID;
    if (m == 0 || end_bit <= begin_bit) return 0;
    if (!keys || !vals || !keys_tmp || !vals_tmp || !workspace) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    bool in_b = false;
    int rc = radix_sort_impl<uint64_t, RADIX_BITS>(m, m_dev, keys, vals, keys, vals, keys_tmp, vals_tmp, begin_bit, end_bit,
                                       (uint32_t*)workspace, st, &in_b);
    if (rc) return rc;
    if (in_b) {
        FRB_CUDA_OK(cudaMemcpyAsync(keys, keys_tmp, sizeof(uint64_t) * (size_t)m, cudaMemcpyDeviceToDevice, st));
        FRB_CUDA_OK(cudaMemcpyAsync(vals, vals_tmp, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

extern "C" int frb_radix_sort_pairs(int m, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                                    uint32_t* vals_tmp, int begin_bit, int end_bit, void* workspace,
                                    void* stream) {
    return sort_pairs(m, nullptr, keys, vals, keys_tmp, vals_tmp, begin_bit, end_bit, workspace, stream);
}

extern "C" int frb_radix_sort_pairs_dev(int m_capacity, const uint32_t* m_dev, uint64_t* keys, uint32_t* vals,
                                        uint64_t* keys_tmp, uint32_t* vals_tmp, int begin_bit, int end_bit,
                                        void* workspace, void* stream) {
    if (!m_dev) return FRB_E_INVALID;
    return sort_pairs(m_capacity, m_dev, keys, vals, keys_tmp, vals_tmp, begin_bit, end_bit, workspace, stream);
}

extern "C" int frb_depth_sort_in_cluster(int mode) {
    if (mode < -1 || mode > 1) return FRB_E_INVALID;
    return __atomic_exchange_n(&g_cluster_sort_mode, mode, __ATOMIC_RELAXED);
}

extern "C" size_t frb_depth_order_workspace_bytes(int n) {
    if (n < 0) n = 0;
    return 3 * align256(sizeof(uint32_t) * (size_t)n) + align256(sizeof(uint32_t) * sort_ws_words(n, 4, RADIX_MAX));
}

extern "C" int frb_depth_order(int n, const uint32_t* depth_bits, uint32_t* order, void* workspace,
                               void* stream) {
    return frb_depth_order_rank(n, depth_bits, order, nullptr, workspace, stream);
}

extern "C" int frb_depth_order_rank(int n, const uint32_t* depth_bits, uint32_t* order, uint32_t* rank,
                                    void* workspace, void* stream) {
    if (n < 0) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!depth_bits || !order || !workspace) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    size_t a = align256(sizeof(uint32_t) * (size_t)n);
    uint32_t* keys_a = (uint32_t*)w;
    uint32_t* keys_b = (uint32_t*)(w + a);
    uint32_t* vals_a = (uint32_t*)(w + 2 * a);
    uint32_t* ws = (uint32_t*)(w + 3 * a);
    {
        const int rc = cluster_sort_try(n, depth_bits, KeyRange{0u, 0u, 0, 0}, 32, order, rank, ws + WS_ERROR, st);
        if (rc != 1) return rc;
    }
    // 4 passes: depth_bits -> A -> B -> A -> B ; the value buffer B is `order` itself
    bool in_b = false;
    int rc = radix_sort_impl<uint32_t, RADIX_BITS>(n, nullptr, depth_bits, nullptr, keys_a, vals_a, keys_b, order, 0, 32, ws, st, &in_b,
                                       false, rank);
    if (rc) return rc;
    if (!in_b) return FRB_E_INVALID;   // cannot happen with an even number of passes
    return 0;
}

// The depth order of what a camera can see: keys = depth bits clamped to [near_bits, far_bits] minus near_bits
// (key_xform), sorted on the bits that range needs (27 for the default [0.01, 100]: four 8-bit passes as before; three
// when far < 2 near, one or two for thin slabs); ranges of 2^27 values and more fall back to the full 32-bit sort.
// The order of the VISIBLE Gaussians (near < depth < far) is exactly frb_depth_order's; culled ones end up at the two
// ends in index order.
extern "C" int frb_depth_order_range(int n, const uint32_t* depth_bits, float near_depth, float far_depth,
                                     uint32_t* order, uint32_t* rank, void* workspace, void* stream) {
    if (n < 0) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!depth_bits || !order || !workspace) return FRB_E_INVALID;
    uint32_t lo, hi;
    memcpy(&lo, &near_depth, 4);
    memcpy(&hi, &far_depth, 4);
    const bool usable = near_depth > 0.0f && far_depth > near_depth && far_depth < 3.0e38f &&
                        (hi - lo) < (1u << (3 * RADIX_BITS_MAX));
    if (!usable) return frb_depth_order_rank(n, depth_bits, order, rank, workspace, stream);
    int nbits = 1;
    while ((1u << nbits) <= (hi - lo)) ++nbits;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    size_t a = align256(sizeof(uint32_t) * (size_t)n);
    uint32_t* keys_a = (uint32_t*)w;
    uint32_t* keys_b = (uint32_t*)(w + a);
    uint32_t* vals_a = (uint32_t*)(w + 2 * a);
    uint32_t* ws = (uint32_t*)(w + 3 * a);
    KeyRange kr{lo, hi, 1, 0};
    bool in_b = false;
    int rc = cluster_sort_try(n, depth_bits, kr, nbits, order, rank, ws + WS_ERROR, st);
    if (rc != 1) return rc;
    uint32_t *va = vals_a, *vb = order;
    // Digit width: 8 bits.  Three 9-bit passes for the default [0.01, 100] (27 key bits) were measured against four
    // 8-bit ones: 45-48 us against 41 us at 100k keys - a 512-digit pass costs more (twice the counters to clear and
    // scan per warp, two look-back words per thread) than the pass it saves; the 9-bit instantiation is kept out of
    // the build.  The range still pays whenever it fits 24 bits (near / far within a factor of two: three passes).
    // (Measured again with the ballot form of the digit match, which does not mind 512 distinct digits: 43.5-44.9 us
    // against 43.2 us for four 8-bit passes on one box - still no gain.)
    const int passes8 = (nbits + 7) / 8;
    if (passes8 % 2 == 1) { va = order; vb = vals_a; }
    rc = radix_sort_impl<uint32_t, RADIX_BITS>(n, nullptr, depth_bits, nullptr, keys_a, va, keys_b, vb, 0, nbits, ws, st,
                                               &in_b, false, rank, kr);
    if (rc) return rc;
    if (in_b != (passes8 % 2 == 0)) return FRB_E_INVALID;
    return 0;
}

// Device address of the word the radix passes of frb_depth_order* set when a look-back gave up (SPIN_LIMIT): a
// consumer that synchronises anyway (frb_tile_scan's instance count) carries it to the host.
extern "C" const uint32_t* frb_depth_order_error_word(int n, const void* workspace) {
    if (n <= 0 || !workspace) return nullptr;
    const size_t a = align256(sizeof(uint32_t) * (size_t)n);
    return (const uint32_t*)((const char*)workspace + 3 * a) + WS_ERROR;
}

extern "C" size_t frb_scan_workspace_bytes(int n) {
    if (n < 0) n = 0;
    return sizeof(unsigned long long) * (size_t)(frb_div_up(n, EMIT_TILE) + 2);   // covers both scan kernels
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
    FRB_CUDA_OK(cudaMemsetAsync(workspace, 0, frb_scan_workspace_bytes(n), st));
    frb_launch(offsets_scan_kernel, dim3(nb), dim3(SCAN_THREADS), 0, st, n, tiles_touched, order, offsets,
                                                     (unsigned long long*)workspace);
    frb_note_launches(1);
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
    frb_launch(bin_emit_kernel, dim3(frb_div_up(n, 256)), dim3(256), 0, (cudaStream_t)stream, 
        n, n / n_views, tiles_x, tiles_x * tiles_y, (const float4*)records, depth_bits, order, offsets, keys,
        gids);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// Whole-pass binning (pipeline.cu): scan + emit + sort histograms in one kernel, then the one-sweep passes on the tile
// bits.  Same result as frb_tile_offsets + frb_bin_emit + frb_radix_sort_pairs_dev; sorted pairs end in keys / gids.
// m_out: device word that receives the true instance count (also read by the passes); scan_ws:
// frb_scan_workspace_bytes(n); sort_ws: frb_sort_workspace_bytes(m_capacity).
extern "C" int frb_bin_sort_dev(int n, int n_views, int width, int height, const float* records,
                                const uint32_t* depth_bits, const uint32_t* tiles_touched, const uint32_t* order,
                                int m_capacity, uint32_t* m_out, uint64_t* keys, uint32_t* gids, uint64_t* keys_tmp,
                                uint32_t* vals_tmp, int tile_bits, void* scan_ws, void* sort_ws, void* stream) {
    if (n <= 0 || n_views < 1 || n_views > FRB_MAX_VIEWS || n % n_views != 0 || m_capacity <= 0) return FRB_E_INVALID;
    if (width < 1 || height < 1 || tile_bits < 1 || tile_bits > 2 * RADIX_BITS) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    if (!records || !depth_bits || !tiles_touched || !m_out || !keys || !gids || !keys_tmp || !vals_tmp ||
        !scan_ws || !sort_ws)
        return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_passes = tile_bits > RADIX_BITS ? 2 : 1;
    const uint32_t mask1 = tile_bits > RADIX_BITS ? (1u << (tile_bits - RADIX_BITS)) - 1u : 0u;
    FRB_CUDA_OK(cudaMemsetAsync(scan_ws, 0, frb_scan_workspace_bytes(n), st));
    FRB_CUDA_OK(cudaMemsetAsync(sort_ws, 0, sizeof(uint32_t) * sort_ws_words(m_capacity, n_passes), st));
    int tiles_x = frb_div_up(width, FRB_TILE), tiles_y = frb_div_up(height, FRB_TILE);
    frb_launch(scan_emit_kernel, dim3(frb_div_up(n, EMIT_TILE)), dim3(EMIT_TILE), 0, st, n, n / n_views, tiles_x,
               tiles_x * tiles_y, (const float4*)records, depth_bits, tiles_touched, order, keys, gids,
               (unsigned long long*)scan_ws, (uint32_t*)sort_ws + WS_HIST, mask1, m_out);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    bool in_b = false;
    int rc = radix_sort_impl<uint64_t, RADIX_BITS>(m_capacity, m_out, keys, gids, keys, gids, keys_tmp, vals_tmp, 32, 32 + tile_bits,
                                       (uint32_t*)sort_ws, st, &in_b, /*hist_ready=*/true);
    if (rc) return rc;
    if (in_b) {
        FRB_CUDA_OK(cudaMemcpyAsync(keys, keys_tmp, sizeof(uint64_t) * (size_t)m_capacity, cudaMemcpyDeviceToDevice, st));
        FRB_CUDA_OK(cudaMemcpyAsync(gids, vals_tmp, sizeof(uint32_t) * (size_t)m_capacity, cudaMemcpyDeviceToDevice, st));
    }
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
    frb_launch(gather_records_kernel, dim3(min(frb_div_up(3ll * m, 256), 148 * 32)), dim3(256), 0, (cudaStream_t)stream, 
        m, nullptr, gids, (const float4*)records, (float4*)sorted_records, phases, sorted_phases, nullptr, nullptr);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

static int ranges_and_gather(int m, const uint32_t* m_dev, const uint64_t* keys, const uint32_t* gids,
                             int n_tiles, int32_t* ranges, const float* records, float* sorted_records,
                             const float* phases, float* sorted_phases, void* stream) {
    if (m < 0 || n_tiles < 0 || !ranges) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    FRB_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)n_tiles, st));
    if (m == 0) return 0;
    if (!keys || !gids || !records || !sorted_records) return FRB_E_INVALID;
    if ((phases == nullptr) != (sorted_phases == nullptr)) return FRB_E_INVALID;
    frb_launch(gather_records_kernel, dim3(min(frb_div_up(3ll * m, 256), 148 * 32)), dim3(256), 0, st, 
        m, m_dev, gids, (const float4*)records, (float4*)sorted_records, phases, sorted_phases, keys,
        (int2*)ranges);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_ranges_and_gather(int m, const uint64_t* keys, const uint32_t* gids, int n_tiles,
                                     int32_t* ranges, const float* records, float* sorted_records,
                                     const float* phases, float* sorted_phases, void* stream) {
    return ranges_and_gather(m, nullptr, keys, gids, n_tiles, ranges, records, sorted_records, phases,
                             sorted_phases, stream);
}

extern "C" int frb_ranges_and_gather_dev(int m_capacity, const uint32_t* m_dev, const uint64_t* keys,
                                         const uint32_t* gids, int n_tiles, int32_t* ranges, const float* records,
                                         float* sorted_records, const float* phases, float* sorted_phases,
                                         void* stream) {
    if (!m_dev) return FRB_E_INVALID;
    return ranges_and_gather(m_capacity, m_dev, keys, gids, n_tiles, ranges, records, sorted_records, phases,
                             sorted_phases, stream);
}
