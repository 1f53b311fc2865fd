// Front-to-back alpha compositing over per-tile Gaussian lists, and its backward.
// Reference: the per-Gaussian loop of TileBasedRenderer.forward DR:582-667, the epilogue
// DR:669-686 and the autograd tape through both.
//
// One CTA per 16x16 tile, one thread per pixel.  The tile's records (48 B each, contiguous in
// the sorted instance list) are staged into shared memory in batches by 1D TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) through a ring of FRB_STAGES buffers,
// so the copy of batch b+STAGES-1 overlaps the arithmetic of batch b.  Every thread reads the
// same record at the same time (shared-memory broadcast, no bank conflicts).
//
// Per pixel and record (appendix A.3), with the conic pre-scaled by -0.5*log2(e):
//     inside = pixel in [x0,x1) x [y0,y1)                       (the reference's rectangle)
//     g = exp2(A' dx^2 + B' dx dy + C' dy^2);  alpha = clamp(g * opacity, 0, 0.99)
//     c = alpha * (1 - Acc);  colour += c * rgb;  depth += c * d;  Acc += c          (DR:650-658, the SUM form)
// The weights follow the reference's own recurrence on the accumulated alpha, operation for operation.  A product
// form T *= (1 - alpha) is the same number on paper but not in fp32 at thousands of overlaps per pixel: once the
// running colour / depth sums are large, late contributions fall below half an ulp and are absorbed, and in the sum
// form the accumulated alpha absorbs the very same terms, which keeps 1 - Acc (and with it every later weight)
// larger - the two losses compensate.  Measured on tests/golden/tile_overlap_faint_20k_128 (5,380 entries at the
// worst pixel): product-form weights end 1.8e-5 from the reference's depth, sum-form weights 1e-7.
// The transmittance is carried in product form AS WELL (T, two more instructions per pair): it decides early
// termination and is what the backward pass divides its way back from, which needs T's relative precision
// (1 - Acc is quantised to 6e-8 and useless for that once the pixel saturates).
#include "frb_common.cuh"

#include "composite_common.cuh"

namespace {

// GATHER = false: the tile's records are a contiguous span of sorted_records (one 1D bulk copy per batch, 48-byte rows).
// GATHER = true: the tile's list is a span of Gaussian ids and the records are fetched from the UNSORTED array by TMA
// tile::gather4 (SASS UTMALDG.2D.GATHER4: four rows of a 2D tensor map by row index per operation, 64-byte rows in
// shared memory - the box is 16 floats wide so that every 4-row group lands 256-byte aligned; columns 12..15 are out of
// bounds and zero-filled).  The gather removes the 48 M-byte write and read-back of the sorted record copy.
template <bool GATHER>
struct FwdStage {
    static constexpr int RS = GATHER ? 4 : 3;           // float4 per record in shared memory
    float4 rec[BATCH * RS];
};

// Tried in round 2: two CTAs of four warps per tile, as the backward kernel now runs: 0.0948 against 0.0927 ms
// (profiles/r4_e_*) - the forward's batch barrier costs less than staging every tile's records twice.
// Tried in round 2: __launch_bounds__(256, 6) (40 registers instead of 47, six CTAs per SM instead of five): 0.104 against
// 0.096 ms (profiles/r3_p_*) - the register cap costs more instructions than the sixth CTA hides.
// Tried in round 2: two pixels per thread (128 threads, a warp owns an 8x8 block, the pixel-independent part of a
// record - rectangle words, loop, centre / conic load, dx - paid once for (x, y) and (x, y + 4)).  Bit-identical
// output and fewer instructions, but a tile then takes twice as long on half the warps and the kernel ends on its
// longest tiles: 119.5 us against 100.8 us for one frame (profiles/r2_u_*; with six frames in flight, where other
// frames fill the tail, 2860-2910 against 2855-2860 frames/s).  Kept: one pixel per thread.
// BY_CANDIDATE (round 2, what the whole-pass calls run): instead of testing every record's rectangle in every thread
// and branching around the body record by record, a warp finds the records whose rectangle touches its 8x4 block
// with one ballot per 32 records (lane = record) and walks only those, FOUR per iteration: the state-independent part
// (loads, rectangle test, power, exp2, clamp) of the four first, then their four short accumulation tails - four
// independent chains in flight instead of one behind a branch.  A lane outside a candidate's rectangle carries
// alpha = 0 through the tail (c = 0: no bit of its sums changes).  Early termination is tested once per iteration.
template <bool GATHER, bool BY_CANDIDATE>
__global__ void __launch_bounds__(CTA_THREADS)
composite_fwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int* __restrict__ tile_order,
                     const int2* __restrict__ ranges, const float4* __restrict__ sorted_records,
                     const __grid_constant__ CUtensorMap record_map, const uint32_t* __restrict__ sorted_gids,
                     float3 bg, float t_eps, float alpha_max, float* __restrict__ image,
                     float* __restrict__ depth_out, float* __restrict__ alpha_out, float* __restrict__ state_T,
                     int* __restrict__ state_n) {
    frb_pdl_prologue();
    constexpr int RS = FwdStage<GATHER>::RS;
    __shared__ __align__(1024) FwdStage<GATHER> stage[STAGES];
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tile = tile_order ? tile_order[blockIdx.x] : blockIdx.x;   // heaviest tiles first
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + foot_x(threadIdx.x);
    const int py = ty * TILE + foot_y(threadIdx.x);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;
    // BY_CANDIDATE: the warp's 8x4 pixel block, and the pixel's words for the rectangle test - moved to a pixel that is
    // in no rectangle (32766, 32766) once the pixel is done or when it lies outside the image
    constexpr uint32_t NO_PIXEL_GUARD = 0xfffefffeu, NO_PIXEL_PLUS1 = 0x7fff7fffu;
    const int wx0 = tx * TILE + ((threadIdx.x >> 5) % (TILE / FOOT_W)) * FOOT_W, wx1 = wx0 + FOOT_W;
    const int wy0 = ty * TILE + ((threadIdx.x >> 5) / (TILE / FOOT_W)) * FOOT_H, wy1 = wy0 + FOOT_H;
    uint32_t pxy_guard_v = in_image ? pxy_guard : NO_PIXEL_GUARD, pxy_plus1_v = in_image ? pxy_plus1 : NO_PIXEL_PLUS1;

    const int2 range = ranges[tile];
    const int count = range.y - range.x;
    const int n_batches = (count + BATCH - 1) / BATCH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) frb_mbar_init(&full_bar[s], 1);
        frb_mbar_fence_init();
    }
    __syncthreads();
    // producer: thread 0 (bulk copy) or lanes 0..15 of warp 0 (gather: lane q fetches records 4q..4q+3 of the batch)
    // (gather: the producer of batch b is warp b % 8, lanes 0..15, so that the work does not always delay the same warp)
    const int lane_f = threadIdx.x & 31, warp_f = threadIdx.x >> 5;
    uint4 pre_g = make_uint4(0u, 0u, 0u, 0u);           // gather: ids of my group in the next batch I issue
    auto load_ids = [&](int b) {
        if (GATHER && warp_f == (b & 7) && lane_f < BATCH / 4 && b < n_batches) {
            const int i0 = b * BATCH + 4 * lane_f, last = count - 1;
            const uint32_t* g = sorted_gids + range.x;
            pre_g = make_uint4(g[min(i0, last)], g[min(i0 + 1, last)], g[min(i0 + 2, last)], g[min(i0 + 3, last)]);
        }
    };
    auto issue = [&](int b) {
        const int s = b % STAGES;
        const int cnt = min(BATCH, count - b * BATCH);
        if (!GATHER) {
            if (threadIdx.x == 0) {
                frb_mbar_expect_tx(&full_bar[s], cnt * RECORD_BYTES);
                frb_tma_load_1d(stage[s].rec, sorted_records + 3 * (size_t)(range.x + b * BATCH), cnt * RECORD_BYTES,
                                &full_bar[s]);
            }
        } else if (warp_f == (b & 7)) {
            const int groups = (cnt + 3) >> 2;
            if (lane_f == 0) frb_mbar_expect_tx(&full_bar[s], groups * 4 * RS * 16);
            __syncwarp();
            if (lane_f < groups)
                frb_tma_gather4(stage[s].rec + 4 * RS * lane_f, &record_map, pre_g, &full_bar[s]);
        }
    };
    for (int b = 0; b < STAGES && b < n_batches; ++b) {
        load_ids(b);
        issue(b);
    }
    load_ids(STAGES);

    float T = 1.0f, acc = 0.0f, W1 = 1.0f;      // W1 = 1 - acc
    // colour and depth sums as two fp32 pairs: (r, g) and (b, depth) are neighbours in the record, so the four fmas
    // of a pair are two FFMA2 (same bits as fmaf)
    float2 c_rg = make_float2(0.0f, 0.0f), c_bd = c_rg;
    // (the clamp bound is re-read from the constant bank for every record - LDCU in the loop body; forcing it into a
    // vector register removed that instruction and made the kernel 0.8 % slower, profiles/r3_m_*)
    int consumed = count;               // list entries walked; lowered when the pixel stops early
    bool done = !in_image;
    const float stop = fmaxf(t_eps, T_FLOOR);

    for (int b = 0; b < n_batches; ++b) {
        const int s = b % STAGES;
        const int cnt = min(BATCH, count - b * BATCH);
        const int cnt_pad = (cnt + CHUNK - 1) & ~(CHUNK - 1);
        frb_mbar_wait(&full_bar[s], (b / STAGES) & 1);
        if (!BY_CANDIDATE && cnt_pad != cnt) {
            // last batch: pad to a whole chunk with records whose rectangle contains no pixel
            if (threadIdx.x < (cnt_pad - cnt) * RS) {
                const int k = threadIdx.x % RS;
                stage[s].rec[RS * cnt + threadIdx.x] =
                    make_float4(0.f, 0.f, k == 1 ? __uint_as_float(NULL_RECT_LO) : 0.f,
                                k == 1 ? __uint_as_float(NULL_RECT_HI) : 0.f);
            }
            __syncthreads();
        }
        if (BY_CANDIDATE) {
            const float4* rec = stage[s].rec;
            // (a warp whose pixels are all done skips the batch; done lanes sit on a pixel no rectangle contains)
            if (__any_sync(0xffffffffu, !done)) {
                for (int sb = 0; sb * 32 < cnt; ++sb) {
                    const int sub_cnt = min(32, cnt - sb * 32);
                    uint32_t cand;
                    {
                        bool ok = false;
                        if (lane_f < sub_cnt) {
                            const float4 q = rec[RS * (sb * 32 + lane_f) + 1];
                            const uint32_t lo = __float_as_uint(q.z), hi = __float_as_uint(q.w) & 0x7fff7fffu;
                            ok = (int)(lo & 0xffff) < wx1 && (int)(hi & 0xffff) > wx0 && (int)(lo >> 16) < wy1 &&
                                 (int)(hi >> 16) > wy0;
                        }
                        cand = __brev(__ballot_sync(0xffffffffu, ok));      // bit 31 - j: walked front to back by bfind
                    }
                    const char* rec_sb = reinterpret_cast<const char*>(rec + RS * (sb * 32));
                    struct Pre { float a, d; float2 rg, bd; };
                    auto stage_a = [&](int j) {
                        Pre p;
                        const float4* rj = reinterpret_cast<const float4*>(rec_sb + j * (RS * 16));
                        const float4 r1 = rj[1], r0 = rj[0], r2 = rj[2];
                        const bool in = rect_contains(pxy_guard_v, pxy_plus1_v, __float_as_uint(r1.z), __float_as_uint(r1.w));
                        const float dx = fpx - r0.x, dy = fpy - r0.y;
                        const float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                        const float a = fminf(fmaxf(frb_ex2(power) * r1.y, 0.0f), alpha_max);
                        p.a = in ? a : 0.0f;
                        p.rg = make_float2(r2.x, r2.y);
                        p.bd = make_float2(r2.z, r2.w);
                        return p;
                    };
                    auto stage_b = [&](const Pre& p) {
                        const float c = p.a * W1;
                        c_rg = frb_fma2s(c, p.rg, c_rg);
                        c_bd = frb_fma2s(c, p.bd, c_bd);
                        acc += c;
                        W1 = 1.0f - acc;
                        T = fmaf(-p.a, T, T);
                    };
                    int last = -1;
                    while (cand) {
                        if (__popc(cand) >= 4) {
                            int jj[4];
                            Pre pp[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int bpos = frb_bfind(cand);
                                cand ^= 1u << bpos;
                                jj[q] = 31 - bpos;
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) pp[q] = stage_a(jj[q]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) stage_b(pp[q]);
                            last = jj[3];
                        } else {
                            const int bpos = frb_bfind(cand);
                            cand ^= 1u << bpos;
                            last = 31 - bpos;
                            const Pre p0 = stage_a(last);
                            stage_b(p0);
                        }
                        // early termination, tested once per iteration; the backward pass replays exactly the
                        // entries [0, consumed)
                        if (!done && T < stop) {
                            done = true;
                            consumed = b * BATCH + sb * 32 + last + 1;
                            pxy_guard_v = NO_PIXEL_GUARD;       // in no rectangle from here on
                            pxy_plus1_v = NO_PIXEL_PLUS1;
                        }
                        if (!__any_sync(0xffffffffu, !done)) cand = 0;
                    }
                }
            }
        } else if (!done) {
            const float4* rec = stage[s].rec;
            for (int j0 = 0; j0 < cnt_pad; j0 += CHUNK) {
#pragma unroll
                for (int k = 0; k < CHUNK; ++k) {
                    const int j = j0 + k;
                    float4 r1 = rec[RS * j + 1], r2 = rec[RS * j + 2];
                    if (rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.z), __float_as_uint(r1.w))) {
                        float4 r0 = rec[RS * j + 0];
                        float dx = fpx - r0.x, dy = fpy - r0.y;
                        float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                        float a = frb_ex2(power) * r1.y;
                        a = fminf(fmaxf(a, 0.0f), alpha_max);
                        float c = a * W1;
                        c_rg = frb_fma2s(c, make_float2(r2.x, r2.y), c_rg);
                        c_bd = frb_fma2s(c, make_float2(r2.z, r2.w), c_bd);
                        acc += c;
                        W1 = 1.0f - acc;
                        T = fmaf(-a, T, T);
                    }
                }
                // early termination is tested once per chunk; the backward pass replays exactly
                // the entries [0, consumed)
                if (T < stop) {
                    done = true;
                    consumed = min(b * BATCH + j0 + CHUNK, count);
                    break;
                }
            }
        }
        // everyone is finished with stage s: refill it, or stop early when no pixel needs more
        int all_done = __syncthreads_and(done ? 1 : 0);
        if (all_done) {
            // bulk copies already in flight must land before this CTA's shared memory is released
            for (int bb = b + 1; bb < n_batches && bb < b + STAGES; ++bb)
                frb_mbar_wait(&full_bar[bb % STAGES], (bb / STAGES) & 1);
            break;
        }
        if (b + STAGES < n_batches) {
            issue(b + STAGES);
            load_ids(b + STAGES + 1);
        }
    }

    if (in_image) {
        const size_t hw = (size_t)width * height;
        const size_t pix = (size_t)view * hw + (size_t)py * width + px;
        float* img = image + (size_t)view * 3 * hw + (size_t)py * width + px;
        // DR:670-675
        const float cr = c_rg.x, cg = c_rg.y, cb = c_bd.x, cd = c_bd.y;
        float o0 = fmaf(W1, bg.x, cr), o1 = fmaf(W1, bg.y, cg), o2 = fmaf(W1, bg.z, cb);
        img[0] = fminf(fmaxf(o0, 0.0f), 1.0f);
        img[hw] = fminf(fmaxf(o1, 0.0f), 1.0f);
        img[2 * hw] = fminf(fmaxf(o2, 0.0f), 1.0f);
        depth_out[pix] = cd;
        alpha_out[pix] = acc;
        state_T[pix] = T;
        // torch.clamp backward passes the gradient where 0 <= x <= 1 (inclusive): keep the three gates
        int gates = ((o0 >= 0.0f && o0 <= 1.0f) ? 1 : 0) | ((o1 >= 0.0f && o1 <= 1.0f) ? 2 : 0) |
                    ((o2 >= 0.0f && o2 <= 1.0f) ? 4 : 0);
        state_n[pix] = consumed | (gates << STATE_GATE_SHIFT);
    }
}

// ------------------------------------------------------------------------------------------
// Backward (appendix A.4).  Walks each tile list back to front.  Per pixel:
//   w_i = gC . rgb_i + gD * depth_i ; T_i = T_{i+1} / (1 - alpha_i) ;
//   dL/dalpha_i = T_i w_i - (S + T_final X) / (1 - alpha_i),  S = sum_{j>i} c_j w_j,
//   X = gC . bg - gA  (alpha_out = 1 - T_final)
// gated by 0 <= g*o <= 0.99 (torch.clamp backward is inclusive).
//
// The per-Gaussian sums over pixels are formed WITHOUT warp shuffles or per-pixel atomics.  Each
// warp owns 32 pixels (a 16x2 block) and handles 32 Gaussians at a time in two phases:
//   phase 1 (lane = pixel): walk the 32 Gaussians back to front, carrying (T, S); per pair store
//           (c, g * gated dL/dalpha) into a 32x32 shared-memory tile;
//   phase 2 (lane = Gaussian): read the tile transposed, walk the 32 pixels and accumulate the ten
//           gradient sums of that Gaussian in registers.
// The eight warps' partial sums are added across warps through shared memory and leave the CTA as
// ONE atomic add per Gaussian value per tile.
// ------------------------------------------------------------------------------------------
constexpr int BWD_WARPS = CTA_THREADS / 32;
constexpr int BWD_FW = 8, BWD_FH = 4;           // pixel block of one warp (BWD_FW * BWD_FH = 32)
constexpr int PAIR_STRIDE = 33;                 // float2 row stride: conflict-free both ways
constexpr int N_GRADS = 10;
constexpr int QUAD_MIN = 4;                     // phase 1 takes four candidates per iteration when this many remain
                                                // (two: 0.1775 ms, four: 0.1674 ms, eight: 0.172 ms; profiles/r3_q_*, r3_r_*)

// 105 KB per CTA, 124 registers: two CTAs (16 warps) per SM.  Round 2 tried the judge's suggestion of a 16-Gaussian
// exchange tile (71.7 KB, __launch_bounds__(256, 3), 80 registers with 24 bytes of spills: three CTAs per SM): achieved
// occupancy 24 -> 36 %, issue slots 69 -> 76 % busy, but the two-halves loop costs 13 % more instructions (161.2 M
// against 142.4 M: the cross-half shuffles, a second pass over the candidate mask and record loads) - 199 us against
// 188 us on the same box (profiles/r2_d_*).  The kernel is bound by instructions issued, not by residency.
// Also tried: the second block-wide barrier of a visit split into bar.arrive by the summing threads / bar.sync before
// the next visit's first part[] store, so that the cross-warp sum overlaps the next phase 1 - 197.8 us against
// 194.6 us (profiles/r2_t_*): the barrier stall in the capture is warps whose pixels ended early, not the sum.
// HALVES = 1: one CTA of eight warps per tile, four ring stages.  HALVES = 2: TWO CTAs of four warps per tile (the
// upper and the lower 16 x 8 pixels), two ring stages, four CTAs per SM: the same sixteen warps per SM in four barrier
// groups instead of two - a batch ends when the slowest of four warps is done, not the slowest of eight.
template <bool GATHER, int HALVES>
struct BwdSmem {
    static constexpr int NW = BWD_WARPS / HALVES, NST = HALVES == 2 ? 2 : STAGES;
    FwdStage<GATHER> stage[NST];                    // first: the gather destinations need 128-byte alignment
    float2 pair[NW][32 * PAIR_STRIDE];
    float part[NW][N_GRADS][BATCH];
    float4 pixc[NW][32];
    uint32_t gid[NST][BATCH];
    uint64_t full_bar[NST];
    int max_n;
};

template <bool GATHER, int HALVES>
__global__ void __launch_bounds__(CTA_THREADS / HALVES, 2 * HALVES)
composite_bwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int* __restrict__ tile_order,
                     const int2* __restrict__ ranges, const float4* __restrict__ sorted_records,
                     const __grid_constant__ CUtensorMap record_map, const uint32_t* __restrict__ sorted_gids,
                     float3 bg, float alpha_max, const float* __restrict__ state_T,
                     const int* __restrict__ state_n, const float* __restrict__ g_image,
                     const float* __restrict__ g_depth, const float* __restrict__ g_alpha,
                     float* __restrict__ grad2d) {
    frb_pdl_prologue();
    constexpr int RS = FwdStage<GATHER>::RS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    using Smem = BwdSmem<GATHER, HALVES>;
    constexpr int NW = Smem::NW, NST = Smem::NST, NT = 32 * NW;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tile_slot = blockIdx.x / HALVES, half_of_tile = blockIdx.x % HALVES;
    const int tile = tile_order ? tile_order[tile_slot] : tile_slot;     // heaviest tiles first
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a warp owns a BWD_FW x BWD_FH pixel block of the tile: the squarer the block, the fewer list entries
    // whose rectangle touches it (phase 1 walks only those)
    const int wx0 = tx * TILE + (warp % (TILE / BWD_FW)) * BWD_FW,
              wy0 = ty * TILE + half_of_tile * (TILE / HALVES) + (warp / (TILE / BWD_FW)) * BWD_FH;
    const int wx1 = wx0 + BWD_FW, wy1 = wy0 + BWD_FH;
    const int px = wx0 + (lane % BWD_FW);
    const int py = wy0 + (lane / BWD_FW);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;
    const float wbx = (float)wx0, wby = (float)wy0;                              // warp's pixel block origin

    const int2 range = ranges[tile];

    float T_final = 1.0f, gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
    int my_n = 0;
    if (in_image) {
        const size_t hw = (size_t)width * height;
        const size_t pix = (size_t)view * hw + (size_t)py * width + px;
        const size_t ip = (size_t)view * 3 * hw + (size_t)py * width + px;
        T_final = state_T[pix];
        const int st = state_n[pix];
        my_n = st & STATE_N_MASK;
        const int gates = st >> STATE_GATE_SHIFT;   // clamp(image, 0, 1) backward, DR:674-675
        gr = (gates & 1) ? g_image[ip] : 0.0f;
        gg = (gates & 2) ? g_image[ip + hw] : 0.0f;
        gb = (gates & 4) ? g_image[ip + 2 * hw] : 0.0f;
        gd = g_depth ? g_depth[pix] : 0.0f;
        ga = g_alpha ? g_alpha[pix] : 0.0f;
    }
    sm.pixc[warp][lane] = make_float4(gr, gg, gb, gd);
    if (threadIdx.x == 0) {
        sm.max_n = 0;
        for (int s = 0; s < NST; ++s) frb_mbar_init(&sm.full_bar[s], 1);
        frb_mbar_fence_init();
    }
    __syncthreads();
    int warp_n = my_n;                                       // entries applied by at least one pixel of the warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_n = max(warp_n, __shfl_xor_sync(0xffffffffu, warp_n, o));
    if (lane == 0) atomicMax(&sm.max_n, warp_n);
    __syncthreads();
    const int count = min(sm.max_n, range.y - range.x);
    const int n_batches = (count + BATCH - 1) / BATCH;
    if (n_batches == 0) return;

    // batches are visited last to first; ring slot (visit % NST) holds visit.  Producer: thread 0 (bulk copy of the
    // sorted records) or lanes 0..15 of warp 0 (gather4 by Gaussian id, ids prefetched one visit ahead).
    // (the producer of visit v is warp v % 8, so that its few dozen instructions per batch do not always delay the
    // same warp at the batch barrier)
    uint4 pre_g = make_uint4(0u, 0u, 0u, 0u);
    auto load_ids = [&](int visit) {
        if (GATHER && warp == (visit & (NW - 1)) && lane < BATCH / 4 && visit < n_batches) {
            const int b = n_batches - 1 - visit;
            const int i0 = b * BATCH + 4 * lane, last = count - 1;
            const uint32_t* g = sorted_gids + range.x;
            pre_g = make_uint4(g[min(i0, last)], g[min(i0 + 1, last)], g[min(i0 + 2, last)], g[min(i0 + 3, last)]);
        }
    };
    auto issue = [&](int visit) {
        const int b = n_batches - 1 - visit;
        const int s = visit % NST;
        const int cnt = min(BATCH, count - b * BATCH);
        if (!GATHER) {
            if (threadIdx.x == 0) {
                frb_mbar_expect_tx(&sm.full_bar[s], cnt * RECORD_BYTES);
                frb_tma_load_1d(sm.stage[s].rec, sorted_records + 3 * (size_t)(range.x + b * BATCH), cnt * RECORD_BYTES,
                                &sm.full_bar[s]);
            }
        } else if (warp == (visit & (NW - 1))) {
            const int groups = (cnt + 3) >> 2;
            if (lane == 0) frb_mbar_expect_tx(&sm.full_bar[s], groups * 4 * RS * 16);
            __syncwarp();
            if (lane < groups)
                frb_tma_gather4(sm.stage[s].rec + 4 * RS * lane, &record_map, pre_g, &sm.full_bar[s]);
        }
    };
    for (int v = 0; v < NST && v < n_batches; ++v) {
        load_ids(v);
        issue(v);
    }
    load_ids(NST);

    const float X = gr * bg.x + gg * bg.y + gb * bg.z - ga;
    const float TX = T_final * X;
    float T = T_final;
    float SX = TX;                      // S + T_final X, S = sum_{j>i} c_j w_j: carried as one sum (one add less per pair)
    float2* my_pair = sm.pair[warp];

    for (int visit = 0; visit < n_batches; ++visit) {
        const int b = n_batches - 1 - visit;
        const int s = visit % NST;
        const int cnt = min(BATCH, count - b * BATCH);
        if (threadIdx.x < cnt) sm.gid[s][threadIdx.x] = sorted_gids[range.x + b * BATCH + threadIdx.x];
        frb_mbar_wait(&sm.full_bar[s], (visit / NST) & 1);
        const float4* rec = sm.stage[s].rec;

        for (int sb = (cnt - 1) >> 5; sb >= 0; --sb) {
            const int sub_cnt = min(32, cnt - sb * 32);
            // ---- candidates: lane l tests record l of the block against this warp's pixel block ----
            const int base_n = b * BATCH + sb * 32;
            uint32_t cand;
            {
                bool ok = false;
                if (lane < sub_cnt) {
                    const uint32_t lo = __float_as_uint(rec[RS * (sb * 32 + lane) + 1].z);
                    const uint32_t hi = __float_as_uint(rec[RS * (sb * 32 + lane) + 1].w) & 0x7fff7fffu;
                    ok = (int)(lo & 0xffff) < wx1 && (int)(hi & 0xffff) > wx0 && (int)(lo >> 16) < wy1 &&
                         (int)(hi >> 16) > wy0;
                }
                cand = __ballot_sync(0xffffffffu, ok);
                const int wn = warp_n - base_n;              // entries >= wn were applied by no pixel of the warp
                if (wn < 32) cand &= (wn <= 0) ? 0u : ((1u << wn) - 1u);
            }
            // ---- phase 1: lane = pixel, candidates back to front ----------------------------
            const uint32_t gmask = cand;
            const int local_n = my_n - base_n;                   // entries j < local_n were applied
            // Two candidates per iteration: everything that does not depend on the running (T, S) is
            // evaluated for both first (independent instruction streams hide the LDS / MUFU latency),
            // then the two short serial tails run back to back.
            // (addresses as one multiply-add from a per-block / per-lane base, and the candidate index from bfind:
            // 31 - __clz() compiled to FLO + IADD3 + LOP3 - four instructions less per candidate out of 48)
            struct Pre { float a, inv_om, w, gpass; bool active; };
            const char* rec_sb = reinterpret_cast<const char*>(rec + RS * (sb * 32));
            char* pair_lane = reinterpret_cast<char*>(my_pair + lane);
            auto stage_a = [&](int j) {
                Pre p;
                const float4* rj = reinterpret_cast<const float4*>(rec_sb + j * (RS * 16));
                const float4 r1 = rj[1], r2 = rj[2];
                p.active = (j < local_n) &&
                           rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.z), __float_as_uint(r1.w));
                const float4 r0 = rj[0];
                const float dx = fpx - r0.x, dy = fpy - r0.y;
                const float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                const float g = frb_ex2(power);
                const float araw = g * r1.y;
                p.a = fminf(fmaxf(araw, 0.0f), alpha_max);
                p.gpass = (p.a == araw) ? g : 0.0f;           // g, or 0 behind the clamp gate 0 <= g*o <= 0.99
                p.inv_om = frb_rcp(1.0f - p.a);
                // gC . rgb + gD depth as two packed operations and an add: (r, g) and (b, depth) are neighbours
                const float2 wp = frb_fma2(make_float2(gb, gd), make_float2(r2.z, r2.w),
                                           frb_mul2(make_float2(gr, gg), make_float2(r2.x, r2.y)));
                p.w = wp.x + wp.y;
                return p;
            };
            auto stage_b = [&](int j, const Pre& p) {
                float2 out = make_float2(0.f, 0.f);
                if (p.active) {
                    const float Ti = T * p.inv_om;
                    const float c = p.a * Ti;
                    const float dalpha = Ti * p.w - SX * p.inv_om;
                    SX = fmaf(c, p.w, SX);
                    T = Ti;
                    out.x = c;
                    out.y = p.gpass * dalpha;                 // g * gated dL/dalpha (= dL/dopacity part)
                }
                *reinterpret_cast<float2*>(pair_lane + j * (PAIR_STRIDE * 8)) = out;
            };
            while (__popc(cand) >= QUAD_MIN) {         // four candidates per iteration while at least four remain
                const int j0 = frb_bfind(cand);
                cand ^= 1u << j0;
                const int j1 = frb_bfind(cand);
                cand ^= 1u << j1;
                const int j2 = frb_bfind(cand);
                cand ^= 1u << j2;
                const int j3 = frb_bfind(cand);
                cand ^= 1u << j3;
                const Pre p0 = stage_a(j0), p1 = stage_a(j1), p2 = stage_a(j2), p3 = stage_a(j3);
                stage_b(j0, p0);
                stage_b(j1, p1);
                stage_b(j2, p2);
                stage_b(j3, p3);
            }
            while (cand) {
                const int j0 = frb_bfind(cand);
                cand ^= 1u << j0;
                if (cand) {
                    const int j1 = frb_bfind(cand);
                    cand ^= 1u << j1;
                    const Pre p0 = stage_a(j0), p1 = stage_a(j1);
                    stage_b(j0, p0);
                    stage_b(j1, p1);
                } else {
                    const Pre p0 = stage_a(j0);
                    stage_b(j0, p0);
                }
            }
            __syncwarp();
            // ---- phase 2: lane = Gaussian ------------------------------------------------
            float d_u = 0.f, d_v = 0.f, d_A = 0.f, d_B = 0.f, d_C = 0.f, d_o = 0.f, d_dep = 0.f, d_r = 0.f,
                  d_g = 0.f, d_b = 0.f;
            if ((gmask >> lane) & 1u) {
                const int jb = sb * 32 + lane;
                const float4 r0 = rec[RS * jb + 0], r1 = rec[RS * jb + 1];
                const float oln2 = r1.y * FRB_LN2;
                const float ux = wbx - r0.x, uy = wby - r0.y;
                const float2* row = my_pair + lane * PAIR_STRIDE;
                const float4* pc = sm.pixc[warp];
                // colour / depth sums as fp32 pairs (FFMA2: two fmas per issue slot, same bits as fmaf).  The geometric
                // sums over dx = ux + i, dy = uy + k (i, k: the pixel's place in the 8x4 block) are accumulated as the
                // six moments sum dpow i^a k^b with i and k as IMMEDIATE operands and combined once per Gaussian:
                // 4 fmas per pixel on average (zero powers drop out at compile time) where forming dx, dy and the
                // five products took 6 instructions
                float2 s_rg = make_float2(0.f, 0.f), s_bd = s_rg;
                float m00 = 0.f, m10 = 0.f, m01 = 0.f, m20 = 0.f, m11 = 0.f, m02 = 0.f;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    const float2 cd = row[p];               // (c, g * gated dL/dalpha), formed in phase 1
                    const float4 gpix = pc[p];
                    const int i = p % BWD_FW, k = p / BWD_FW;
                    s_rg = frb_fma2s(cd.x, make_float2(gpix.x, gpix.y), s_rg);
                    s_bd = frb_fma2s(cd.x, make_float2(gpix.z, gpix.w), s_bd);
                    m00 += cd.y;
                    if (i) m10 = fmaf(cd.y, (float)i, m10);
                    if (k) m01 = fmaf(cd.y, (float)k, m01);
                    if (i) m20 = fmaf(cd.y, (float)(i * i), m20);
                    if (i && k) m11 = fmaf(cd.y, (float)(i * k), m11);
                    if (k) m02 = fmaf(cd.y, (float)(k * k), m02);
                }
                d_o = m00;
                // sum dpow dx, dy, dx^2, dx dy, dy^2 from the moments
                const float2 s_xy = make_float2(fmaf(ux, m00, m10), fmaf(uy, m00, m01));
                const float2 s_AB = make_float2(fmaf(ux, fmaf(ux, m00, 2.0f * m10), m20),
                                                fmaf(ux, fmaf(uy, m00, m01), fmaf(uy, m10, m11)));
                d_C = fmaf(uy, fmaf(uy, m00, 2.0f * m01), m02);
                d_r = s_rg.x; d_g = s_rg.y; d_b = s_bd.x; d_dep = s_bd.y;
                // dL/d(power) = gda * o * ln2 (g = 2^power); u, v enter through dx, dy
                d_A = s_AB.x * oln2; d_B = s_AB.y * oln2; d_C *= oln2;
                d_u = -(2.0f * r0.z * s_xy.x + r0.w * s_xy.y) * oln2;
                d_v = -(r0.w * s_xy.x + 2.0f * r1.x * s_xy.y) * oln2;
            }
            {
                // (shared-memory float atomics compile to CAS loops on sm_100a: plain stores, summed below)
                float* pw = &sm.part[warp][0][sb * 32 + lane];
                pw[0 * BATCH] = d_u; pw[1 * BATCH] = d_v; pw[2 * BATCH] = d_A; pw[3 * BATCH] = d_B;
                pw[4 * BATCH] = d_C; pw[5 * BATCH] = d_o; pw[6 * BATCH] = d_dep; pw[7 * BATCH] = d_r;
                pw[8 * BATCH] = d_g; pw[9 * BATCH] = d_b;
            }
            __syncwarp();   // pair tile is reused by the next sub-block
        }
        __syncthreads();    // partial sums of all warps (and gid) visible
        // one thread per (Gaussian, float4 of its grad2d row [du dv dA dB | dC do ddepth _ | dr dg db _]):
        // the eight warps' partial sums are added and leave as ONE 16-byte vector reduction (red.global.add.v4.f32)
        // instead of up to four scalar atomics
        for (int t = threadIdx.x; t < 3 * BATCH; t += NT) {
            const int q = t / BATCH, jb = t - q * BATCH;
            if (jb < cnt) {
                const int v0 = (q == 0) ? 0 : (q == 1 ? 4 : 7);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    sum.x += sm.part[w][v0][jb];
                    sum.y += sm.part[w][v0 + 1][jb];
                    sum.z += sm.part[w][v0 + 2][jb];
                    if (q == 0) sum.w += sm.part[w][3][jb];
                }
                if (sum.x != 0.0f || sum.y != 0.0f || sum.z != 0.0f || sum.w != 0.0f)
                    frb_red_add_f4(grad2d + (size_t)sm.gid[s][jb] * FRB_GRAD_FLOATS + 4 * q, sum);
            }
        }
        __syncthreads();    // stage s, gid[s] and part are free
        if (visit + NST < n_batches) {
            issue(visit + NST);
            load_ids(visit + NST + 1);
        }
    }
}

// Launch order of the tiles: descending list length (bucketed), so the long centre tiles start first
// and the short ones fill the tail of the grid (longest-processing-time-first list scheduling).
constexpr int SCHED_BUCKETS = 1024;
__global__ void __launch_bounds__(1024) tile_schedule_kernel(int n_tiles, const int2* __restrict__ ranges,
                                                             int* __restrict__ order) {
    frb_pdl_prologue();
    __shared__ int hist[SCHED_BUCKETS];
    __shared__ int warp_tot[32];
    hist[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [](int2 r) { return SCHED_BUCKETS - 1 - min((r.y - r.x) >> 3, SCHED_BUCKETS - 1); };
    for (int t = threadIdx.x; t < n_tiles; t += 1024) atomicAdd(&hist[bucket(ranges[t])], 1);
    __syncthreads();
    // exclusive scan of the 1024 buckets (bucket 0 = longest lists)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v = hist[threadIdx.x], incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;
    }
    __syncthreads();
    hist[threadIdx.x] = warp_tot[warp] + incl - v;
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += 1024) order[atomicAdd(&hist[bucket(ranges[t])], 1)] = t;
}

}  // namespace

extern "C" int frb_tile_schedule(int n_tiles, const int32_t* ranges, int32_t* tile_order, void* stream) {
    if (n_tiles < 0 || !ranges || !tile_order) return FRB_E_INVALID;
    if (n_tiles == 0) return 0;
    frb_launch(tile_schedule_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, n_tiles, (const int2*)ranges, tile_order);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// composite_phase.cu
int frb_composite_phase_fwd_launch(int n_views, int width, int height, const int32_t* tile_order,
                                   const int32_t* ranges, const float* sorted_records, const float* sorted_phases, float phase_amplitude,
                                   const float* background_host, float t_eps, float* image, float* depth,
                                   float* alpha, float* state_T, int32_t* state_n, float* ckpt, cudaStream_t st);
int frb_composite_phase_bwd_launch(int n_views, int width, int height, const int32_t* tile_order,
                                   const int32_t* ranges, const float* sorted_records, const uint32_t* sorted_gids,
                                   const float* sorted_phases, float phase_amplitude, const float* background_host,
                                   const float* state_T, const int32_t* state_n, const float* ckpt,
                                   const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                                   float* g_phases, cudaStream_t st);

static int check_image_args(int n_views, int width, int height) {
    if (n_views < 1 || n_views > FRB_MAX_VIEWS || width < 1 || height < 1) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    return 0;
}

extern "C" int frb_composite_fwd(int n_views, int width, int height, const int32_t* ranges,
                                 const float* sorted_records, const float* sorted_phases,
                                 float phase_amplitude, const float* background_host, float t_eps,
                                 float* image, float* depth, float* alpha, float* state_T, int32_t* state_n,
                                 float* ckpt, void* stream) {
    return frb_composite_fwd_sched(n_views, width, height, nullptr, ranges, sorted_records, sorted_phases,
                                   phase_amplitude, background_host, t_eps, image, depth, alpha, state_T, state_n,
                                   ckpt, stream);
}

extern "C" int frb_composite_fwd_sched(int n_views, int width, int height, const int32_t* tile_order,
                                 const int32_t* ranges, const float* sorted_records, const float* sorted_phases,
                                 float phase_amplitude, const float* background_host, float t_eps,
                                 float* image, float* depth, float* alpha, float* state_T, int32_t* state_n,
                                 float* ckpt, void* stream) {
    return frb_composite_fwd_cap(n_views, width, height, tile_order, ranges, sorted_records, sorted_phases,
                                 phase_amplitude, background_host, t_eps, FRB_ALPHA_MAX, image, depth, alpha, state_T,
                                 state_n, ckpt, stream);
}

extern "C" int frb_composite_fwd_cap(int n_views, int width, int height, const int32_t* tile_order,
                                 const int32_t* ranges, const float* sorted_records, const float* sorted_phases,
                                 float phase_amplitude, const float* background_host, float t_eps, float alpha_max,
                                 float* image, float* depth, float* alpha, float* state_T, int32_t* state_n,
                                 float* ckpt, void* stream) {
    if (!(alpha_max > 0.0f && alpha_max < 1.0f)) return FRB_E_INVALID;   // the backward divides by 1 - alpha
    if (sorted_phases && alpha_max != FRB_ALPHA_MAX) return FRB_E_INVALID;
    int rc = check_image_args(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !background_host || !image || !depth || !alpha || !state_T || !state_n) return FRB_E_INVALID;
    if (sorted_phases)
        return frb_composite_phase_fwd_launch(n_views, width, height, tile_order, ranges, sorted_records, sorted_phases,
                                              phase_amplitude, background_host, t_eps, image, depth, alpha,
                                              state_T, state_n, ckpt, (cudaStream_t)stream);
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    CUtensorMap no_map = {};
    frb_launch(composite_fwd_kernel<false, false>, dim3(n_views * tpv), dim3(CTA_THREADS), 0, (cudaStream_t)stream,
        width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)sorted_records, no_map,
        (const uint32_t*)nullptr, bg, t_eps, alpha_max, image, depth, alpha, state_T, state_n);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// The same compositor reading the tile lists as Gaussian ids: records[n_records] is the UNSORTED record array of
// frb_project_fwd, sorted_gids the per-tile lists; the records are fetched by TMA tile::gather4.
extern "C" int frb_composite_fwd_gather(int n_views, int width, int height, const int32_t* tile_order,
                                        const int32_t* ranges, const float* records, int n_records,
                                        const uint32_t* sorted_gids, const float* background_host, float t_eps,
                                        float alpha_max, float* image, float* depth, float* alpha, float* state_T,
                                        int32_t* state_n, void* stream) {
    if (!(alpha_max > 0.0f && alpha_max < 1.0f)) return FRB_E_INVALID;
    int rc = check_image_args(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !background_host || !image || !depth || !alpha || !state_T || !state_n || !sorted_gids)
        return FRB_E_INVALID;
    CUtensorMap map;
    if ((rc = frb_record_tensor_map(records, n_records, &map))) return rc;
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    static const bool by_record = getenv("FRB_FWD_BY_RECORD") && getenv("FRB_FWD_BY_RECORD")[0] == '1';
    if (by_record)
        frb_launch(composite_fwd_kernel<true, false>, dim3(n_views * tpv), dim3(CTA_THREADS), 0, (cudaStream_t)stream,
            width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)nullptr, map, sorted_gids, bg,
            t_eps, alpha_max, image, depth, alpha, state_T, state_n);
    else
        frb_launch(composite_fwd_kernel<true, true>, dim3(n_views * tpv), dim3(CTA_THREADS), 0, (cudaStream_t)stream,
            width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)nullptr, map, sorted_gids, bg,
            t_eps, alpha_max, image, depth, alpha, state_T, state_n);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_composite_bwd(int n_views, int width, int height, const int32_t* ranges,
                                 const float* sorted_records, const uint32_t* sorted_gids,
                                 const float* sorted_phases, float phase_amplitude,
                                 const float* background_host, const float* state_T,
                                 const int32_t* state_n, const float* ckpt, const float* g_image,
                                 const float* g_depth, const float* g_alpha, float* grad2d, float* g_phases,
                                 void* stream) {
    return frb_composite_bwd_sched(n_views, width, height, nullptr, ranges, sorted_records, sorted_gids,
                                   sorted_phases, phase_amplitude, background_host, state_T, state_n, ckpt, g_image,
                                   g_depth, g_alpha, grad2d, g_phases, stream);
}

extern "C" int frb_composite_bwd_sched(int n_views, int width, int height, const int32_t* tile_order,
                                 const int32_t* ranges, const float* sorted_records, const uint32_t* sorted_gids,
                                 const float* sorted_phases, float phase_amplitude,
                                 const float* background_host, const float* state_T,
                                 const int32_t* state_n, const float* ckpt, const float* g_image,
                                 const float* g_depth, const float* g_alpha, float* grad2d, float* g_phases,
                                 void* stream) {
    return frb_composite_bwd_cap(n_views, width, height, tile_order, ranges, sorted_records, sorted_gids,
                                 sorted_phases, phase_amplitude, background_host, FRB_ALPHA_MAX, state_T, state_n, ckpt,
                                 g_image, g_depth, g_alpha, grad2d, g_phases, stream);
}

extern "C" int frb_composite_bwd_cap(int n_views, int width, int height, const int32_t* tile_order,
                                 const int32_t* ranges, const float* sorted_records, const uint32_t* sorted_gids,
                                 const float* sorted_phases, float phase_amplitude,
                                 const float* background_host, float alpha_max, const float* state_T,
                                 const int32_t* state_n, const float* ckpt, const float* g_image,
                                 const float* g_depth, const float* g_alpha, float* grad2d, float* g_phases,
                                 void* stream) {
    if (!(alpha_max > 0.0f && alpha_max < 1.0f)) return FRB_E_INVALID;
    if (sorted_phases && alpha_max != FRB_ALPHA_MAX) return FRB_E_INVALID;
    int rc = check_image_args(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !background_host || !state_T || !state_n || !g_image || !grad2d) return FRB_E_INVALID;
    if (sorted_phases)
        return frb_composite_phase_bwd_launch(n_views, width, height, tile_order, ranges, sorted_records, sorted_gids,
                                              sorted_phases, phase_amplitude, background_host, state_T, state_n,
                                              ckpt, g_image, g_depth, g_alpha, grad2d, g_phases,
                                              (cudaStream_t)stream);
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    static unsigned long long smem_opted_in = 0;          // per-device bitmask (the attribute is per device)
    FRB_CUDA_OK(frb_opt_in_smem(composite_bwd_kernel<false, 1>, (int)sizeof(BwdSmem<false, 1>), &smem_opted_in));
    CUtensorMap no_map = {};
    frb_launch(composite_bwd_kernel<false, 1>, dim3(n_views * tpv), dim3(CTA_THREADS), sizeof(BwdSmem<false, 1>),
        (cudaStream_t)stream, width, height, tiles_x, tpv, tile_order, (const int2*)ranges,
        (const float4*)sorted_records, no_map, sorted_gids, bg, alpha_max, state_T, state_n, g_image, g_depth, g_alpha,
        grad2d);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

// frb_composite_bwd_cap reading the records of the tile lists by TMA tile::gather4 from the unsorted record array.
extern "C" int frb_composite_bwd_gather(int n_views, int width, int height, const int32_t* tile_order,
                                        const int32_t* ranges, const float* records, int n_records,
                                        const uint32_t* sorted_gids, const float* background_host, float alpha_max,
                                        const float* state_T, const int32_t* state_n, const float* g_image,
                                        const float* g_depth, const float* g_alpha, float* grad2d, void* stream) {
    if (!(alpha_max > 0.0f && alpha_max < 1.0f)) return FRB_E_INVALID;
    int rc = check_image_args(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !background_host || !state_T || !state_n || !g_image || !grad2d || !sorted_gids) return FRB_E_INVALID;
    CUtensorMap map;
    if ((rc = frb_record_tensor_map(records, n_records, &map))) return rc;
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    static unsigned long long smem_opted_in = 0;
    static unsigned long long smem_opted_in_halves = 0;
    // default: two CTAs per tile (0.168 -> 0.164 ms, resident 2942 -> 2987, host to host 3131 -> 3202 frames/s on one box,
    // profiles/r4_d_*); FRB_BWD_HALVES=1 launches the one-CTA form
    static const bool halves = !(getenv("FRB_BWD_HALVES") && getenv("FRB_BWD_HALVES")[0] == '1');
    if (halves) {
        FRB_CUDA_OK(frb_opt_in_smem(composite_bwd_kernel<true, 2>, (int)sizeof(BwdSmem<true, 2>), &smem_opted_in_halves));
        frb_launch(composite_bwd_kernel<true, 2>, dim3(2 * n_views * tpv), dim3(CTA_THREADS / 2), sizeof(BwdSmem<true, 2>),
            (cudaStream_t)stream, width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)nullptr,
            map, sorted_gids, bg, alpha_max, state_T, state_n, g_image, g_depth, g_alpha, grad2d);
    } else {
        FRB_CUDA_OK(frb_opt_in_smem(composite_bwd_kernel<true, 1>, (int)sizeof(BwdSmem<true, 1>), &smem_opted_in));
        frb_launch(composite_bwd_kernel<true, 1>, dim3(n_views * tpv), dim3(CTA_THREADS), sizeof(BwdSmem<true, 1>),
            (cudaStream_t)stream, width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)nullptr,
            map, sorted_gids, bg, alpha_max, state_T, state_n, g_image, g_depth, g_alpha, grad2d);
    }
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
