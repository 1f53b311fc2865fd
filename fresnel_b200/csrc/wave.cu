// Complex wave-field splatting (order-free additive accumulation), forward and backward.
// Reference: WaveFieldRenderer.forward DR:747-926 and the per-plane splat loop of
// ASMWaveFieldRenderer.forward DR:1233-1283 (same arithmetic, one complex field per depth plane).
//
// Per Gaussian and pixel of its rectangle:  amp = g * opacity  (no clamp, no transmittance)
//     Re_c += amp * colour_c * cos(phi_c) ;  Im_c += amp * colour_c * sin(phi_c)
//     [wave only]  D += amp * depth ;  Wsum += amp
// colour_c cos / sin(phi_c) are formed once per Gaussian (frb_wave_prepare, accurate sincosf) and
// travel with the sorted instance list as a 32-byte side record.
//
// Forward: one CTA per 16x16 tile, one thread per pixel, both record streams staged by TMA bulk
// copies.  ASM: the tile list is sorted by depth plane; the accumulator is flushed to the plane's
// complex field whenever the plane changes.
// Backward: no order dependence, so it is a gather: one CTA per tile stages the tile's per-pixel
// upstream gradients in shared memory; wave: one THREAD per list entry walks the entry's rectangle inside the
// tile and keeps its thirteen sums in registers; ASM (short per-plane groups): one WARP per entry, lanes =
// pixels of a 4/8/16-wide patch, butterfly reduction.  One atomic per value per (Gaussian, tile).
#include "composite_common.cuh"

namespace {

constexpr int WC_FLOATS = 8;
constexpr int WC_BYTES = WC_FLOATS * 4;

struct WaveStage {
    float4 rec[BATCH * 3];
    float4 wc[BATCH * 2];
};

__global__ void __launch_bounds__(256)
wave_prepare_kernel(int n, const float* __restrict__ colors, const float* __restrict__ phases, int phase_stride,
                    float4* __restrict__ wc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float cc[3], cs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ph = phases[(size_t)i * phase_stride + (phase_stride == 3 ? c : 0)];
        float sn, cn;
        sincosf(ph, &sn, &cn);
        float col = colors[3 * (size_t)i + c];
        cc[c] = col * cn;
        cs[c] = col * sn;
    }
    wc[2 * (size_t)i + 0] = make_float4(cc[0], cc[1], cc[2], cs[0]);
    wc[2 * (size_t)i + 1] = make_float4(cs[1], cs[2], 0.f, 0.f);
}

__global__ void __launch_bounds__(256)
gather_wc_kernel(int m, const uint32_t* __restrict__ gids, const float4* __restrict__ wc,
                 float4* __restrict__ sorted_wc) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2ll * m) return;
    int i = (int)(t >> 1), part = (int)(t & 1);
    sorted_wc[t] = wc[2 * (size_t)gids[i] + part];
}

// ASM = false: accum[view][8][H][W] planar (Re rgb, Im rgb, D, Wsum) + per-view max of sqrt(I + 1e-8).
// ASM = true : fields[view][plane][3][H][W] float2 (Re, Im); plane = low word of the sorted key.
template <bool ASM>
__global__ void __launch_bounds__(CTA_THREADS)
wave_splat_fwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int2* __restrict__ ranges,
                      const float4* __restrict__ sorted_records, const float4* __restrict__ sorted_wc,
                      const uint64_t* __restrict__ keys, int n_planes, float* __restrict__ accum,
                      uint32_t* __restrict__ rmax_bits, float2* __restrict__ fields) {
    __shared__ __align__(16) WaveStage stage[STAGES];
    __shared__ int group_end[64];          // ASM: list position (relative to the tile's range) where plane p ends
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tile = blockIdx.x;
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + foot_x(threadIdx.x);
    const int py = ty * TILE + foot_y(threadIdx.x);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;
    const size_t hw = (size_t)width * height;
    const size_t pix_in_view = (size_t)py * width + px;

    const int2 range = ranges[tile];
    const int count = range.y - range.x;
    const int n_batches = (count + BATCH - 1) / BATCH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) frb_mbar_init(&full_bar[s], 1);
        frb_mbar_fence_init();
    }
    if (ASM && threadIdx.x < n_planes) {
        // the tile list is sorted by plane (low word of the key): upper bound of plane p by binary search, so that
        // the inner loop below runs over whole plane groups without a per-record plane test
        int lo = range.x, hi = range.y;
        const uint32_t p = threadIdx.x;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if ((uint32_t)keys[mid] <= p) lo = mid + 1; else hi = mid;
        }
        group_end[threadIdx.x] = (threadIdx.x == n_planes - 1) ? count : lo - range.x;
    }
    __syncthreads();
    auto issue = [&](int b) {
        int s = b % STAGES;
        int cnt = min(BATCH, count - b * BATCH);
        frb_mbar_expect_tx(&full_bar[s], cnt * (RECORD_BYTES + WC_BYTES));
        frb_tma_load_1d(stage[s].rec, sorted_records + 3 * (size_t)(range.x + b * BATCH), cnt * RECORD_BYTES,
                        &full_bar[s]);
        frb_tma_load_1d(stage[s].wc, sorted_wc + 2 * (size_t)(range.x + b * BATCH), cnt * WC_BYTES, &full_bar[s]);
    };
    if (threadIdx.x == 0)
        for (int b = 0; b < STAGES && b < n_batches; ++b) issue(b);

    // (scalar FMAs on purpose: as FFMA2 pairs this loop ran 9 % slower - FFMA2 issues on the heavy FMA pipe only,
    // which this kernel already keeps at 45 %)
    float re0 = 0.f, re1 = 0.f, re2 = 0.f, im0 = 0.f, im1 = 0.f, im2 = 0.f, ad = 0.f, aw = 0.f;
    int grp = 0;                           // ASM: plane whose group is being accumulated

    // ASM: every (pixel, plane) of the tile is written exactly once, zeros included: no memset of the fields
    auto flush = [&](int plane) {
        if (ASM && in_image) {
            float2* f = fields + ((size_t)view * n_planes + plane) * 3 * hw + pix_in_view;
            f[0] = make_float2(re0, im0);
            f[hw] = make_float2(re1, im1);
            f[2 * hw] = make_float2(re2, im2);
        }
        re0 = re1 = re2 = im0 = im1 = im2 = 0.f;
    };

    for (int b = 0; b < n_batches; ++b) {
        const int s = b % STAGES;
        const int cnt = min(BATCH, count - b * BATCH);
        frb_mbar_wait(&full_bar[s], (b / STAGES) & 1);
        const float4* rec = stage[s].rec;
        const float4* wcs = stage[s].wc;
        int j = 0;
        while (j < cnt) {
            int run_end = cnt;
            if (ASM) {
                const int pos = b * BATCH + j;
                while (pos >= group_end[grp]) flush(grp++);      // uniform; group_end[n_planes - 1] = count > pos
                run_end = min(cnt, group_end[grp] - b * BATCH);
            }
#pragma unroll 4
            for (; j < run_end; ++j) {
                float4 r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                if (rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.z), __float_as_uint(r1.w))) {
                    float4 r0 = rec[3 * j + 0];
                    float4 wa = wcs[2 * j + 0], wb = wcs[2 * j + 1];
                    float dx = fpx - r0.x, dy = fpy - r0.y;
                    float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                    float amp = frb_ex2(power) * r1.y;
                    re0 = fmaf(amp, wa.x, re0); re1 = fmaf(amp, wa.y, re1); re2 = fmaf(amp, wa.z, re2);
                    im0 = fmaf(amp, wa.w, im0); im1 = fmaf(amp, wb.x, im1); im2 = fmaf(amp, wb.y, im2);
                    if (!ASM) {
                        ad = fmaf(amp, r2.w, ad);
                        aw += amp;
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && b + STAGES < n_batches) issue(b + STAGES);
    }

    if (ASM) {
        for (; grp < n_planes; ++grp) flush(grp);
        return;
    }
    float rm = 0.0f;
    if (in_image) {
        float* a = accum + (size_t)view * 8 * hw + pix_in_view;
        a[0] = re0; a[hw] = re1; a[2 * hw] = re2; a[3 * hw] = im0; a[4 * hw] = im1; a[5 * hw] = im2;
        a[6 * hw] = ad; a[7 * hw] = aw;
        // sqrt(I + 1e-8), DR:894-898 (correctly rounded so that finish kernels recompute the same bits)
        float r0 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(re0, re0), __fmul_rn(im0, im0)), 1e-8f));
        float r1 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(re1, re1), __fmul_rn(im1, im1)), 1e-8f));
        float r2 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(re2, re2), __fmul_rn(im2, im2)), 1e-8f));
        rm = fmaxf(r0, fmaxf(r1, r2));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
    if ((threadIdx.x & 31) == 0) atomicMax(rmax_bits + view, __float_as_uint(rm));
}

// Thirteen sums per list entry over the pixels of its rectangle inside this tile.
// gpix: ASM = false: [view][8][H][W] (dRe rgb, dIm rgb, dD, dW); ASM = true: [view][plane][3][H][W] float2.
template <bool ASM>
__global__ void __launch_bounds__(CTA_THREADS, 4)
wave_splat_bwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int2* __restrict__ ranges,
                      const float4* __restrict__ sorted_records, const float4* __restrict__ sorted_wc,
                      const uint32_t* __restrict__ sorted_gids, const uint64_t* __restrict__ keys, int n_planes,
                      const float* __restrict__ gpix, float scale, float* __restrict__ grad2d,
                      float* __restrict__ gwc) {
    __shared__ float4 gp_a[CTA_THREADS];   // dRe r, g, b, dIm r
    __shared__ float4 gp_b[CTA_THREADS];   // dIm g, b, dD, dW
    __shared__ int group_start[64 + 1];

    const int tile = blockIdx.x;
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + (threadIdx.x & (TILE - 1));
    const int py = ty * TILE + (threadIdx.x / TILE);
    const bool in_image = (px < width) && (py < height);
    const size_t hw = (size_t)width * height;
    const size_t pix_in_view = (size_t)py * width + px;
    const int2 range = ranges[tile];
    if (range.y == range.x) return;

    const int n_groups = ASM ? n_planes : 1;
    if (ASM) {
        // group_start[p] = first entry of the tile list whose plane >= p (binary search, one thread per plane)
        if (threadIdx.x <= n_planes) {
            int lo = range.x, hi = range.y;
            const uint32_t p = threadIdx.x;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if ((uint32_t)keys[mid] < p) lo = mid + 1; else hi = mid;
            }
            group_start[threadIdx.x] = (threadIdx.x == n_planes) ? range.y : lo;
        }
    } else if (threadIdx.x == 0) {
        group_start[0] = range.x;
        group_start[1] = range.y;
    }
    __syncthreads();

    const float bx = (float)(tx * TILE), by = (float)(ty * TILE);
    const int tx0 = tx * TILE, ty0 = ty * TILE;

    for (int grp = 0; grp < n_groups; ++grp) {
        const int g0 = group_start[grp], g1 = group_start[grp + 1];
        if (g0 == g1) continue;                                  // uniform
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (in_image) {
            if (ASM) {
                const float2* f = reinterpret_cast<const float2*>(gpix) +
                                  ((size_t)view * n_planes + grp) * 3 * hw + pix_in_view;
                float2 c0 = f[0], c1 = f[hw], c2 = f[2 * hw];
                a = make_float4(c0.x * scale, c1.x * scale, c2.x * scale, c0.y * scale);
                b = make_float4(c1.y * scale, c2.y * scale, 0.f, 0.f);
            } else {
                const float* f = gpix + (size_t)view * 8 * hw + pix_in_view;
                a = make_float4(f[0], f[hw], f[2 * hw], f[3 * hw]);
                b = make_float4(f[4 * hw], f[5 * hw], f[6 * hw], f[7 * hw]);
            }
        }
        __syncthreads();                                         // previous group's readers are done
        gp_a[threadIdx.x] = a;
        gp_b[threadIdx.x] = b;
        __syncthreads();

        if constexpr (ASM) {
        // ASM lists are split into up to 16 plane groups of a few dozen entries: one thread per entry would leave
        // most of the CTA idle, so a whole warp takes an entry (measured at config 5: 3.56 -> 2.50 ms).
            // One WARP per list entry: the lanes cover the entry's rectangle inside the tile as a (cols x rows) patch
            // that is moved down the rectangle; cols = 4, 8 or 16 (the smallest that holds the rectangle's width), so
            // at least half of the lanes are on pixels of the rectangle and neighbouring lanes read neighbouring
            // shared-memory words.  The thirteen partial sums are reduced over the warp with the halving butterfly.
            // (round 2: HALF a warp per entry - at config 5 a rectangle clipped to the tile is ~8 x 8 pixels, i.e. two
            // trips of the pixel loop against ~130 instructions of fixed cost per entry (prefetch, rectangle, the
            // 16-shuffle butterfly, atomics); two entries per warp instruction halve that fixed cost.)
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int hl = lane & 15, upper_half = lane >> 4;
            constexpr int N_SUB = CTA_THREADS / 16;                // half-warps per CTA
            // the entry's 80 bytes come straight from global memory (L2): the next entry of this half-warp is loaded
            // before the current one is processed, so the load latency hides behind the pixel loop
            float4 n0, n1, n2, na, nb;
            uint32_t ngid = 0;
            auto load_entry = [&](int e) {
                n0 = sorted_records[3 * (size_t)e + 0]; n1 = sorted_records[3 * (size_t)e + 1];
                n2 = sorted_records[3 * (size_t)e + 2];
                na = sorted_wc[2 * (size_t)e + 0]; nb = sorted_wc[2 * (size_t)e + 1];
                ngid = sorted_gids[e];
            };
            if (g0 + 2 * warp + upper_half < g1) load_entry(g0 + 2 * warp + upper_half);
            for (int eb = g0 + 2 * warp; eb < g1; eb += N_SUB) {            // eb: the lower half's entry (uniform)
                const int e = eb + upper_half;
                const bool have = e < g1;
                const float4 r0 = n0, r1 = n1, r2 = n2, wa = na, wb = nb;
                const uint32_t gid = ngid;
                if (e + N_SUB < g1) load_entry(e + N_SUB);
                const uint32_t lo = __float_as_uint(r1.z), hi = __float_as_uint(r1.w) & 0x7fff7fffu;
                // rectangle clipped to this tile, in tile-local pixel coordinates (empty for a half without an entry)
                const int x0 = max((int)(lo & 0xffff) - tx0, 0), x1 = have ? min((int)(hi & 0xffff) - tx0, TILE) : 0;
                const int y0 = max((int)(lo >> 16) - ty0, 0), y1 = have ? min((int)(hi >> 16) - ty0, TILE) : 0;
                const int wr = x1 - x0;
                const int shift = (wr <= 4) ? 2 : ((wr <= 8) ? 3 : 4);           // patch width 4 / 8 / 16
                const int lx = x0 + (hl & ((1 << shift) - 1));
                const int rows = 16 >> shift;
                const bool col_ok = lx < x1;
                const float dx = bx - r0.x + (float)lx;
                float part[13];
    #pragma unroll
                for (int q = 0; q < 13; ++q) part[q] = 0.0f;
                float2 c01 = make_float2(0.f, 0.f), c23 = c01, c45 = c01;
                for (int ly = y0 + (hl >> shift); ly < y1; ly += rows) {
                    if (col_ok) {
                        const float dy = by - r0.y + (float)ly;
                        const float4 ga = gp_a[ly * TILE + lx], gb = gp_b[ly * TILE + lx];
                        float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                        float g = frb_ex2(power);
                        float amp = g * r1.y;
                        // dL/damp
                        // packed pairs (FFMA2): the six-term dot product and the six colour sums
                        float2 dp = frb_mul2(make_float2(wa.x, wa.y), make_float2(ga.x, ga.y));
                        dp = frb_fma2(make_float2(wa.z, wa.w), make_float2(ga.z, ga.w), dp);
                        dp = frb_fma2(make_float2(wb.x, wb.y), make_float2(gb.x, gb.y), dp);
                        float damp = dp.x + dp.y;
                        if (!ASM) damp += r2.w * gb.z + gb.w;
                        c01 = frb_fma2s(amp, make_float2(ga.x, ga.y), c01);
                        c23 = frb_fma2s(amp, make_float2(ga.z, ga.w), c23);
                        c45 = frb_fma2s(amp, make_float2(gb.x, gb.y), c45);
                        if (!ASM) part[6] = fmaf(amp, gb.z, part[6]);
                        float gd = g * damp;                          // dL/dopacity contribution
                        part[5] += gd;
                        float ty_ = dy * gd;
                        part[1] += ty_;                               // sy
                        part[4] = fmaf(dy, ty_, part[4]);
                    }
                }
                // dx is fixed per lane (a lane stays in its column of the patch): the moments in x follow from
                // the lane's sums of gd and dy * gd
                part[7] = c01.x; part[8] = c01.y; part[9] = c23.x; part[10] = c23.y; part[11] = c45.x; part[12] = c45.y;
                part[0] = dx * part[5];                               // sx
                part[2] = dx * part[0];                               // sum dx^2 gd
                part[3] = dx * part[1];                               // sum dx dy gd
                const float tot = halfwarp_reduce_multi<13>(part, hl);
                const int slot = hl;
                const float sx = __shfl_sync(0xffffffffu, tot, lane & 16), sy = __shfl_sync(0xffffffffu, tot, (lane & 16) | 1);
                // lane k of the half holds the total of value k: thirteen lanes per entry of ONE atomic instruction on
                // two contiguous rows (as four 16-byte vector reductions gathered by shuffles this was 25 % slower)
                if (have && slot < 13) {
                    const float oln2 = r1.y * FRB_LN2;                // dL/d(power) = g * damp * o * ln2
                    float* g2 = grad2d + (size_t)gid * FRB_GRAD_FLOATS;
                    float* gw = gwc + (size_t)gid * WC_FLOATS;
                    float v = tot;
                    if (slot == 0) v = -(2.0f * r0.z * sx + r0.w * sy) * oln2;
                    else if (slot == 1) v = -(r0.w * sx + 2.0f * r1.x * sy) * oln2;
                    else if (slot <= 4) v = tot * oln2;
                    if (slot <= 5 || (slot == 6 && !ASM)) { if (v != 0.0f) atomicAdd(g2 + slot, v); }
                    else if (slot >= 7 && v != 0.0f) atomicAdd(gw + (slot - 7), v);
                }
            }
        } else {
        // Long single-group lists (WaveFieldRenderer): one THREAD per entry walks its rectangle (measured at
        // config 2: 0.50 ms, against 0.60 ms for the warp-per-entry form).
            for (int e = g0 + threadIdx.x; e < g1; e += CTA_THREADS) {
                const float4 r0 = sorted_records[3 * (size_t)e + 0], r1 = sorted_records[3 * (size_t)e + 1],
                             r2 = sorted_records[3 * (size_t)e + 2];
                const float4 wa = sorted_wc[2 * (size_t)e + 0], wb = sorted_wc[2 * (size_t)e + 1];
                const uint32_t lo = __float_as_uint(r1.z), hi = __float_as_uint(r1.w) & 0x7fff7fffu;
                // rectangle clipped to this tile, in tile-local pixel coordinates
                const int x0 = max((int)(lo & 0xffff) - tx0, 0), x1 = min((int)(hi & 0xffff) - tx0, TILE);
                const int y0 = max((int)(lo >> 16) - ty0, 0), y1 = min((int)(hi >> 16) - ty0, TILE);
                const float ux = bx - r0.x, uy = by - r0.y;
                float s_amp = 0.f, sx = 0.f, sy = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f, s_dep = 0.f;
                float dcc0 = 0.f, dcc1 = 0.f, dcc2 = 0.f, dcs0 = 0.f, dcs1 = 0.f, dcs2 = 0.f;
                for (int ly = y0; ly < y1; ++ly) {
                    const float dy = uy + (float)ly;
                    for (int lx = x0; lx < x1; ++lx) {
                        const float dx = ux + (float)lx;
                        const float4 ga = gp_a[ly * TILE + lx], gb = gp_b[ly * TILE + lx];
                        float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                        float g = frb_ex2(power);
                        float amp = g * r1.y;
                        // dL/damp
                        float damp = wa.x * ga.x + wa.y * ga.y + wa.z * ga.z + wa.w * ga.w + wb.x * gb.x + wb.y * gb.y;
                        if (!ASM) damp += r2.w * gb.z + gb.w;
                        dcc0 = fmaf(amp, ga.x, dcc0); dcc1 = fmaf(amp, ga.y, dcc1); dcc2 = fmaf(amp, ga.z, dcc2);
                        dcs0 = fmaf(amp, ga.w, dcs0); dcs1 = fmaf(amp, gb.x, dcs1); dcs2 = fmaf(amp, gb.y, dcs2);
                        if (!ASM) s_dep = fmaf(amp, gb.z, s_dep);
                        float gd = g * damp;                          // dL/dopacity contribution
                        s_amp += gd;
                        float tx_ = dx * gd, ty_ = dy * gd;
                        sx += tx_; sy += ty_;
                        sxx = fmaf(dx, tx_, sxx); sxy = fmaf(dx, ty_, sxy); syy = fmaf(dy, ty_, syy);
                    }
                }
                const float oln2 = r1.y * FRB_LN2;                    // dL/d(power) = g * damp * o * ln2
                const uint32_t gid = sorted_gids[e];
                float* g2 = grad2d + (size_t)gid * FRB_GRAD_FLOATS;
                // four 16-byte vector reductions instead of thirteen scalar atomics
                frb_red_add_f4(g2 + 0, make_float4(-(2.0f * r0.z * sx + r0.w * sy) * oln2,
                                                   -(r0.w * sx + 2.0f * r1.x * sy) * oln2, sxx * oln2, sxy * oln2));
                frb_red_add_f4(g2 + 4, make_float4(syy * oln2, s_amp, ASM ? 0.0f : s_dep, 0.0f));
                float* gw = gwc + (size_t)gid * WC_FLOATS;
                frb_red_add_f4(gw + 0, make_float4(dcc0, dcc1, dcc2, dcs0));
                frb_red_add_f4(gw + 4, make_float4(dcs1, dcs2, 0.0f, 0.0f));
            }
        }
    }
}

// d colour, d phase from d(colour cos phi), d(colour sin phi); writes colours into grad2d slots 8..10.
__global__ void __launch_bounds__(256)
wave_chain_bwd_kernel(int n, const float* __restrict__ colors, const float* __restrict__ phases, int phase_stride,
                      const float* __restrict__ gwc, float* __restrict__ grad2d, float* __restrict__ g_phases) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gph[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ph = phases[(size_t)i * phase_stride + (phase_stride == 3 ? c : 0)];
        float sn, cn;
        sincosf(ph, &sn, &cn);
        float col = colors[3 * (size_t)i + c];
        float dcc = gwc[(size_t)i * WC_FLOATS + c], dcs = gwc[(size_t)i * WC_FLOATS + 3 + c];
        grad2d[(size_t)i * FRB_GRAD_FLOATS + 8 + c] = dcc * cn + dcs * sn;
        gph[c] = col * (dcs * cn - dcc * sn);
    }
    if (phase_stride == 3) {
        g_phases[3 * (size_t)i + 0] = gph[0]; g_phases[3 * (size_t)i + 1] = gph[1]; g_phases[3 * (size_t)i + 2] = gph[2];
    } else {
        g_phases[i] = gph[0] + gph[1] + gph[2];
    }
}

// ---- WaveFieldRenderer epilogue DR:893-924 and its backward ---------------------------------
__device__ __forceinline__ float wave_r(float re, float im) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)), 1e-8f));
}

__global__ void __launch_bounds__(256)
wave_finish_fwd_kernel(int n_views, int width, int height, const float* __restrict__ accum,
                       const uint32_t* __restrict__ rmax_bits, float3 bg, float* __restrict__ image,
                       float* __restrict__ depth) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float* a = accum + (size_t)view * 8 * hw + pix;
    const float mx = fmaxf(__uint_as_float(rmax_bits[view]), 1.0f);         // DR:902
    float re[3] = {a[0], a[hw], a[2 * hw]}, im[3] = {a[3 * hw], a[4 * hw], a[5 * hw]};
    float inten = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) inten += re[c] * re[c] + im[c] * im[c];
    float tot = fminf(fmaxf(sqrtf(inten + 1e-8f), 0.f), 1.f);              // DR:908-909
    const float bgc[3] = {bg.x, bg.y, bg.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float rend = fminf(fmaxf(wave_r(re[c], im[c]) / mx, 0.f), 1.f);     // DR:898-905
        image[(size_t)view * 3 * hw + c * hw + pix] = fminf(fmaxf(rend + bgc[c] * (1.0f - tot), 0.f), 1.f);
    }
    if (depth) depth[t] = a[6 * hw] / (a[7 * hw] + 1e-8f);                   // DR:924
}

// red[view] = { S = sum g_img * gates * r / mx^2 , K = number of elements equal to the maximum }
__global__ void __launch_bounds__(256)
wave_finish_bwd_reduce_kernel(int n_views, int width, int height, const float* __restrict__ accum,
                              const uint32_t* __restrict__ rmax_bits, float3 bg, const float* __restrict__ g_image,
                              float* __restrict__ red) {
    const size_t hw = (size_t)width * height;
    const int view = blockIdx.y;
    const float rmax = __uint_as_float(rmax_bits[view]);
    const float mx = fmaxf(rmax, 1.0f);
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float S = 0.f, K = 0.f;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (size_t)gridDim.x * blockDim.x) {
        const float* a = accum + (size_t)view * 8 * hw + pix;
        float re[3] = {a[0], a[hw], a[2 * hw]}, im[3] = {a[3 * hw], a[4 * hw], a[5 * hw]};
        float inten = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) inten += re[c] * re[c] + im[c] * im[c];
        float tot = fminf(fmaxf(sqrtf(inten + 1e-8f), 0.f), 1.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float r = wave_r(re[c], im[c]);
            float q = r / mx;
            float out = fminf(fmaxf(q, 0.f), 1.f) + bgc[c] * (1.0f - tot);
            float gi = (out >= 0.f && out <= 1.f) ? g_image[(size_t)view * 3 * hw + c * hw + pix] : 0.f;
            if (q >= 0.f && q <= 1.f) S += gi * r / (mx * mx);
            if (r == rmax) K += 1.f;
        }
    }
    S = frb_warp_sum(S);
    K = frb_warp_sum(K);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(red + 2 * view + 0, S);
        atomicAdd(red + 2 * view + 1, K);
    }
}

__global__ void __launch_bounds__(256)
wave_finish_bwd_kernel(int n_views, int width, int height, const float* __restrict__ accum,
                       const uint32_t* __restrict__ rmax_bits, float3 bg, const float* __restrict__ g_image,
                       const float* __restrict__ g_depth, const float* __restrict__ red, float* __restrict__ gpix) {
    const size_t hw = (size_t)width * height;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_views * hw) return;
    const int view = (int)(t / hw);
    const size_t pix = t - (size_t)view * hw;
    const float* a = accum + (size_t)view * 8 * hw + pix;
    const float rmax = __uint_as_float(rmax_bits[view]);
    const float mx = fmaxf(rmax, 1.0f);
    const float mxb = (rmax >= 1.0f) ? -red[2 * view] / fmaxf(red[2 * view + 1], 1.0f) : 0.f;  // to each arg max
    const float bgc[3] = {bg.x, bg.y, bg.z};
    float re[3] = {a[0], a[hw], a[2 * hw]}, im[3] = {a[3 * hw], a[4 * hw], a[5 * hw]};
    float inten = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) inten += re[c] * re[c] + im[c] * im[c];
    float t0 = sqrtf(inten + 1e-8f);
    float tot = fminf(fmaxf(t0, 0.f), 1.f);
    float totb = 0.f, rb[3], rr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float r = wave_r(re[c], im[c]);
        float q = r / mx;
        float out = fminf(fmaxf(q, 0.f), 1.f) + bgc[c] * (1.0f - tot);
        float gi = (out >= 0.f && out <= 1.f) ? g_image[(size_t)view * 3 * hw + c * hw + pix] : 0.f;
        totb -= gi * bgc[c];
        rb[c] = ((q >= 0.f && q <= 1.f) ? gi / mx : 0.f) + ((r == rmax) ? mxb : 0.f);
        rr[c] = r;
    }
    float t0b = (t0 >= 0.f && t0 <= 1.f) ? totb : 0.f;
    float* g = gpix + (size_t)view * 8 * hw + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ib = rb[c] / (2.0f * rr[c]) + t0b / (2.0f * t0);                 // dL/dI_c
        g[c * hw] = 2.0f * re[c] * ib;
        g[(3 + c) * hw] = 2.0f * im[c] * ib;
    }
    float gd = g_depth ? g_depth[t] : 0.f;
    float den = a[7 * hw] + 1e-8f;
    g[6 * hw] = gd / den;
    g[7 * hw] = -gd * a[6 * hw] / (den * den);
}

}  // namespace

static int tiles_of(int width, int height, int* tiles_x) {
    *tiles_x = frb_div_up(width, TILE);
    return *tiles_x * frb_div_up(height, TILE);
}

static int check_views(int n_views, int width, int height) {
    if (n_views < 1 || n_views > FRB_MAX_VIEWS || width < 1 || height < 1) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    return 0;
}

extern "C" int frb_wave_prepare(int n, const float* colors, const float* phases, int phase_stride, float* wc,
                                void* stream) {
    if (n < 0 || (phase_stride != 1 && phase_stride != 3)) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!colors || !phases || !wc) return FRB_E_INVALID;
    wave_prepare_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, colors, phases, phase_stride,
                                                                            (float4*)wc);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_wave_gather(int m, const uint32_t* gids, const float* wc, float* sorted_wc, void* stream) {
    if (m < 0) return FRB_E_INVALID;
    if (m == 0) return 0;
    if (!gids || !wc || !sorted_wc) return FRB_E_INVALID;
    gather_wc_kernel<<<frb_div_up(2ll * m, 256), 256, 0, (cudaStream_t)stream>>>(m, gids, (const float4*)wc,
                                                                               (float4*)sorted_wc);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_wave_splat_fwd(int n_views, int width, int height, const int32_t* ranges,
                                  const float* sorted_records, const float* sorted_wc, float* accum,
                                  uint32_t* rmax_bits, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !accum || !rmax_bits) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x, tpv = tiles_of(width, height, &tiles_x);
    FRB_CUDA_OK(cudaMemsetAsync(rmax_bits, 0, sizeof(uint32_t) * n_views, st));
    wave_splat_fwd_kernel<false><<<n_views * tpv, CTA_THREADS, 0, st>>>(
        width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records, (const float4*)sorted_wc,
        nullptr, 0, accum, rmax_bits, nullptr);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_asm_splat_fwd(int n_views, int width, int height, int n_planes, const int32_t* ranges,
                                 const float* sorted_records, const float* sorted_wc, const uint64_t* keys,
                                 float* fields, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !fields || n_planes < 1 || n_planes > 64) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x, tpv = tiles_of(width, height, &tiles_x);
    // no memset: the kernel writes every (view, plane, channel, pixel) of the fields, zeros included
    wave_splat_fwd_kernel<true><<<n_views * tpv, CTA_THREADS, 0, st>>>(
        width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records, (const float4*)sorted_wc,
        keys, n_planes, nullptr, nullptr, (float2*)fields);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_wave_finish_fwd(int n_views, int width, int height, const float* accum,
                                   const uint32_t* rmax_bits, const float* background_host, float* image,
                                   float* depth, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!accum || !rmax_bits || !background_host || !image) return FRB_E_INVALID;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    long long total = (long long)n_views * width * height;
    wave_finish_fwd_kernel<<<frb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(n_views, width, height, accum,
                                                                                    rmax_bits, bg, image, depth);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_wave_finish_bwd(int n_views, int width, int height, const float* accum,
                                   const uint32_t* rmax_bits, const float* background_host, const float* g_image,
                                   const float* g_depth, float* red, float* gpix, void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!accum || !rmax_bits || !background_host || !g_image || !red || !gpix) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    long long hw = (long long)width * height;
    FRB_CUDA_OK(cudaMemsetAsync(red, 0, sizeof(float) * 2 * n_views, st));
    dim3 grid(min(frb_div_up(hw, 256), 592), n_views);
    wave_finish_bwd_reduce_kernel<<<grid, 256, 0, st>>>(n_views, width, height, accum, rmax_bits, bg, g_image, red);
    wave_finish_bwd_kernel<<<frb_div_up(hw * n_views, 256), 256, 0, st>>>(n_views, width, height, accum, rmax_bits,
                                                                          bg, g_image, g_depth, red, gpix);
    frb_note_launches(2);
    FRB_LAUNCH_CHECK();
    return 0;
}

// ASM variant when n_planes > 0: gpix = d(fields) [view][plane][3][H][W] float2, scaled by `scale`.
extern "C" int frb_wave_splat_bwd(int n_views, int width, int height, int n_planes, const int32_t* ranges,
                                  const float* sorted_records, const float* sorted_wc, const uint32_t* sorted_gids,
                                  const uint64_t* keys, const float* gpix, float scale, float* grad2d, float* gwc,
                                  void* stream) {
    int rc = check_views(n_views, width, height);
    if (rc) return rc;
    if (!ranges || !gpix || !grad2d || !gwc || n_planes < 0 || n_planes > 64) return FRB_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x, tpv = tiles_of(width, height, &tiles_x);
    if (n_planes == 0)
        wave_splat_bwd_kernel<false><<<n_views * tpv, CTA_THREADS, 0, st>>>(
            width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records,
            (const float4*)sorted_wc, sorted_gids, nullptr, 0, gpix, 1.0f, grad2d, gwc);
    else
        wave_splat_bwd_kernel<true><<<n_views * tpv, CTA_THREADS, 0, st>>>(
            width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records,
            (const float4*)sorted_wc, sorted_gids, keys, n_planes, gpix, scale, grad2d, gwc);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_wave_chain_bwd(int n, const float* colors, const float* phases, int phase_stride,
                                  const float* gwc, float* grad2d, float* g_phases, void* stream) {
    if (n < 0 || (phase_stride != 1 && phase_stride != 3)) return FRB_E_INVALID;
    if (n == 0) return 0;
    if (!colors || !phases || !gwc || !grad2d || !g_phases) return FRB_E_INVALID;
    wave_chain_bwd_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, colors, phases, phase_stride, gwc,
                                                                              grad2d, g_phases);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
