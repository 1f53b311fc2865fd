// Fresnel phase-blending compositor (use_phase_blending=True), forward and backward.
// Reference: TileBasedRenderer.forward DR:571-575 (state), DR:629-645 (interference modulation of
// alpha), DR:660-667 (running weighted-average phase), SURVEY.md appendix A.3 / A.5.
//
// Per pixel and record, in depth order, with state (T = 1 - accumulated_alpha, Phi):
//     a0    = g * opacity
//     d0    = |phi - Phi| ;  d = min(d0, 1 - d0)
//     m     = (1 - A) + A * cos(d * 2 * 3.14159)            (the reference's literal)
//     alpha = clamp(a0 * m, 0, 0.99) ;  c = alpha * T
//     colour += c * rgb ; depth += c * z ; T' = T - c
//     pc    = c / max(1 - T', 1e-6) ;  Phi' = Phi * (1 - pc) + phi * pc
//
// The reference cannot backpropagate through this path (in-place slice updates, SURVEY.md note 3);
// the gradient definition is the autograd of the same expressions (oracle: .clone() restatement).
// The Phi recurrence has no stable inverse (pc = 1 on the first hit), so the forward pass
// checkpoints (T, Phi) per pixel every 32 list entries and the backward pass recomputes each
// 32-entry block forward from its checkpoint (keeping Phi_j in registers) before walking it back.
#include "composite_common.cuh"

namespace {

constexpr int SUB = 32;                          // checkpoint interval (list entries)
constexpr float TWO_PI_REF = 2.0f * 3.14159f;    // DR:642 uses the literal 3.14159
constexpr int N_PHASE_GRADS = 11;

__device__ __forceinline__ size_t ckpt_slot(int range_x, int tile, int sub_block) {
    return (size_t)(range_x >> 5) + (size_t)tile + (size_t)sub_block;
}

struct PhaseStep {
    float a0, m, d, d0, a1, alpha, c, accn, den, pc, Phin;
};

// One forward step of the recurrence (shared by the forward kernel and the backward recompute).
// The state is the reference's accumulated_alpha itself (sum form, DR:650-658): pc divides by it, and
// 1 - (1 - c) would lose the low bits of a small first contribution.
__device__ __forceinline__ void phase_step(float g, float o, float phi, float A, float acc, float Phi,
                                           PhaseStep& s) {
    s.a0 = g * o;
    s.d0 = fabsf(phi - Phi);
    s.d = fminf(s.d0, 1.0f - s.d0);
    s.m = (1.0f - A) + A * __cosf(s.d * TWO_PI_REF);
    s.a1 = s.a0 * s.m;
    s.alpha = fminf(fmaxf(s.a1, 0.0f), FRB_ALPHA_MAX);
    s.c = s.alpha * (1.0f - acc);
    s.accn = acc + s.c;
    s.den = fmaxf(s.accn, 1e-6f);
    s.pc = __fdividef(s.c, s.den);
    s.Phin = Phi * (1.0f - s.pc) + phi * s.pc;
}

__global__ void __launch_bounds__(CTA_THREADS)
composite_phase_fwd_kernel(int width, int height, int tiles_x, int tiles_per_view,
                           const int* __restrict__ tile_order, const int2* __restrict__ ranges,
                           const float4* __restrict__ sorted_records, const float* __restrict__ sorted_phases,
                           float A, float3 bg, float t_eps, float* __restrict__ image,
                           float* __restrict__ depth_out, float* __restrict__ alpha_out,
                           float* __restrict__ state_T, int* __restrict__ state_n, float2* __restrict__ ckpt) {
    __shared__ StageBuf stage[STAGES];
    __shared__ float phase_s[STAGES][BATCH];
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;   // heaviest tiles first
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + foot_x(threadIdx.x);
    const int py = ty * TILE + foot_y(threadIdx.x);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;

    const int2 range = ranges[tile];
    const int count = range.y - range.x;
    const int n_batches = (count + BATCH - 1) / BATCH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) frb_mbar_init(&full_bar[s], 1);
        frb_mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int b) {
        int s = b % STAGES;
        int cnt = min(BATCH, count - b * BATCH);
        frb_mbar_expect_tx(&full_bar[s], cnt * RECORD_BYTES);
        frb_tma_load_1d(stage[s].rec, sorted_records + 3 * (size_t)(range.x + b * BATCH), cnt * RECORD_BYTES,
                        &full_bar[s]);
    };
    if (threadIdx.x == 0)
        for (int b = 0; b < STAGES && b < n_batches; ++b) issue(b);

    float acc = 0.0f, Phi = 0.0f;
    float2 c_rg = make_float2(0.0f, 0.0f), c_bd = c_rg;        // (r, g) and (b, depth) sums as fp32 pairs (FFMA2)
    int consumed = count;
    bool done = !in_image;
    // the warp's 8x4 pixel block; the pixel's words for the rectangle test move to a pixel that is in no rectangle
    // (32766, 32766) once the pixel is done or when it lies outside the image
    constexpr uint32_t NO_PIXEL_GUARD = 0xfffefffeu, NO_PIXEL_PLUS1 = 0x7fff7fffu;
    const int wx0 = tx * TILE + ((threadIdx.x >> 5) % (TILE / FOOT_W)) * FOOT_W, wx1 = wx0 + FOOT_W;
    const int wy0 = ty * TILE + ((threadIdx.x >> 5) / (TILE / FOOT_W)) * FOOT_H, wy1 = wy0 + FOOT_H;
    uint32_t pxy_guard_v = in_image ? pxy_guard : NO_PIXEL_GUARD, pxy_plus1_v = in_image ? pxy_plus1 : NO_PIXEL_PLUS1;
    const float stop = fmaxf(t_eps, T_FLOOR);

    for (int b = 0; b < n_batches; ++b) {
        const int s = b % STAGES;
        const int cnt = min(BATCH, count - b * BATCH);
        if (threadIdx.x < cnt) phase_s[s][threadIdx.x] = sorted_phases[range.x + b * BATCH + threadIdx.x];
        frb_mbar_wait(&full_bar[s], (b / STAGES) & 1);
        __syncthreads();
        // Per 32-entry block (= checkpoint interval): termination test and checkpoint, then the warp walks the records
        // whose rectangle touches its 8x4 block, four per iteration - the state-independent part (loads, rectangle
        // test, exp2, phase fetch) of the four first, then the four serial steps; a lane outside a rectangle carries
        // a0 = 0 through its step (alpha = c = pc = 0: neither acc nor Phi changes a bit).  Same arithmetic per pixel
        // as the record-by-record loop this replaces (0.226 ms at configs[3]).
        {
            const float4* rec = stage[s].rec;
            const int lane = threadIdx.x & 31;
            for (int sb = 0; sb * SUB < cnt; ++sb) {
                if (!done) {
                    if (1.0f - acc < stop) {              // tested at checkpoint boundaries
                        done = true;
                        consumed = b * BATCH + sb * SUB;
                        pxy_guard_v = NO_PIXEL_GUARD;     // in no rectangle from here on
                        pxy_plus1_v = NO_PIXEL_PLUS1;
                    } else if (ckpt) {
                        ckpt[ckpt_slot(range.x, tile, (b * BATCH + sb * SUB) >> 5) * CTA_THREADS + threadIdx.x] =
                            make_float2(acc, Phi);
                    }
                }
                if (!__any_sync(0xffffffffu, !done)) break;
                const int sub_cnt = min(SUB, cnt - sb * SUB);
                uint32_t cand;
                {
                    bool ok = false;
                    if (lane < sub_cnt) {
                        const float4 q = rec[3 * (sb * SUB + lane) + 1];
                        const uint32_t lo = __float_as_uint(q.z), hi = __float_as_uint(q.w) & 0x7fff7fffu;
                        ok = (int)(lo & 0xffff) < wx1 && (int)(hi & 0xffff) > wx0 && (int)(lo >> 16) < wy1 &&
                             (int)(hi >> 16) > wy0;
                    }
                    cand = __brev(__ballot_sync(0xffffffffu, ok));          // bit 31 - j: front to back by bfind
                }
                const char* rec_sb = reinterpret_cast<const char*>(rec + 3 * (sb * SUB));
                const float* phase_sb = &phase_s[s][sb * SUB];
                struct Pre { float g, o, phi; float2 rg, bd; };
                auto pre = [&](int j) {
                    Pre p;
                    const float4* rj = reinterpret_cast<const float4*>(rec_sb + j * 48);
                    const float4 r1 = rj[1], r0 = rj[0], r2 = rj[2];
                    const bool in = rect_contains(pxy_guard_v, pxy_plus1_v, __float_as_uint(r1.z), __float_as_uint(r1.w));
                    const float dx = fpx - r0.x, dy = fpy - r0.y;
                    const float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                    p.g = in ? frb_ex2(power) : 0.0f;
                    p.o = r1.y;
                    p.phi = phase_sb[j];
                    p.rg = make_float2(r2.x, r2.y);
                    p.bd = make_float2(r2.z, r2.w);
                    return p;
                };
                auto step = [&](const Pre& p) {
                    PhaseStep st;
                    phase_step(p.g, p.o, p.phi, A, acc, Phi, st);
                    c_rg = frb_fma2s(st.c, p.rg, c_rg);
                    c_bd = frb_fma2s(st.c, p.bd, c_bd);
                    acc = st.accn;
                    Phi = st.Phin;
                };
                while (cand) {
                    if (__popc(cand) >= 4) {
                        Pre pp[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int bpos = frb_bfind(cand);
                            cand ^= 1u << bpos;
                            pp[q] = pre(31 - bpos);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) step(pp[q]);
                    } else {
                        const int bpos = frb_bfind(cand);
                        cand ^= 1u << bpos;
                        const Pre p0 = pre(31 - bpos);
                        step(p0);
                    }
                }
            }
        }
        int all_done = __syncthreads_and(done ? 1 : 0);
        if (all_done) {
            for (int bb = b + 1; bb < n_batches && bb < b + STAGES; ++bb)
                frb_mbar_wait(&full_bar[bb % STAGES], (bb / STAGES) & 1);
            break;
        }
        if (threadIdx.x == 0 && b + STAGES < n_batches) issue(b + STAGES);
    }

    if (in_image) {
        const size_t hw = (size_t)width * height;
        const size_t pix = (size_t)view * hw + (size_t)py * width + px;
        float* img = image + (size_t)view * 3 * hw + (size_t)py * width + px;
        const float T = 1.0f - acc;
        const float cr = c_rg.x, cg = c_rg.y, cb = c_bd.x, cd = c_bd.y;
        float o0 = fmaf(T, bg.x, cr), o1 = fmaf(T, bg.y, cg), o2 = fmaf(T, bg.z, cb);
        img[0] = fminf(fmaxf(o0, 0.0f), 1.0f);
        img[hw] = fminf(fmaxf(o1, 0.0f), 1.0f);
        img[2 * hw] = fminf(fmaxf(o2, 0.0f), 1.0f);
        depth_out[pix] = cd;
        alpha_out[pix] = acc;
        state_T[pix] = T;
        int gates = ((o0 >= 0.0f && o0 <= 1.0f) ? 1 : 0) | ((o1 >= 0.0f && o1 <= 1.0f) ? 2 : 0) |
                    ((o2 >= 0.0f && o2 <= 1.0f) ? 4 : 0);
        state_n[pix] = consumed | (gates << STATE_GATE_SHIFT);
    }
}

// Backward.  The tile list is visited in 32-entry blocks (= checkpoint interval) from last to first.
// Per block every warp (32 pixels) runs three phases, the same decomposition as composite_bwd_kernel:
//   phase 1a (lane = pixel): recompute the block forward from its checkpoint over the warp's candidate
//            records, leaving the state (acc, Phi) BEFORE every entry in a 32x32 shared-memory tile;
//   phase 1b (lane = pixel): walk the candidates back, reverse mode through the step with the adjoints
//            (Tb = dL/dT, Pb = dL/dPhi) carried per pixel; per pair the tile slot is overwritten with
//            (c, g * dL/da0) and a second tile receives dL/dphi;
//   phase 2  (lane = Gaussian): read both tiles transposed, walk the 32 pixels and keep the eleven sums of
//            the lane's Gaussian in registers.
// The eight warps' sums meet in shared memory (shared atomics) and leave the CTA as ONE global atomic per
// Gaussian value per tile.
constexpr int PH_STAGES = 4;
constexpr int PH_STRIDE = 33;                    // row stride of the per-warp tiles: conflict-free both ways
constexpr int PH_WARPS = CTA_THREADS / 32;

struct PhaseBwdSmem {
    float4 rec[PH_STAGES][SUB * 3];
    float2 pair[PH_WARPS][32 * PH_STRIDE];
    float phib[PH_WARPS][32 * PH_STRIDE];
    float4 pixc[PH_WARPS][32];
    float sums[2][N_PHASE_GRADS][SUB];
    uint64_t full_bar[PH_STAGES];
    int max_n;
};

__global__ void __launch_bounds__(CTA_THREADS, 2)
composite_phase_bwd_kernel(int width, int height, int tiles_x, int tiles_per_view,
                           const int* __restrict__ tile_order, const int2* __restrict__ ranges,
                           const float4* __restrict__ sorted_records, const uint32_t* __restrict__ sorted_gids,
                           const float* __restrict__ sorted_phases, float A, float3 bg,
                           const float* __restrict__ state_T, const int* __restrict__ state_n,
                           const float2* __restrict__ ckpt, const float* __restrict__ g_image,
                           const float* __restrict__ g_depth, const float* __restrict__ g_alpha,
                           float* __restrict__ grad2d, float* __restrict__ g_phases) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseBwdSmem& sm = *reinterpret_cast<PhaseBwdSmem*>(smem_raw);

    const int tile = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;   // heaviest tiles first
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + foot_x(threadIdx.x);
    const int py = ty * TILE + foot_y(threadIdx.x);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wx0 = tx * TILE + (warp % (TILE / FOOT_W)) * FOOT_W, wy0 = ty * TILE + (warp / (TILE / FOOT_W)) * FOOT_H;
    const int wx1 = wx0 + FOOT_W, wy1 = wy0 + FOOT_H;
    const float wbx = (float)wx0, wby = (float)wy0;                              // warp's pixel block origin
    const int2 range = ranges[tile];

    float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
    int my_n = 0;
    if (in_image) {
        const size_t hw = (size_t)width * height;
        const size_t pix = (size_t)view * hw + (size_t)py * width + px;
        const size_t ip = (size_t)view * 3 * hw + (size_t)py * width + px;
        const int st = state_n[pix];
        my_n = st & STATE_N_MASK;
        const int gates = st >> STATE_GATE_SHIFT;
        gr = (gates & 1) ? g_image[ip] : 0.0f;
        gg = (gates & 2) ? g_image[ip + hw] : 0.0f;
        gb = (gates & 4) ? g_image[ip + 2 * hw] : 0.0f;
        gd = g_depth ? g_depth[pix] : 0.0f;
        ga = g_alpha ? g_alpha[pix] : 0.0f;
    }
    sm.pixc[warp][lane] = make_float4(gr, gg, gb, gd);
    for (int i = threadIdx.x; i < 2 * N_PHASE_GRADS * SUB; i += CTA_THREADS) (&sm.sums[0][0][0])[i] = 0.0f;
    if (threadIdx.x == 0) {
        sm.max_n = 0;
        for (int s = 0; s < PH_STAGES; ++s) frb_mbar_init(&sm.full_bar[s], 1);
        frb_mbar_fence_init();
    }
    __syncthreads();
    int warp_n = my_n;                           // entries applied by at least one pixel of the warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_n = max(warp_n, __shfl_xor_sync(0xffffffffu, warp_n, o));
    if (lane == 0) atomicMax(&sm.max_n, warp_n);
    __syncthreads();
    const int count = min(sm.max_n, range.y - range.x);
    const int n_blocks = (count + SUB - 1) / SUB;
    if (n_blocks == 0) return;

    auto issue = [&](int visit) {
        int k = n_blocks - 1 - visit;
        int s = visit % PH_STAGES;
        int cnt = min(SUB, count - k * SUB);
        frb_mbar_expect_tx(&sm.full_bar[s], cnt * RECORD_BYTES);
        frb_tma_load_1d(sm.rec[s], sorted_records + 3 * (size_t)(range.x + k * SUB), cnt * RECORD_BYTES,
                        &sm.full_bar[s]);
    };
    if (threadIdx.x == 0)
        for (int v = 0; v < PH_STAGES && v < n_blocks; ++v) issue(v);

    float Tb = gr * bg.x + gg * bg.y + gb * bg.z - ga;     // dL/dT_final
    float Pb = 0.0f;                                        // dL/dPhi_final
    float2* my_pair = sm.pair[warp];
    float* my_phib = sm.phib[warp];

    for (int visit = 0; visit < n_blocks; ++visit) {
        const int k = n_blocks - 1 - visit;
        const int s = visit % PH_STAGES;
        const int cnt = min(SUB, count - k * SUB);
        const int base_n = k * SUB;
        // lane l keeps the phase of record l of the block (read by shuffle: uniform record index)
        const float phi_l = (lane < cnt) ? sorted_phases[range.x + base_n + lane] : 0.0f;
        frb_mbar_wait(&sm.full_bar[s], (visit / PH_STAGES) & 1);
        const float4* rec = sm.rec[s];

        // ---- candidates: lane l tests record l against this warp's pixel block ----
        uint32_t cand;
        {
            bool ok = false;
            if (lane < cnt) {
                const uint32_t lo = __float_as_uint(rec[3 * lane + 1].z);
                const uint32_t hi = __float_as_uint(rec[3 * lane + 1].w) & 0x7fff7fffu;
                ok = (int)(lo & 0xffff) < wx1 && (int)(hi & 0xffff) > wx0 && (int)(lo >> 16) < wy1 &&
                     (int)(hi >> 16) > wy0;
            }
            cand = __ballot_sync(0xffffffffu, ok);
            const int wn = warp_n - base_n;          // entries >= wn were applied by no pixel of the warp
            if (wn < 32) cand &= (wn <= 0) ? 0u : ((1u << wn) - 1u);
        }
        const uint32_t gmask = cand;
        const int local_n = my_n - base_n;           // entries j < local_n were applied by this pixel

        if (gmask != 0) {
            // Both walks handle TWO candidates per iteration: everything that does not depend on the running state is
            // evaluated for both first, then the two serial chains run back to back.  A warp issues in order, so
            // without this the ~25-instruction dependent chain of one candidate (|phi - Phi| -> cos -> clamp -> c ->
            // 1 / acc -> Phi' forward; the adjoints Tb, Pb backward) stalls the 70 independent instructions of the
            // next one behind it (ncu, profiles/r2_g_*: issue slots 57 % busy, "wait" the top stall at 16 warps / SM).
            // ---- phase 1a: forward recompute from the checkpoint ----
            float acc = 0.0f, Phi = 0.0f;
            if (local_n > 0) {
                const float2 c0 = ckpt[ckpt_slot(range.x, tile, k) * CTA_THREADS + threadIdx.x];
                acc = c0.x; Phi = c0.y;
            }
            struct FwdPre { float a0, phi; bool active; };
            auto fwd_pre = [&](int j) {
                FwdPre p;
                p.phi = __shfl_sync(0xffffffffu, phi_l, j);
                const float4 r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                p.active = (j < local_n) &&
                           rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.z), __float_as_uint(r1.w));
                const float4 r0 = rec[3 * j + 0];
                const float dx = fpx - r0.x, dy = fpy - r0.y;
                const float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                p.a0 = frb_ex2(power) * r1.y;
                return p;
            };
            auto fwd_step = [&](int j, const FwdPre& p) {
                my_pair[j * PH_STRIDE + lane] = make_float2(acc, Phi);          // the state BEFORE entry j
                if (p.active) {
                    // phase_step with a0 = g * o already formed
                    const float d0 = fabsf(p.phi - Phi);
                    const float d = fminf(d0, 1.0f - d0);
                    const float mm = (1.0f - A) + A * __cosf(d * TWO_PI_REF);
                    const float alpha = fminf(fmaxf(p.a0 * mm, 0.0f), FRB_ALPHA_MAX);
                    const float c = alpha * (1.0f - acc);
                    acc = acc + c;
                    const float pc = __fdividef(c, fmaxf(acc, 1e-6f));
                    Phi = Phi * (1.0f - pc) + p.phi * pc;
                }
            };
            uint32_t m = gmask;
            while (__popc(m) >= 4) {                    // four candidates per iteration while at least four remain
                int jj[4];
                FwdPre pp[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    jj[q] = __ffs(m) - 1;
                    m &= m - 1;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) pp[q] = fwd_pre(jj[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q) fwd_step(jj[q], pp[q]);
            }
            while (m) {
                const int j0 = __ffs(m) - 1;
                m &= m - 1;
                if (m) {
                    const int j1 = __ffs(m) - 1;
                    m &= m - 1;
                    const FwdPre p0 = fwd_pre(j0), p1 = fwd_pre(j1);
                    fwd_step(j0, p0);
                    fwd_step(j1, p1);
                } else {
                    const FwdPre p0 = fwd_pre(j0);
                    fwd_step(j0, p0);
                }
            }
            // ---- phase 1b: walk back ----
            // State-independent part of entry j (replayed from the remembered state before it), reduced to the
            // coefficients the adjoint chain needs:
            //   pcb = Pb K1;  Tb_tot = Tb + pcb q;  cb = pcb iden + w - Tb_tot;  Tb' = Tb_tot + cb alpha;
            //   a1b = gate ? cb T0 : 0;  t = a1b E;  dL/dphi = Pb pc + t;  Pb' = Pb (1 - pc) - t;  dL/da0 = a1b m
            struct BwdPre { float K1, pc, iden, q, w, alpha, T0, mm, E, c, g; bool active, gate; };
            auto bwd_pre = [&](int j) {
                BwdPre p;
                const float phi = __shfl_sync(0xffffffffu, phi_l, j);
                const float2 before = my_pair[j * PH_STRIDE + lane];
                const float4 r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                p.active = (j < local_n) &&
                           rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.z), __float_as_uint(r1.w));
                const float4 r0 = rec[3 * j + 0];
                const float Phi0 = before.y;
                const float dx = fpx - r0.x, dy = fpy - r0.y;
                const float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                p.g = frb_ex2(power);
                PhaseStep st;
                phase_step(p.g, r1.y, phi, A, before.x, Phi0, st);           // replay from the remembered state
                const float sn = __sinf(st.d * TWO_PI_REF);
                p.T0 = 1.0f - before.x;
                p.iden = frb_rcp(st.den);
                p.w = gr * r2.x + gg * r2.y + gb * r2.z + gd * r2.w;
                p.K1 = phi - Phi0;
                p.pc = st.pc;
                p.q = (st.accn >= 1e-6f) ? st.c * p.iden * p.iden : 0.0f;
                p.alpha = st.alpha;
                p.gate = (st.alpha == st.a1);                                  // clamp gate (inclusive)
                p.mm = st.m;
                // d0b * sg with db = -mb A 2pi sn, mb = a1b a0: everything but a1b
                const float s1 = (st.d0 < 1.0f - st.d0) ? 1.0f : ((st.d0 > 1.0f - st.d0) ? -1.0f : 0.0f);
                const float sg = (phi > Phi0) ? 1.0f : ((phi < Phi0) ? -1.0f : 0.0f);
                p.E = -st.a0 * A * TWO_PI_REF * sn * s1 * sg;
                p.c = st.c;
                return p;
            };
            auto bwd_step = [&](int j, const BwdPre& p) {
                float2 out = make_float2(0.f, 0.f);
                float phib = 0.0f;
                if (p.active) {
                    const float pcb = Pb * p.K1;
                    const float Tb_tot = fmaf(pcb, p.q, Tb);
                    const float cb_ = fmaf(pcb, p.iden, p.w) - Tb_tot;
                    const float a1b = p.gate ? cb_ * p.T0 : 0.0f;
                    const float t = a1b * p.E;
                    phib = fmaf(Pb, p.pc, t);
                    Pb = fmaf(Pb, 1.0f - p.pc, -t);
                    Tb = fmaf(cb_, p.alpha, Tb_tot);
                    out.x = p.c;
                    out.y = p.g * (a1b * p.mm);                                 // dL/dopacity part
                }
                my_pair[j * PH_STRIDE + lane] = out;
                my_phib[j * PH_STRIDE + lane] = phib;
            };
            m = gmask;
            while (__popc(m) >= 4) {
                int jj[4];
                BwdPre pp[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    jj[q] = frb_bfind(m);
                    m ^= 1u << jj[q];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) pp[q] = bwd_pre(jj[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q) bwd_step(jj[q], pp[q]);
            }
            while (m) {
                const int j0 = frb_bfind(m);
                m ^= 1u << j0;
                if (m) {
                    const int j1 = frb_bfind(m);
                    m ^= 1u << j1;
                    const BwdPre p0 = bwd_pre(j0), p1 = bwd_pre(j1);
                    bwd_step(j0, p0);
                    bwd_step(j1, p1);
                } else {
                    const BwdPre p0 = bwd_pre(j0);
                    bwd_step(j0, p0);
                }
            }
            __syncwarp();
            // ---- phase 2: lane = Gaussian ----
            if ((gmask >> lane) & 1u) {
                const float4 r0 = rec[3 * lane + 0], r1 = rec[3 * lane + 1];
                const float oln2 = r1.y * FRB_LN2;
                const float ux = wbx - r0.x, uy = wby - r0.y;
                const float2* row = my_pair + lane * PH_STRIDE;
                const float* prow = my_phib + lane * PH_STRIDE;
                const float4* pc = sm.pixc[warp];
                float d_phi = 0.f;
                // colour / depth sums as fp32 pairs (FFMA2); the geometric sums as the six moments sum dpow i^a k^b over
                // the pixel's place (i, k) in the 8x4 block (immediate operands), combined once per Gaussian - see
                // composite.cu phase 2
                float2 s_rg = make_float2(0.f, 0.f), s_bd = s_rg;
                float m00 = 0.f, m10 = 0.f, m01 = 0.f, m20 = 0.f, m11 = 0.f, m02 = 0.f;
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    const float2 cd = row[p];
                    const float4 gpix = pc[p];
                    const int i = p % FOOT_W, k = p / FOOT_W;
                    s_rg = frb_fma2s(cd.x, make_float2(gpix.x, gpix.y), s_rg);
                    s_bd = frb_fma2s(cd.x, make_float2(gpix.z, gpix.w), s_bd);
                    d_phi += prow[p];
                    m00 += cd.y;
                    if (i) m10 = fmaf(cd.y, (float)i, m10);
                    if (k) m01 = fmaf(cd.y, (float)k, m01);
                    if (i) m20 = fmaf(cd.y, (float)(i * i), m20);
                    if (i && k) m11 = fmaf(cd.y, (float)(i * k), m11);
                    if (k) m02 = fmaf(cd.y, (float)(k * k), m02);
                }
                const float d_o = m00;
                const float sx = fmaf(ux, m00, m10), sy = fmaf(uy, m00, m01);
                const float d_A = fmaf(ux, fmaf(ux, m00, 2.0f * m10), m20);
                const float d_B = fmaf(ux, fmaf(uy, m00, m01), fmaf(uy, m10, m11));
                const float d_C = fmaf(uy, fmaf(uy, m00, 2.0f * m01), m02);
                const float d_r = s_rg.x, d_g = s_rg.y, d_b = s_bd.x, d_dep = s_bd.y;
                float* sp = &sm.sums[visit & 1][0][lane];
                atomicAdd(sp + 0 * SUB, -(2.0f * r0.z * sx + r0.w * sy) * oln2);
                atomicAdd(sp + 1 * SUB, -(r0.w * sx + 2.0f * r1.x * sy) * oln2);
                atomicAdd(sp + 2 * SUB, d_A * oln2);
                atomicAdd(sp + 3 * SUB, d_B * oln2);
                atomicAdd(sp + 4 * SUB, d_C * oln2);
                atomicAdd(sp + 5 * SUB, d_o);
                atomicAdd(sp + 6 * SUB, d_dep);
                atomicAdd(sp + 7 * SUB, d_r);
                atomicAdd(sp + 8 * SUB, d_g);
                atomicAdd(sp + 9 * SUB, d_b);
                atomicAdd(sp + 10 * SUB, d_phi);
            }
        }
        __syncthreads();    // every warp's sums of this block are in; stage s is free
        if (threadIdx.x == 0 && visit + PH_STAGES < n_blocks) issue(visit + PH_STAGES);
        // one thread per (Gaussian, float4 of its grad2d row [du dv dA dB | dC do ddepth _ | dr dg db _]) leaves with a
        // 16-byte vector reduction; a fourth group of threads takes the phase gradients
        if (threadIdx.x < 4 * SUB) {
            const int q = threadIdx.x / SUB, jb = threadIdx.x - q * SUB;
            float* base = &sm.sums[visit & 1][0][jb];
            if (q < 3) {
                const int v0 = (q == 0) ? 0 : (q == 1 ? 4 : 7);
                float4 sum = make_float4(base[v0 * SUB], base[(v0 + 1) * SUB], base[(v0 + 2) * SUB],
                                         q == 0 ? base[3 * SUB] : 0.0f);
                if (sum.x != 0.0f || sum.y != 0.0f || sum.z != 0.0f || sum.w != 0.0f) {
                    base[v0 * SUB] = 0.0f; base[(v0 + 1) * SUB] = 0.0f; base[(v0 + 2) * SUB] = 0.0f;
                    if (q == 0) base[3 * SUB] = 0.0f;        // this buffer is used again two blocks later
                    const uint32_t gid = sorted_gids[range.x + base_n + jb];
                    frb_red_add_f4(grad2d + (size_t)gid * FRB_GRAD_FLOATS + 4 * q, sum);
                }
            } else {
                const float sum = base[10 * SUB];
                if (sum != 0.0f) {
                    base[10 * SUB] = 0.0f;
                    atomicAdd(g_phases + sorted_gids[range.x + base_n + jb], sum);
                }
            }
        }
    }
}

}  // namespace

extern "C" size_t frb_phase_ckpt_floats(int m, int n_tiles) {
    if (m < 0) m = 0;
    if (n_tiles < 0) n_tiles = 0;
    return ((size_t)(m >> 5) + (size_t)n_tiles + 2) * CTA_THREADS * 2;
}

int frb_composite_phase_fwd_launch(int n_views, int width, int height, const int32_t* tile_order,
                                   const int32_t* ranges,
                                   const float* sorted_records, const float* sorted_phases, float phase_amplitude,
                                   const float* background_host, float t_eps, float* image, float* depth,
                                   float* alpha, float* state_T, int32_t* state_n, float* ckpt, cudaStream_t st) {
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    composite_phase_fwd_kernel<<<n_views * tpv, CTA_THREADS, 0, st>>>(
        width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)sorted_records, sorted_phases,
        phase_amplitude, bg, t_eps, image, depth, alpha, state_T, state_n, (float2*)ckpt);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

int frb_composite_phase_bwd_launch(int n_views, int width, int height, const int32_t* tile_order,
                                   const int32_t* ranges,
                                   const float* sorted_records, const uint32_t* sorted_gids,
                                   const float* sorted_phases, float phase_amplitude, const float* background_host,
                                   const float* state_T, const int32_t* state_n, const float* ckpt,
                                   const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                                   float* g_phases, cudaStream_t st) {
    if (!ckpt || !g_phases) return FRB_E_INVALID;
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    static unsigned long long smem_opted_in = 0;          // per-device bitmask
    FRB_CUDA_OK(frb_opt_in_smem(composite_phase_bwd_kernel, (int)sizeof(PhaseBwdSmem), &smem_opted_in));
    composite_phase_bwd_kernel<<<n_views * tpv, CTA_THREADS, sizeof(PhaseBwdSmem), st>>>(
        width, height, tiles_x, tpv, tile_order, (const int2*)ranges, (const float4*)sorted_records, sorted_gids,
        sorted_phases, phase_amplitude, bg, state_T, state_n, (const float2*)ckpt, g_image, g_depth, g_alpha, grad2d,
        g_phases);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
