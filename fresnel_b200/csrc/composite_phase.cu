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
composite_phase_fwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int2* __restrict__ ranges,
                           const float4* __restrict__ sorted_records, const float* __restrict__ sorted_phases,
                           float A, float3 bg, float t_eps, float* __restrict__ image,
                           float* __restrict__ depth_out, float* __restrict__ alpha_out,
                           float* __restrict__ state_T, int* __restrict__ state_n, float2* __restrict__ ckpt) {
    __shared__ StageBuf stage[STAGES];
    __shared__ float phase_s[STAGES][BATCH];
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int tile = blockIdx.x;
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + (threadIdx.x & (TILE - 1));
    const int py = ty * TILE + (threadIdx.x / TILE);
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

    float acc = 0.0f, Phi = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, cd = 0.0f;
    int consumed = count;
    bool done = !in_image;
    const float stop = fmaxf(t_eps, T_FLOOR);

    for (int b = 0; b < n_batches; ++b) {
        const int s = b % STAGES;
        const int cnt = min(BATCH, count - b * BATCH);
        if (threadIdx.x < cnt) phase_s[s][threadIdx.x] = sorted_phases[range.x + b * BATCH + threadIdx.x];
        frb_mbar_wait(&full_bar[s], (b / STAGES) & 1);
        __syncthreads();
        if (!done) {
            const float4* rec = stage[s].rec;
            for (int j = 0; j < cnt; ++j) {
                if ((j & (SUB - 1)) == 0) {
                    if (1.0f - acc < stop) {              // tested at checkpoint boundaries
                        done = true;
                        consumed = b * BATCH + j;
                        break;
                    }
                    if (ckpt) ckpt[ckpt_slot(range.x, tile, (b * BATCH + j) >> 5) * CTA_THREADS + threadIdx.x] =
                        make_float2(acc, Phi);
                }
                float4 r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                if (rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.w), __float_as_uint(r2.w))) {
                    float4 r0 = rec[3 * j + 0];
                    float dx = fpx - r0.x, dy = fpy - r0.y;
                    float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                    PhaseStep st;
                    phase_step(frb_ex2(power), r1.y, phase_s[s][j], A, acc, Phi, st);
                    cr = fmaf(st.c, r2.x, cr);
                    cg = fmaf(st.c, r2.y, cg);
                    cb = fmaf(st.c, r2.z, cb);
                    cd = fmaf(st.c, r1.z, cd);
                    acc = st.accn;
                    Phi = st.Phin;
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

// Backward: per pixel thread, tile list visited in 32-entry blocks from last to first.
//   reverse-mode through the step above with adjoints (Tb = dL/dT, Pb = dL/dPhi) carried per pixel;
//   the eleven per-Gaussian partials are reduced over the warp with the halving butterfly and
//   leave as one atomic instruction per warp and record.
__global__ void __launch_bounds__(CTA_THREADS)
composite_phase_bwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int2* __restrict__ ranges,
                           const float4* __restrict__ sorted_records, const uint32_t* __restrict__ sorted_gids,
                           const float* __restrict__ sorted_phases, float A, float3 bg,
                           const float* __restrict__ state_T, const int* __restrict__ state_n,
                           const float2* __restrict__ ckpt, const float* __restrict__ g_image,
                           const float* __restrict__ g_depth, const float* __restrict__ g_alpha,
                           float* __restrict__ grad2d, float* __restrict__ g_phases) {
    __shared__ StageBuf stage[2];
    __shared__ float phase_s[2][SUB];
    __shared__ uint32_t gid_s[2][SUB];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ int max_n_s;

    const int tile = blockIdx.x;
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * TILE + (threadIdx.x & (TILE - 1));
    const int py = ty * TILE + (threadIdx.x / TILE);
    const bool in_image = (px < width) && (py < height);
    const float fpx = (float)px, fpy = (float)py;
    const uint32_t pxy = (uint32_t)px | ((uint32_t)py << 16);
    const uint32_t pxy_guard = pxy | 0x80008000u, pxy_plus1 = pxy + 0x00010001u;
    const int lane = threadIdx.x & 31;
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
        const int gates = st >> STATE_GATE_SHIFT;
        gr = (gates & 1) ? g_image[ip] : 0.0f;
        gg = (gates & 2) ? g_image[ip + hw] : 0.0f;
        gb = (gates & 4) ? g_image[ip + 2 * hw] : 0.0f;
        gd = g_depth ? g_depth[pix] : 0.0f;
        ga = g_alpha ? g_alpha[pix] : 0.0f;
    }
    if (threadIdx.x == 0) {
        max_n_s = 0;
        for (int s = 0; s < 2; ++s) frb_mbar_init(&full_bar[s], 1);
        frb_mbar_fence_init();
    }
    __syncthreads();
    {
        int wmax = my_n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        if (lane == 0) atomicMax(&max_n_s, wmax);
    }
    __syncthreads();
    const int count = min(max_n_s, range.y - range.x);
    const int n_blocks = (count + SUB - 1) / SUB;
    if (n_blocks == 0) return;

    auto issue = [&](int visit) {
        int k = n_blocks - 1 - visit;
        int s = visit & 1;
        int cnt = min(SUB, count - k * SUB);
        frb_mbar_expect_tx(&full_bar[s], cnt * RECORD_BYTES);
        frb_tma_load_1d(stage[s].rec, sorted_records + 3 * (size_t)(range.x + k * SUB), cnt * RECORD_BYTES,
                        &full_bar[s]);
    };
    if (threadIdx.x == 0) {
        issue(0);
        if (n_blocks > 1) issue(1);
    }

    float Tb = gr * bg.x + gg * bg.y + gb * bg.z - ga;     // dL/dT_final
    float Pb = 0.0f;                                        // dL/dPhi_final
    const int gslot = warp_reduce_multi_index(lane);
    const int gofs = gslot + (gslot >= 7 ? 1 : 0);          // [du dv dA dB | dC do ddepth _ | dr dg db _]

    for (int visit = 0; visit < n_blocks; ++visit) {
        const int k = n_blocks - 1 - visit;
        const int s = visit & 1;
        const int cnt = min(SUB, count - k * SUB);
        if (threadIdx.x < cnt) {
            phase_s[s][threadIdx.x] = sorted_phases[range.x + k * SUB + threadIdx.x];
            gid_s[s][threadIdx.x] = sorted_gids[range.x + k * SUB + threadIdx.x];
        }
        frb_mbar_wait(&full_bar[s], (visit >> 1) & 1);
        __syncthreads();
        const float4* rec = stage[s].rec;
        const int local_n = my_n - k * SUB;                 // entries j < local_n were applied by this pixel

        // ---- recompute forward from the checkpoint, remembering Phi before every entry ----
        float phi_before[SUB], acc_before[SUB];
        float acc = 0.0f, Phi = 0.0f;
        uint32_t amask = 0;
        if (local_n > 0) {
            float2 c0 = ckpt[ckpt_slot(range.x, tile, k) * CTA_THREADS + threadIdx.x];
            acc = c0.x; Phi = c0.y;
        }
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
            phi_before[j] = Phi;
            acc_before[j] = acc;
            if (j < cnt && j < local_n) {
                float4 r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                if (rect_contains(pxy_guard, pxy_plus1, __float_as_uint(r1.w), __float_as_uint(r2.w))) {
                    float4 r0 = rec[3 * j + 0];
                    float dx = fpx - r0.x, dy = fpy - r0.y;
                    float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                    PhaseStep st;
                    phase_step(frb_ex2(power), r1.y, phase_s[s][j], A, acc, Phi, st);
                    acc = st.accn; Phi = st.Phin;
                    amask |= 1u << j;
                }
            }
        }

        // ---- walk back ----
#pragma unroll
        for (int j = SUB - 1; j >= 0; --j) {
            if (j >= cnt) continue;                                   // uniform
            const bool active = (amask >> j) & 1u;
            if (!__any_sync(0xffffffffu, active)) continue;
            float part[N_PHASE_GRADS];
#pragma unroll
            for (int q = 0; q < N_PHASE_GRADS; ++q) part[q] = 0.0f;
            if (active) {
                float4 r0 = rec[3 * j + 0], r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
                const float phi = phase_s[s][j];
                const float Phi0 = phi_before[j];
                float dx = fpx - r0.x, dy = fpy - r0.y;
                float power = dx * (r0.z * dx + r0.w * dy) + r1.x * (dy * dy);
                float g = frb_ex2(power);
                // replay the step from the remembered state before the entry
                PhaseStep st;
                phase_step(g, r1.y, phi, A, acc_before[j], Phi0, st);
                float sn = __sinf(st.d * TWO_PI_REF);
                const float a0 = st.a0, d0 = st.d0, m = st.m, a1 = st.a1, alpha = st.alpha, c = st.c;
                const float T0 = 1.0f - acc_before[j];
                const float acc_n = st.accn;
                const float iden = frb_rcp(st.den);
                const float pc = st.pc;
                // reverse mode
                float w = gr * r2.x + gg * r2.y + gb * r2.z + gd * r1.z;
                float pcb = Pb * (phi - Phi0);
                float phib = Pb * pc;
                float Pb0 = Pb * (1.0f - pc);
                float cb_ = pcb * iden + w;
                float Tb_tot = Tb + ((acc_n >= 1e-6f) ? pcb * c * iden * iden : 0.0f);
                cb_ -= Tb_tot;
                float Tb0 = Tb_tot + cb_ * alpha;
                float alphab = cb_ * T0;
                float a1b = (alpha == a1) ? alphab : 0.0f;            // clamp gate (inclusive)
                float a0b = a1b * m;
                float mb = a1b * a0;
                float db = -mb * A * TWO_PI_REF * sn;
                float d0b = (d0 < 1.0f - d0) ? db : ((d0 > 1.0f - d0) ? -db : 0.0f);
                float sg = (phi > Phi0) ? 1.0f : ((phi < Phi0) ? -1.0f : 0.0f);
                phib += d0b * sg;
                Pb0 -= d0b * sg;
                float gda = g * a0b;                                   // dL/dopacity
                float dpow = gda * r1.y * FRB_LN2;
                part[0] = -(2.0f * r0.z * dx + r0.w * dy) * dpow;
                part[1] = -(r0.w * dx + 2.0f * r1.x * dy) * dpow;
                part[2] = dx * dx * dpow;
                part[3] = dx * dy * dpow;
                part[4] = dy * dy * dpow;
                part[5] = gda;
                part[6] = c * gd;
                part[7] = c * gr;
                part[8] = c * gg;
                part[9] = c * gb;
                part[10] = phib;
                Tb = Tb0; Pb = Pb0;
            }
            float tot = warp_reduce_multi<N_PHASE_GRADS>(part, lane);
            if ((lane & 1) == 0 && gslot < N_PHASE_GRADS && tot != 0.0f) {
                const uint32_t gid = gid_s[s][j];
                if (gslot == 10) atomicAdd(g_phases + gid, tot);
                else atomicAdd(grad2d + (size_t)gid * FRB_GRAD_FLOATS + gofs, tot);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && visit + 2 < n_blocks) issue(visit + 2);
    }
}

}  // namespace

extern "C" size_t frb_phase_ckpt_floats(int m, int n_tiles) {
    if (m < 0) m = 0;
    if (n_tiles < 0) n_tiles = 0;
    return ((size_t)(m >> 5) + (size_t)n_tiles + 2) * CTA_THREADS * 2;
}

int frb_composite_phase_fwd_launch(int n_views, int width, int height, const int32_t* ranges,
                                   const float* sorted_records, const float* sorted_phases, float phase_amplitude,
                                   const float* background_host, float t_eps, float* image, float* depth,
                                   float* alpha, float* state_T, int32_t* state_n, float* ckpt, cudaStream_t st) {
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    composite_phase_fwd_kernel<<<n_views * tpv, CTA_THREADS, 0, st>>>(
        width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records, sorted_phases,
        phase_amplitude, bg, t_eps, image, depth, alpha, state_T, state_n, (float2*)ckpt);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

int frb_composite_phase_bwd_launch(int n_views, int width, int height, const int32_t* ranges,
                                   const float* sorted_records, const uint32_t* sorted_gids,
                                   const float* sorted_phases, float phase_amplitude, const float* background_host,
                                   const float* state_T, const int32_t* state_n, const float* ckpt,
                                   const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                                   float* g_phases, cudaStream_t st) {
    if (!ckpt || !g_phases) return FRB_E_INVALID;
    int tiles_x = frb_div_up(width, TILE), tiles_y = frb_div_up(height, TILE);
    int tpv = tiles_x * tiles_y;
    float3 bg = make_float3(background_host[0], background_host[1], background_host[2]);
    composite_phase_bwd_kernel<<<n_views * tpv, CTA_THREADS, 0, st>>>(
        width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records, sorted_gids, sorted_phases,
        phase_amplitude, bg, state_T, state_n, (const float2*)ckpt, g_image, g_depth, g_alpha, grad2d, g_phases);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
