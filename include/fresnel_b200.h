/*
 * fresnel_b200 C-ABI: the drop-in boundary of the B200-native Gaussian-splatting
 * renderer.  Plain pointers and sizes only; every pointer is a DEVICE pointer
 * unless the name ends in _host.  All functions are asynchronous on `stream`
 * (a cudaStream_t passed as void*), return 0 on success and a positive
 * cudaError_t / negative FRB_E_* code otherwise, and never fall back to the CPU.
 *
 * Reference interfaces replaced (paths relative to the reference repository,
 * DR = scripts/models/differentiable_renderer.py):
 *   frb_project_fwd / frb_project_bwd     compute_2d_covariance DR:123-195,
 *                                         quaternion_to_rotation_matrix DR:98-120,
 *                                         TileBasedRenderer._compute_radius DR:452-487,
 *                                         visibility + rectangles DR:541-543, 594-600,
 *                                         pinv(cov + 1e-4 I) DR:578-579, and their autograd
 *   frb_depth_order / frb_tile_offsets /  torch.argsort(depths) DR:527-538 and the mask
 *   frb_bin_emit / frb_radix_sort_pairs / compaction DR:554-562, restated as a stable
 *   frb_tile_ranges / frb_gather_records  (tile|depth) 64-bit key sort (no reference
 *                                         counterpart for tiles; nearest native code is
 *                                         src/core/compute/radix_sort.cpp)
 *   frb_composite_fwd / frb_composite_bwd the per-Gaussian loop DR:582-667, the epilogue
 *                                         DR:669-686 and the autograd tape through them
 *                                         (phases != NULL: the Fresnel phase blending
 *                                         DR:571-575, 629-645, 660-667)
 *   frb_wave_*                            WaveFieldRenderer.forward DR:747-926
 *   frb_asm_*                             AngularSpectrumPropagator DR:929-1065 and
 *                                         ASMWaveFieldRenderer.forward DR:1150-1344
 *
 * Views.  Every entry point renders `n_views` independent views in one call (the per-view
 * loop of scripts/training/train_gaussian_decoder.py:1209-1223).  The n Gaussians are
 * n_views consecutive groups of n / n_views, group k seen by camera k; tile ids are global:
 * view * tiles_per_view + tile_y * tiles_x + tile_x.  n_views <= FRB_MAX_VIEWS.
 *
 * Layouts
 *   camera_host : n_views x 20 floats = view-matrix rows 0..2 (12), fx, fy, cx, cy, width,
 *                 height, near, far  (Camera DR:24-52); width/height equal for all views
 *   records     : FRB_RECORD_FLOATS floats per Gaussian
 *                 [u, v, A', B' | C', opacity, rect_lo, rect_hi | r, g, b, depth]
 *                 (the two rectangle words are one 8-byte load; (r, g) and (b, depth) feed packed fp32 fmas)
 *                 A',B',C' = conic * (-0.5*log2 e), B = inv01 + inv10 (DR:618);
 *                 rect_lo = x0 | y0 << 16, rect_hi = x1 | y1 << 16 | 0x80008000 (bit-cast);
 *                 the rectangle is [x0,x1) x [y0,y1) of DR:594-597, all zero when culled
 *   rects       : int32 [x0, x1, y0, y1] per Gaussian (same numbers, unpacked)
 *   grad2d      : FRB_GRAD_FLOATS floats per Gaussian, accumulated by composite_bwd
 *                 [du, dv, dA', dB' | dC', dopacity, ddepth, 0 | dr, dg, db, 0]
 *   keys        : (global_tile_id << 32) | IEEE-754 bits of the fp32 depth
 *   ranges      : [start, end) into the sorted instance list, per 16x16 tile
 *   image       : [n_views][3][H][W]; depth, alpha, state_T, state_n : [n_views][H][W]
 */
#ifndef FRESNEL_B200_H
#define FRESNEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRB_TILE 16
#define FRB_RECORD_FLOATS 12
#define FRB_GRAD_FLOATS 12
#define FRB_CAMERA_FLOATS 20
#define FRB_DEBUG_FLOATS 8
#define FRB_MAX_VIEWS 32
#define FRB_MAX_IMAGE_SIDE 32767

#define FRB_E_INVALID (-1)   /* bad argument (size, alignment, null pointer) */
#define FRB_E_TOO_LARGE (-2) /* image side > FRB_MAX_IMAGE_SIDE, too many views or instances */

int frb_version(void);
const char* frb_error_string(int code);
/* Number of CUDA kernels launched by this library since it was loaded. */
unsigned long long frb_launch_count(void);

/* ---- projection ------------------------------------------------------- */
/* rects, debug nullable.  debug: [cov a, b, c, d, radius, visible, 0, 0] per Gaussian. */
int frb_project_fwd(int n, int n_views, const float* positions, const float* scales,
                    const float* rotations, const float* colors, const float* opacities,
                    const float* camera_host, float max_radius, float* records, int32_t* rects,
                    uint32_t* depth_bits, uint32_t* tiles_touched, float* debug, void* stream);

/* g_colors / g_opacities nullable (they are plain copies of grad2d columns). */
int frb_project_bwd(int n, int n_views, const float* positions, const float* scales,
                    const float* rotations, const float* camera_host, const float* grad2d,
                    float* g_positions, float* g_scales, float* g_rotations, float* g_colors,
                    float* g_opacities, void* stream);

/* Projection for the other renderers of the same reference file.  mode:
 *   0 = FRB_MODE_TILE     TileBasedRenderer / WaveFieldRenderer / ASMWaveFieldRenderer (== the two functions above)
 *   1 = FRB_MODE_DENSE    DifferentiableGaussianRenderer  differentiable_renderer.py:245-409
 *   2 = FRB_MODE_FOURIER  FourierGaussianRenderer         differentiable_renderer.py:1500-1774
 * Same record layout; only visibility, support rectangle and the conic differ (csrc/frb_math.h). */
int frb_project_fwd_mode(int n, int n_views, const float* positions, const float* scales,
                         const float* rotations, const float* colors, const float* opacities,
                         const float* camera_host, float max_radius, int mode, float* records, int32_t* rects,
                         uint32_t* depth_bits, uint32_t* tiles_touched, float* debug, void* stream);
int frb_project_bwd_mode(int n, int n_views, const float* positions, const float* scales,
                         const float* rotations, const float* camera_host, const float* grad2d, int mode,
                         float* g_positions, float* g_scales, float* g_rotations, float* g_colors,
                         float* g_opacities, void* stream);

/* ---- binning: depth order, offsets, keys, stable radix sort, ranges, gather -------- */
/* Stable LSD radix sort of (key, value) pairs on key bits [begin_bit, end_bit).
 * The sorted result is always left in keys / vals; *_tmp are scratch of the same size. */
size_t frb_sort_workspace_bytes(int m);
int frb_radix_sort_pairs(int m, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                         uint32_t* vals_tmp, int begin_bit, int end_bit, void* workspace,
                         void* stream);
/* order[k] = index of the k-th Gaussian in stable ascending depth_bits order (all n of them). */
size_t frb_depth_order_workspace_bytes(int n);
int frb_depth_order(int n, const uint32_t* depth_bits, uint32_t* order, void* workspace,
                    void* stream);
/* The same, and rank[g] = position of Gaussian g in that order (the inverse permutation; nullable). */
int frb_depth_order_rank(int n, const uint32_t* depth_bits, uint32_t* order, uint32_t* rank, void* workspace,
                         void* stream);
/* The same for what a camera can see: keys are the depth bits clamped to the bit patterns of [near_depth, far_depth]
 * minus the lower one, sorted on the bits that range needs (fewer passes for narrow slabs; 2^27 values and more fall
 * back to frb_depth_order_rank).  The order of the visible Gaussians
 * (near < depth < far, DR:541) is exactly frb_depth_order's; culled ones collapse to the two ends in index order. */
int frb_depth_order_range(int n, const uint32_t* depth_bits, float near_depth, float far_depth, uint32_t* order,
                          uint32_t* rank, void* workspace, void* stream);
/* For n <= 131,072 the three functions above have a second implementation: ONE kernel on one 16-CTA thread-block
 * cluster that keeps all (key, index) pairs in distributed shared memory through all passes (no workspace traffic, no
 * look-back, bit-identical result).  It occupies 16 SMs for longer than the multi-kernel chain occupies all of them:
 * slower for one frame at a time, faster for the whole device when other streams have work for the remaining SMs.
 * mode 1 = use it, 0 = do not, -1 = environment variable FRB_CLUSTER_SORT (unset: 0; the initial state).  Returns the
 * previous mode, FRB_E_INVALID for another value.  Process-wide; it decides which kernels a call enqueues (a captured
 * CUDA graph keeps what it was captured with). */
int frb_depth_sort_in_cluster(int mode);
/* offsets[k] = sum_{j<k} tiles_touched[order[j]] for k = 0..n (order NULL = identity);
 * offsets[n] = number of tile instances M. */
size_t frb_scan_workspace_bytes(int n);
int frb_tile_offsets(int n, const uint32_t* tiles_touched, const uint32_t* order, uint32_t* offsets,
                     void* workspace, void* stream);
/* Slot k emits the keys of Gaussian order[k] at offsets[k] (tiles in row-major order). */
int frb_bin_emit(int n, int n_views, int width, int height, const float* records,
                 const uint32_t* depth_bits, const uint32_t* order, const uint32_t* offsets,
                 uint64_t* keys, uint32_t* gids, void* stream);
int frb_tile_ranges(int m, const uint64_t* keys, int n_tiles, int32_t* ranges, void* stream);
/* phases / sorted_phases nullable (both or neither). */
int frb_gather_records(int m, const uint32_t* gids, const float* records, float* sorted_records,
                       const float* phases, float* sorted_phases, void* stream);

/* frb_tile_ranges + frb_gather_records in one launch (what the renderer calls). */
int frb_ranges_and_gather(int m, const uint64_t* keys, const uint32_t* gids, int n_tiles,
                          int32_t* ranges, const float* records, float* sorted_records,
                          const float* phases, float* sorted_phases, void* stream);

/* Capacity mode (no host synchronisation): buffers hold m_capacity instances, the true count is read
 * on the device from *m_dev (= offsets[n] of frb_tile_offsets); entries past it are never touched. */
int frb_radix_sort_pairs_dev(int m_capacity, const uint32_t* m_dev, uint64_t* keys, uint32_t* vals,
                             uint64_t* keys_tmp, uint32_t* vals_tmp, int begin_bit, int end_bit,
                             void* workspace, void* stream);
int frb_ranges_and_gather_dev(int m_capacity, const uint32_t* m_dev, const uint64_t* keys,
                              const uint32_t* gids, int n_tiles, int32_t* ranges, const float* records,
                              float* sorted_records, const float* phases, float* sorted_phases,
                              void* stream);

/* ---- compositing ------------------------------------------------------ */
/* t_eps: a pixel stops once its transmittance has fallen below max(t_eps, 1e-20).
 * sorted_phases == NULL selects plain alpha compositing; otherwise Fresnel phase blending
 * with phase_amplitude.  state_T / state_n are per-pixel state for the backward pass
 * (final transmittance; number of list entries consumed | image-clamp gates << 28).  ckpt (nullable unless phases are
 * given and a backward pass will follow) receives per-batch (alpha, phase) checkpoints. */
size_t frb_phase_ckpt_floats(int m, int n_tiles);
int frb_composite_fwd(int n_views, int width, int height, const int32_t* ranges,
                      const float* sorted_records, const float* sorted_phases,
                      float phase_amplitude, const float* background_host, float t_eps,
                      float* image, float* depth, float* alpha, float* state_T, int32_t* state_n,
                      float* ckpt, void* stream);

/* Launch order: tile_order[k] = k-th tile to start, longest lists first (nullable = row-major order).
 * The *_sched variants of the compositor take it as an extra argument. */
int frb_tile_schedule(int n_tiles, const int32_t* ranges, int32_t* tile_order, void* stream);
int frb_composite_fwd_sched(int n_views, int width, int height, const int32_t* tile_order,
                            const int32_t* ranges, const float* sorted_records, const float* sorted_phases,
                            float phase_amplitude, const float* background_host, float t_eps, float* image,
                            float* depth, float* alpha, float* state_T, int32_t* state_n, float* ckpt,
                            void* stream);
int frb_composite_bwd_sched(int n_views, int width, int height, const int32_t* tile_order,
                            const int32_t* ranges, const float* sorted_records, const uint32_t* sorted_gids,
                            const float* sorted_phases, float phase_amplitude, const float* background_host,
                            const float* state_T, const int32_t* state_n, const float* ckpt,
                            const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                            float* g_phases, void* stream);

/* g_depth, g_alpha nullable (treated as zero).  grad2d (and g_phases) must be zeroed by the
 * caller; contributions are accumulated with atomics. */
int frb_composite_bwd(int n_views, int width, int height, const int32_t* ranges,
                      const float* sorted_records, const uint32_t* sorted_gids,
                      const float* sorted_phases, float phase_amplitude,
                      const float* background_host, const float* state_T,
                      const int32_t* state_n, const float* ckpt, const float* g_image,
                      const float* g_depth, const float* g_alpha, float* grad2d, float* g_phases,
                      void* stream);

/* ---- tile lists by counting + bitmap ranking (csrc/tile_lists.cu) ----------------------------------------
 * The same sorted per-tile lists as the key sort above (ascending tile, depth bits, Gaussian index), built without
 * sorting M keys: count the instances per tile, scan, emit each Gaussian's DEPTH RANK into its tiles' spans
 * (arbitrary order inside a span), then one CTA per tile orders its span with a shared-memory bitmap over the ranks
 * and gathers the records.  Needs n <= frb_tile_lists_max_gaussians() (the bitmap must fit one SM's shared memory).
 * and the image must have at most frb_tile_lists_max_tiles() tiles over all views (one counter per tile in shared
 * memory).  workspace: frb_tile_lists_workspace_bytes(n, n_tiles), written by frb_tile_count and frb_tile_scan,
 * consumed by frb_tile_emit; order / rank: frb_depth_order_rank's outputs; m_out[0] = total number of instances
 * (may exceed m_capacity: then the lists are truncated to the capacity and the caller must retry with more room);
 * inst_rank:
 * m_capacity words; sorted_keys (nullable): the 64-bit (tile | depth bits) keys, for checks. */
int frb_tile_lists_max_gaussians(void);
int frb_tile_lists_max_tiles(void);
size_t frb_tile_lists_workspace_bytes(int n, int n_tiles);
int frb_tile_count(int n, int n_views, int width, int height, const float* records, void* workspace,
                   void* stream);
int frb_tile_scan(int n, int n_tiles, int m_capacity, int32_t* ranges, int32_t* tile_order, uint32_t* m_out,
                  const uint32_t* upstream_error, void* workspace, void* stream);
/* m_out: TWO words - [0] the instance count, [1] a status: bit 0 = the lists were truncated to m_capacity, bit 1 =
 * *upstream_error was set (nullable; frb_depth_order_error_word: a radix pass of the depth sort gave up on a
 * look-back, its order is not trustworthy).  Callers that read the count read the status with it. */
const uint32_t* frb_depth_order_error_word(int n, const void* workspace);
/* frb_tile_count + frb_tile_scan in ONE kernel (what the renderers call): the chunk CTAs obtain "instances of tile t
 * in earlier chunks" by decoupled look-back over each other's histogram rows, the CTA that started last runs the
 * tile scan on the column totals it ends up with.  Same outputs. */
int frb_tile_count_scan(int n, int n_views, int width, int height, const float* records, int m_capacity,
                        int32_t* ranges, int32_t* tile_order, uint32_t* m_out, const uint32_t* upstream_error,
                        void* workspace, void* stream);
int frb_tile_emit(int n, int n_views, int width, int height, const float* records, const uint32_t* rank,
                  int m_capacity, void* workspace, uint32_t* inst_rank, void* stream);
int frb_tile_rank_gather(int n, int n_tiles, const int32_t* tile_order, const int32_t* ranges,
                         const uint32_t* inst_rank, const uint32_t* order, const float* records,
                         const uint32_t* depth_bits, const float* phases, uint32_t* sorted_gids,
                         float* sorted_records, float* sorted_phases, uint64_t* sorted_keys, void* stream);

/* frb_composite_fwd_cap reading the tile lists as Gaussian ids: records is the UNSORTED array written by
 * frb_project_fwd (n_records rows of 12 floats), sorted_gids the per-tile lists; the records are fetched by TMA
 * tile::gather4 (no sorted copy of the records is needed).  No phase blending. */
int frb_composite_fwd_gather(int n_views, int width, int height, const int32_t* tile_order, const int32_t* ranges,
                             const float* records, int n_records, const uint32_t* sorted_gids,
                             const float* background_host, float t_eps, float alpha_max, float* image, float* depth,
                             float* alpha, float* state_T, int32_t* state_n, void* stream);

int frb_composite_bwd_gather(int n_views, int width, int height, const int32_t* tile_order, const int32_t* ranges,
                             const float* records, int n_records, const uint32_t* sorted_gids,
                             const float* background_host, float alpha_max, const float* state_T,
                             const int32_t* state_n, const float* g_image, const float* g_depth, const float* g_alpha,
                             float* grad2d, void* stream);

/* ---- whole-pass entry points (capacity mode, no host synchronisation) ------------------ */
/* One call enqueues projection, binning and compositing of TileBasedRenderer.forward (DR:489-686)
 * for n_views views; buffers are carved from two arenas laid out by frb_tile_layout: `persist`
 * (kept for the backward pass) and `scratch` (free after the call's work has run).  m_capacity bounds
 * the number of tile instances (n * max tiles a rectangle can cover is always enough). */
/* frb_tile_offsets + frb_bin_emit + frb_radix_sort_pairs_dev(begin 32, end 32 + tile_bits) with the scan, the key
 * emission and the sort's digit histograms fused into one kernel (tile_bits <= 16).  m_out: device word receiving the
 * instance count M; scan_ws: frb_scan_workspace_bytes(n); sort_ws: frb_sort_workspace_bytes(m_capacity).  The sorted
 * pairs end in keys / gids. */
int frb_bin_sort_dev(int n, int n_views, int width, int height, const float* records, const uint32_t* depth_bits,
                     const uint32_t* tiles_touched, const uint32_t* order, int m_capacity, uint32_t* m_out,
                     uint64_t* keys, uint32_t* gids, uint64_t* keys_tmp, uint32_t* vals_tmp, int tile_bits,
                     void* scan_ws, void* sort_ws, void* stream);

typedef struct FrbTileLayout {
    size_t ranges, tile_order, state_T, state_n, sorted_gids, sorted_records, records, persist_bytes;
    size_t depth_bits, touched, order, offsets, depth_ws, scan_ws, keys, keys_tmp, vals_tmp,
        sort_ws, tile_ws, inst_rank, rank, scratch_bytes;
} FrbTileLayout;
int frb_tile_layout(int n, int n_views, int width, int height, int m_capacity, FrbTileLayout* layout);
int frb_tile_render_fwd(int n, int n_views, const float* positions, const float* scales,
                        const float* rotations, const float* colors, const float* opacities,
                        const float* camera_host, float max_radius, int width, int height,
                        const float* background_host, float t_eps, int m_capacity, void* persist,
                        void* scratch, float* image, float* depth, float* alpha, void* stream);
int frb_tile_render_bwd(int n, int n_views, const float* positions, const float* scales,
                        const float* rotations, const float* camera_host, int width, int height,
                        const float* background_host, int m_capacity, const void* persist,
                        const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                        float* g_positions, float* g_scales, float* g_rotations, float* g_colors,
                        float* g_opacities, void* stream);

/* Per-stage timing of the whole-pass entry points (measurement aid, bench.py's roofline line): with timing
 * enabled frb_tile_render_fwd / _bwd bracket every stage they enqueue with CUDA events on the caller's stream.
 * frb_stage_timing_enable(on) clears what was recorded; _get(i) waits for stage i and returns its name (static
 * string) and duration.  Do not enable inside a stream capture. */
int frb_stage_timing_enable(int on);
int frb_stage_timing_count(void);
int frb_stage_timing_get(int i, const char** name, float* ms);

/* ---- complex wave field: WaveFieldRenderer DR:747-926 ------------------------------------ */
/* wc: 8 floats per Gaussian [colour_c cos(phi_c) x3, colour_c sin(phi_c) x3, 0, 0];
 * phases: n x phase_stride floats, phase_stride = 1 (scalar phase) or 3 (per channel), radians.
 * accum / gpix: [n_views][8][H][W] planar (Re rgb, Im rgb, sum amp*depth, sum amp) and its gradient.
 * rmax_bits: per view, bits of max over pixels and channels of sqrt(I + 1e-8) (DR:902).
 * red: 2 floats per view of scratch. */
int frb_wave_prepare(int n, const float* colors, const float* phases, int phase_stride, float* wc,
                     void* stream);
int frb_wave_gather(int m, const uint32_t* gids, const float* wc, float* sorted_wc, void* stream);
int frb_wave_splat_fwd(int n_views, int width, int height, const int32_t* ranges,
                       const float* sorted_records, const float* sorted_wc, float* accum,
                       uint32_t* rmax_bits, void* stream);
int frb_wave_finish_fwd(int n_views, int width, int height, const float* accum,
                        const uint32_t* rmax_bits, const float* background_host, float* image,
                        float* depth, void* stream);
int frb_wave_finish_bwd(int n_views, int width, int height, const float* accum,
                        const uint32_t* rmax_bits, const float* background_host, const float* g_image,
                        const float* g_depth, float* red, float* gpix, void* stream);
/* n_planes == 0: wave (gpix as above); n_planes > 0: ASM (gpix = d_fields, see below, times scale).
 * grad2d (slots 0..6) and gwc (8 floats per Gaussian) must be zeroed by the caller. */
int frb_wave_splat_bwd(int n_views, int width, int height, int n_planes, const int32_t* ranges,
                       const float* sorted_records, const float* sorted_wc, const uint32_t* sorted_gids,
                       const uint64_t* keys, const float* gpix, float scale, float* grad2d, float* gwc,
                       void* stream);
/* colour / phase gradients from gwc: writes grad2d slots 8..10 and g_phases (n x phase_stride). */
int frb_wave_chain_bwd(int n, const float* colors, const float* phases, int phase_stride,
                       const float* gwc, float* grad2d, float* g_phases, void* stream);

/* ---- angular spectrum: AngularSpectrumPropagator DR:929-1065, ASMWaveFieldRenderer DR:1150-1344 */
/* plane_idx[i] = argmin_p |depth_i - depth_planes[p]| (DR:1136-1148); it replaces the depth bits as
 * the low word of the sort key, so every tile list is grouped by plane.
 * fields / d_fields: [n_views][n_planes][3][H][W] complex64; total / g_total: [n_views][3][H][W]. */
int frb_asm_assign_planes(int n, const float* records, int n_planes, const float* depth_planes_host,
                          uint32_t* plane_idx, void* stream);
int frb_asm_splat_fwd(int n_views, int width, int height, int n_planes, const int32_t* ranges,
                      const float* sorted_records, const float* sorted_wc, const uint64_t* keys,
                      float* fields, void* stream);
int frb_asm_propagate_fwd(int n_views, int width, int height, int n_planes,
                          const float* depth_planes_host, float focal_depth, float pixel_pitch,
                          const float* wavelengths_host, const float* background_host, float* fields,
                          float* total, uint32_t* rmax_bits, float* image, void* stream);
int frb_asm_propagate_bwd(int n_views, int width, int height, int n_planes,
                          const float* depth_planes_host, float focal_depth, float pixel_pitch,
                          const float* wavelengths_host, const float* background_host,
                          const float* total, const uint32_t* rmax_bits, const float* g_image, float* red,
                          float* g_total, float* d_fields, void* stream);

/* The plain compositor with the alpha clamp as an argument (0 < alpha_max < 1; the tile renderer's is 0.99,
 * differentiable_renderer.py:647).  Used by SimplifiedRenderer, whose clamp(alpha, 0, 1) (:1430) becomes
 * alpha_max = 1 - 2^-24.  sorted_phases must be NULL unless alpha_max is the default. */
int frb_composite_fwd_cap(int n_views, int width, int height, const int32_t* tile_order, const int32_t* ranges,
                          const float* sorted_records, const float* sorted_phases, float phase_amplitude,
                          const float* background_host, float t_eps, float alpha_max, float* image, float* depth,
                          float* alpha, float* state_T, int32_t* state_n, float* ckpt, void* stream);
int frb_composite_bwd_cap(int n_views, int width, int height, const int32_t* tile_order, const int32_t* ranges,
                          const float* sorted_records, const uint32_t* sorted_gids, const float* sorted_phases,
                          float phase_amplitude, const float* background_host, float alpha_max, const float* state_T,
                          const int32_t* state_n, const float* ckpt, const float* g_image, const float* g_depth,
                          const float* g_alpha, float* grad2d, float* g_phases, void* stream);

/* ---- SimplifiedRenderer (differentiable_renderer.py:1347-1458): point splats with an integer radius -------
 * frb_simple_project_fwd: Camera.project (:54-85), radius = min(int(max(mean(scale) fx / depth, 1)), 20),
 * rectangle [int(u) - r, int(u) + r + 1) x [int(v) - r, int(v) + r + 1) clipped to the image, isotropic conic
 * exp(-d^2 / (2 max(r/2, 1)^2)); same record layout as frb_project_fwd.
 * frb_simple_depth_fwd / bwd: depth map = depth of the front-most list entry with alpha > 0.1 (:1438-1442);
 * hit: [views, H, W] Gaussian ids (-1 = none).  frb_simple_project_bwd: chain (u, v, depth) -> positions;
 * scales and rotations receive no gradient in the reference (radius goes through .item()). */
int frb_simple_project_fwd(int n, int n_views, const float* positions, const float* scales, const float* colors,
                           const float* opacities, const float* camera_host, float* records, uint32_t* depth_bits,
                           uint32_t* tiles_touched, void* stream);
int frb_simple_project_bwd(int n, int n_views, const float* positions, const float* camera_host,
                           const float* grad2d, float* g_positions, float* g_colors, float* g_opacities,
                           void* stream);
int frb_simple_depth_fwd(int n_views, int width, int height, const int32_t* ranges, const float* sorted_records,
                         const uint32_t* sorted_gids, float* depth, int32_t* hit, void* stream);
int frb_simple_depth_bwd(int n_views, int width, int height, const int32_t* hit, const float* g_depth,
                         float* grad2d, void* stream);

/* ---- FourierGaussianRenderer epilogue (differentiable_renderer.py:1740-1753) and its backward ----
 * accum: [view][8][H][W] as written by frb_wave_splat_fwd on FRB_MODE_FOURIER records with zero phases
 * (planes 0..2 = channel sums); mx_key: n_views words (global maximum, order-preserving key);
 * red: 2 * n_views floats scratch; gpix: [view][8][H][W] for frb_wave_splat_bwd. */
int frb_fourier_finish_fwd(int n_views, int width, int height, const float* accum, uint32_t* mx_key,
                           const float* background_host, float* image, void* stream);
int frb_fourier_finish_bwd(int n_views, int width, int height, const float* accum, const uint32_t* mx_key,
                           const float* background_host, const float* g_image, float* red, float* gpix,
                           void* stream);

/* ---- on-disk formats --------------------------------------------------------
 * rows: n x 14 floats as stored in the file body [position 3 | scale 3 | rotation wxyz 4 | colour 3 | opacity 1].
 * ply = 0: the reference's .bin (differentiable_renderer.py:1461-1497, renderer.cpp:557-647), a plain split;
 * ply = 1: 3DGS .ply rows (renderer.cpp:649-793): exp / log scale, SH-DC colour, sigmoid / logit opacity. */
int frb_unpack_gaussians(int n, int ply, const float* rows, float* positions, float* scales, float* rotations,
                         float* colors, float* opacities, void* stream);
int frb_pack_gaussians(int n, int ply, const float* positions, const float* scales, const float* rotations,
                       const float* colors, const float* opacities, float* rows, void* stream);

/* ---- decoder output head -----------------------------------------------------
 * DirectPatchDecoder.forward tail, scripts/models/gaussian_decoder_models.py:807-948 (default flags + the
 * edge-aware modulation :881-895), fused with the stochastic subsampling gather of
 * scripts/training/train_gaussian_decoder.py:1154-1187.
 * raw: [B, H*W*K, 16] MLP outputs; depth_grid: [B, H, W] (nullable: z = depth_offset); depth_offset: device
 * scalar; edge: [B, H, W] edge strength (nullable = edge-aware off); idx: n_sel int64 Gaussian indices kept for
 * every view (nullable = all H*W*K).  Outputs are [B, n_out, .] with n_out = n_sel or H*W*K.
 * Backward: g_raw [B, H*W*K, 16] is fully written (zero rows for unselected Gaussians); g_depth_offset[1].
 * Any of the five incoming gradients may be NULL (= zero). */
int frb_decode_head_fwd(int B, int H, int W, int K, const float* raw, const float* depth_grid,
                        const float* depth_offset, const float* edge, float edge_scale_factor,
                        float edge_opacity_boost, const long long* idx, int n_sel, float* positions, float* scales,
                        float* rotations, float* colors, float* opacities, void* stream);
int frb_decode_head_bwd(int B, int H, int W, int K, const float* raw, const float* edge, float edge_scale_factor,
                        float edge_opacity_boost, const long long* idx, int n_sel, const float* g_positions,
                        const float* g_scales, const float* g_rotations, const float* g_colors,
                        const float* g_opacities, float* g_raw, float* g_depth_offset, void* stream);

/* The full tail of DirectPatchDecoder.forward: the above plus the Fresnel depth-zone snap of the depth grid
 * (gaussian_decoder_models.py:833-838 with FresnelZones.get_zone_centers_for_depth, utils/fresnel_zones.py:96-139),
 * the pose rotation of the positions (rotate_positions_for_pose :51-104, applied at :860) and the gradient of the
 * edge strength (the edge detector is a trained module).  ex may be NULL (= all extras off).
 * g_edge: [B, H, W], nullable, zeroed and accumulated here. */
typedef struct FrbHeadExtras {
    const float* edge;                  /* device [B, H, W] edge strength, NULL = edge-aware placement off */
    float edge_scale_factor, edge_opacity_boost;
    int num_zones;                      /* 0 = no zone snap */
    const float* zone_boundaries_host;  /* num_zones + 1 floats (FresnelZones.zone_boundaries), host */
    const float* zone_centers_host;     /* num_zones floats (FresnelZones.zone_centers), host */
    const float* pose_trig;             /* device [B, 4]: cos az, sin az, cos el, sin el; NULL = no rotation */
} FrbHeadExtras;
int frb_decode_head_fwd_ex(int B, int H, int W, int K, const float* raw, const float* depth_grid,
                           const float* depth_offset, const FrbHeadExtras* ex, const long long* idx, int n_sel,
                           float* positions, float* scales, float* rotations, float* colors, float* opacities,
                           void* stream);
int frb_decode_head_bwd_ex(int B, int H, int W, int K, const float* raw, const FrbHeadExtras* ex,
                           const long long* idx, int n_sel, const float* g_positions, const float* g_scales,
                           const float* g_rotations, const float* g_colors, const float* g_opacities, float* g_raw,
                           float* g_depth_offset, float* g_edge, void* stream);

/* ---- reconstruction loss front-end -------------------------------------------
 * compute_losses, scripts/training/train_gaussian_decoder.py:838-930 (L1 RGB + normalised-depth L1; the SSIM /
 * LPIPS terms need packages that are absent and are dropped by the reference too).
 * rendered, target: n_rgb floats; rendered_depth, target_depth: n_pix floats (both NULL = no depth term);
 * stats: frb_recon_loss_workspace_bytes() of device scratch shared by forward and backward; loss, g_loss:
 * device scalars. */
size_t frb_recon_loss_workspace_bytes(void);
int frb_recon_loss_fwd(long long n_rgb, long long n_pix, const float* rendered, const float* target,
                       const float* rendered_depth, const float* target_depth, float rgb_weight,
                       float depth_weight, void* stats, float* loss, void* stream);
int frb_recon_loss_bwd(long long n_rgb, long long n_pix, const float* rendered, const float* target,
                       const float* rendered_depth, const float* target_depth, float rgb_weight,
                       float depth_weight, const void* stats, const float* g_loss, float* g_rendered,
                       float* g_rendered_depth, void* stream);

/* The same with the Fresnel boundary-emphasis term of train_gaussian_decoder.py:941-953:
 *   + boundary_weight * mean_{view,pixel} [ mean_c |rendered - target| * mask(target_depth) ],
 * mask = FresnelZones.compute_boundary_mask (scripts/utils/fresnel_zones.py:141-180): soft
 * sigmoid(10 / thr * (thr - min_k |d - boundary_k|)) or hard 1[min_k |d - boundary_k| < thr].
 * rendered / target: [views, 3, hw]; target_depth: [views, hw] (needed for the mask even without rendered_depth);
 * boundaries_host: the n_boundaries = num_zones + 1 zone boundaries (torch.linspace values, at most 65);
 * boundary_weight = 0 disables the term. */
int frb_recon_loss_fwd_ex(long long n_rgb, long long n_pix, long long hw, const float* rendered, const float* target,
                          const float* rendered_depth, const float* target_depth, float rgb_weight,
                          float depth_weight, float boundary_weight, int n_boundaries, const float* boundaries_host,
                          float boundary_threshold, int soft_boundaries, void* stats, float* loss, void* stream);
int frb_recon_loss_bwd_ex(long long n_rgb, long long n_pix, long long hw, const float* rendered, const float* target,
                          const float* rendered_depth, const float* target_depth, float rgb_weight,
                          float depth_weight, float boundary_weight, int n_boundaries, const float* boundaries_host,
                          float boundary_threshold, int soft_boundaries, const void* stats, const float* g_loss,
                          float* g_rendered, float* g_rendered_depth, void* stream);

/* ---- gradient exchange fused with the optimiser step over NVLink peer memory ---
 * Multi-view optimisation of one replicated cloud (BASELINE configs[4]): replaces the pair
 * dist.all_reduce(grad) + torch.optim.Adam.step() that follows loss.backward() in the reference's loop
 * (scripts/training/train_gaussian_decoder.py:1261-1266) by ONE kernel per rank: entry barrier, reduce-scatter
 * by peer loads (fixed rank order), Adam on the owned shard (torch.optim.Adam semantics: no weight decay, no
 * amsgrad; moments sharded), all-gather by peer stores, exit barrier.
 * grad_ptrs / param_ptrs / signal_ptrs: DEVICE arrays of `world` 64-bit addresses, entry k = rank k's buffer
 * as mapped into this process (peer-mapped symmetric memory; world = 1: plain local pointers).  Gradient and
 * parameter buffers hold n_floats fp32 and are 16-byte aligned; a signal pad is 2 * world uint32, zero before
 * the first step.  exp_avg / exp_avg_sq: frb_peer_shard_floats(world, rank, n_floats) floats each, zero before
 * the first step.  state: 2 uint32 on this device, zero before the first step ([1] counts the steps taken, so
 * the arguments never change and the launch replays from a CUDA graph).  The summed gradient is multiplied by
 * grad_scale (1 = SUM, 1/world = mean).  Every rank must call it once per step; the call returns when enqueued. */
long long frb_peer_shard_floats(int world, int rank, long long n_floats);
int frb_peer_adam_step(int world, int rank, long long n_floats, const unsigned long long* grad_ptrs,
                       const unsigned long long* param_ptrs, const unsigned long long* signal_ptrs,
                       float* exp_avg, float* exp_avg_sq, uint32_t* state, double lr, double beta1, double beta2,
                       float eps, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FRESNEL_B200_H */
