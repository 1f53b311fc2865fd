// Whole-pass entry points of the tile renderer: one C call enqueues every stage of the forward
// (or backward) pass on the stream, carving its buffers out of two caller-provided arenas.
// This is the capacity (sync-free) mode of DESIGN.md section 3: instance buffers are sized for
// m_capacity entries, the true count stays on the device, and nothing here touches the host.
// The stage functions it sequences are the ones declared in include/fresnel_b200.h.
#include "frb_common.cuh"

#include <stdlib.h>

#include <mutex>
#include <vector>

namespace {

// ---- per-stage CUDA events INSIDE the whole-pass entry points (bench.py's roofline line) -----------------------
// With timing enabled, frb_tile_render_fwd / _bwd bracket every stage they enqueue with a pair of events on the
// caller's stream, so the stage times come from the same code path (same kernels, same buffers, same fused binning)
// as the throughput figure.  Not for use inside a stream capture (events recorded in a capture cannot be timed).
struct StageRec {
    const char* name;
    cudaEvent_t a, b;
};
std::mutex g_stage_mutex;
std::vector<StageRec> g_stages;
bool g_stage_timing = false;

struct StageScope {
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    const char* name;
    bool on;
    StageScope(const char* n, void* stream) : st((cudaStream_t)stream), name(n), on(g_stage_timing) {
        if (!on) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~StageScope() {
        if (!on) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lock(g_stage_mutex);
        g_stages.push_back(StageRec{name, a, b});
    }
};
#define FRB_STAGE(name, stream, expr)             \
    do {                                          \
        StageScope _scope(name, stream);          \
        if ((rc = (expr))) return rc;             \
    } while (0)

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// tile lists by counting + bitmap ranking (tile_lists.cu) when the rank bitmap and the tile counters fit an SM
bool frb_use_tile_lists(int n, int tiles) {
    return n > 0 && n <= frb_tile_lists_max_gaussians() && tiles <= frb_tile_lists_max_tiles();
}

// The compositor kernels of the whole-pass calls fetch the records of a tile's list by TMA tile::gather4 from the
// unsorted record array (frb_composite_fwd_gather / _bwd_gather) instead of reading a sorted copy: no 48-byte-per-
// instance gather pass, no sorted_records buffer (it was 480 MB of the 560 MB worst-case arena at config 2).
// Measured on one box: 2579 frames/s against 2494 with the sorted copy (frb_tile_rank_gather 29 -> 19 us, the
// backward compositor 188 -> 198 us: its two lane-parallel record reads per 32-entry block conflict 16-way on the
// 64-byte rows the gather needs).  FRB_GATHER=0 selects the sorted copy (A/B).
bool frb_gather_mode() {
    static const bool on = !(getenv("FRB_GATHER") && atoi(getenv("FRB_GATHER")) == 0);
    return on;
}

struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += align256(bytes);
        return o;
    }
};

}  // namespace

extern "C" int frb_stage_timing_enable(int on) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    for (auto& r : g_stages) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_stages.clear();
    g_stage_timing = on != 0;
    return 0;
}

extern "C" int frb_stage_timing_count(void) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    return (int)g_stages.size();
}

// Name and duration of recorded stage i (waits for its closing event).
extern "C" int frb_stage_timing_get(int i, const char** name, float* ms) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    if (i < 0 || i >= (int)g_stages.size() || !name || !ms) return FRB_E_INVALID;
    FRB_CUDA_OK(cudaEventSynchronize(g_stages[i].b));
    FRB_CUDA_OK(cudaEventElapsedTime(ms, g_stages[i].a, g_stages[i].b));
    *name = g_stages[i].name;
    return 0;
}

// Byte offsets inside the arenas.  persist: needed again by the backward pass.  scratch: forward only.
extern "C" int frb_tile_layout(int n, int n_views, int width, int height, int m_capacity, FrbTileLayout* L) {
    if (n < 0 || n_views < 1 || width < 1 || height < 1 || m_capacity < 0 || !L) return FRB_E_INVALID;
    const size_t hw = (size_t)width * height * n_views;
    const size_t tiles = (size_t)n_views * frb_div_up(width, FRB_TILE) * frb_div_up(height, FRB_TILE);
    const size_t m = (size_t)m_capacity;
    Carver p;
    L->ranges = p.take(8 * tiles);
    L->tile_order = p.take(4 * tiles);
    L->state_T = p.take(4 * hw);
    L->state_n = p.take(4 * hw);
    L->sorted_gids = p.take(4 * m);
    // the sorted record copy exists only on the key-sort path and with FRB_GATHER=0
    L->sorted_records = p.take((frb_gather_mode() && frb_use_tile_lists(n, (int)tiles)) ? 0 : 48 * (m + 1));
    L->records = p.take(48 * ((size_t)n + 1));      // persist: the backward pass gathers from it too
    L->persist_bytes = p.off;
    Carver s;
    L->depth_bits = s.take(4 * (size_t)n);
    L->touched = s.take(4 * (size_t)n);
    L->order = s.take(4 * (size_t)n);
    L->rank = s.take(4 * (size_t)n);
    L->offsets = s.take(4 * ((size_t)n + 2));      // [n]: instance count, [n + 1]: status (frb_tile_scan)
    L->depth_ws = s.take(frb_depth_order_workspace_bytes(n));
    L->scan_ws = s.take(frb_scan_workspace_bytes(n));
    if (frb_use_tile_lists(n, (int)tiles)) {
        // tile lists by counting + bitmap ranking: 4 bytes of scratch per instance, no key buffers
        L->keys = L->keys_tmp = L->vals_tmp = L->sort_ws = s.off;
        L->tile_ws = s.take(frb_tile_lists_workspace_bytes(n, (int)tiles));
        L->inst_rank = s.take(4 * m);
    } else {
        L->keys = s.take(8 * m);
        L->keys_tmp = s.take(8 * m);
        L->vals_tmp = s.take(4 * m);
        L->sort_ws = s.take(frb_sort_workspace_bytes(m_capacity));
        L->tile_ws = L->inst_rank = s.off;
    }
    L->scratch_bytes = s.off;
    return 0;
}

extern "C" int frb_tile_render_fwd(int n, int n_views, const float* positions, const float* scales,
                                   const float* rotations, const float* colors, const float* opacities,
                                   const float* camera_host, float max_radius, int width, int height,
                                   const float* background_host, float t_eps, int m_capacity, void* persist,
                                   void* scratch, float* image, float* depth, float* alpha, void* stream) {
    FrbTileLayout L;
    int rc = frb_tile_layout(n, n_views, width, height, m_capacity, &L);
    if (rc) return rc;
    if (!persist || !scratch) return FRB_E_INVALID;
    char* P = (char*)persist;
    char* S = (char*)scratch;
    const int tiles = n_views * frb_div_up(width, FRB_TILE) * frb_div_up(height, FRB_TILE);
    int tile_bits = 1;
    while ((1 << tile_bits) < tiles) ++tile_bits;
    float* records = (float*)(P + L.records);
    uint32_t* depth_bits = (uint32_t*)(S + L.depth_bits);
    uint32_t* touched = (uint32_t*)(S + L.touched);
    uint32_t* order = (uint32_t*)(S + L.order);
    uint32_t* offsets = (uint32_t*)(S + L.offsets);
    uint64_t* keys = (uint64_t*)(S + L.keys);
    uint32_t* gids = (uint32_t*)(P + L.sorted_gids);
    int32_t* ranges = (int32_t*)(P + L.ranges);
    float* sorted_records = (float*)(P + L.sorted_records);

    FRB_STAGE("frb_project_fwd", stream,
              frb_project_fwd(n, n_views, positions, scales, rotations, colors, opacities, camera_host, max_radius,
                              records, nullptr, depth_bits, touched, nullptr, stream));
    int32_t* tile_order = (int32_t*)(P + L.tile_order);
    if (m_capacity > 0 && frb_use_tile_lists(n, tiles)) {
        // count per tile -> scan (ranges, launch order) -> depth sort -> emit depth ranks -> per-tile bitmap ranking + gather
        uint32_t* inst_rank = (uint32_t*)(S + L.inst_rank);
        uint32_t* rank = (uint32_t*)(S + L.rank);
        // (counting and scanning need no depth order; running them on a forked stream beside the depth sort was
        // measured and dropped: 2564-2577 frames/s against 2586-2588 in stream order - inside the replayed graph the
        // cross-stream edges cost the programmatic-launch overlap they replace, profiles/r2_e_overlap_ab.txt)
        // depth order over the camera's [near, far] only (fewer radix passes for narrow slabs): widest range over the views
        float near_d = camera_host[18], far_d = camera_host[19];
        for (int v = 1; v < n_views; ++v) {
            near_d = fminf(near_d, camera_host[v * FRB_CAMERA_FLOATS + 18]);
            far_d = fmaxf(far_d, camera_host[v * FRB_CAMERA_FLOATS + 19]);
        }
        FRB_STAGE("frb_depth_order", stream,
                  frb_depth_order_range(n, depth_bits, near_d, far_d, order, rank, S + L.depth_ws, stream));
        FRB_STAGE("frb_tile_count_scan", stream,
                  frb_tile_count_scan(n, n_views, width, height, records, m_capacity, ranges, tile_order, offsets + n,
                                      frb_depth_order_error_word(n, S + L.depth_ws), S + L.tile_ws, stream));
        FRB_STAGE("frb_tile_emit", stream,
                  frb_tile_emit(n, n_views, width, height, records, rank, m_capacity, S + L.tile_ws, inst_rank,
                                stream));
        FRB_STAGE("frb_tile_rank_gather", stream,
                  frb_tile_rank_gather(n, tiles, tile_order, ranges, inst_rank, order, records, nullptr, nullptr, gids,
                                       frb_gather_mode() ? nullptr : sorted_records, nullptr, nullptr, stream));
    } else if (n > 0 && m_capacity > 0) {
        FRB_STAGE("frb_depth_order", stream, frb_depth_order(n, depth_bits, order, S + L.depth_ws, stream));
        if (tile_bits <= 16) {
            // scan of the tile counts, key emission and the sort histograms in one kernel (offsets[n] = M only)
            FRB_STAGE("frb_bin_sort_dev", stream,
                      frb_bin_sort_dev(n, n_views, width, height, records, depth_bits, touched, order, m_capacity,
                                       offsets + n, keys, gids, (uint64_t*)(S + L.keys_tmp),
                                       (uint32_t*)(S + L.vals_tmp), tile_bits, S + L.scan_ws, S + L.sort_ws, stream));
        } else {
            if ((rc = frb_tile_offsets(n, touched, order, offsets, S + L.scan_ws, stream))) return rc;
            if ((rc = frb_bin_emit(n, n_views, width, height, records, depth_bits, order, offsets, keys, gids,
                                   stream))) return rc;
            if ((rc = frb_radix_sort_pairs_dev(m_capacity, offsets + n, keys, gids, (uint64_t*)(S + L.keys_tmp),
                                               (uint32_t*)(S + L.vals_tmp), 32, 32 + tile_bits, S + L.sort_ws,
                                               stream))) return rc;
        }
        FRB_STAGE("frb_ranges_and_gather", stream,
                  frb_ranges_and_gather_dev(m_capacity, offsets + n, keys, gids, tiles, ranges, records,
                                            sorted_records, nullptr, nullptr, stream));
        FRB_STAGE("frb_tile_schedule", stream, frb_tile_schedule(tiles, ranges, tile_order, stream));
    } else {
        FRB_CUDA_OK(cudaMemsetAsync(ranges, 0, sizeof(int32_t) * 2 * (size_t)tiles, (cudaStream_t)stream));
        FRB_STAGE("frb_tile_schedule", stream, frb_tile_schedule(tiles, ranges, tile_order, stream));
    }
    if (frb_gather_mode() && n > 0 && m_capacity > 0 && frb_use_tile_lists(n, tiles)) {
        FRB_STAGE("frb_composite_fwd", stream,
                  frb_composite_fwd_gather(n_views, width, height, tile_order, ranges, records, n, gids, background_host,
                                           t_eps, FRB_ALPHA_MAX, image, depth, alpha, (float*)(P + L.state_T),
                                           (int32_t*)(P + L.state_n), stream));
        return 0;
    }
    FRB_STAGE("frb_composite_fwd", stream,
              frb_composite_fwd_sched(n_views, width, height, tile_order, ranges, sorted_records, nullptr, 0.0f,
                                      background_host, t_eps, image, depth, alpha, (float*)(P + L.state_T),
                                      (int32_t*)(P + L.state_n), nullptr, stream));
    return 0;
}

// grad2d: scratch of n * FRB_GRAD_FLOATS floats (zeroed here).
extern "C" int frb_tile_render_bwd(int n, int n_views, const float* positions, const float* scales,
                                   const float* rotations, const float* camera_host, int width, int height,
                                   const float* background_host, int m_capacity, const void* persist,
                                   const float* g_image, const float* g_depth, const float* g_alpha, float* grad2d,
                                   float* g_positions, float* g_scales, float* g_rotations, float* g_colors,
                                   float* g_opacities, void* stream) {
    FrbTileLayout L;
    int rc = frb_tile_layout(n, n_views, width, height, m_capacity, &L);
    if (rc) return rc;
    if (!persist || !grad2d) return FRB_E_INVALID;
    const char* P = (const char*)persist;
    FRB_CUDA_OK(cudaMemsetAsync(grad2d, 0, sizeof(float) * FRB_GRAD_FLOATS * (size_t)n, (cudaStream_t)stream));
    const int tiles_b = n_views * frb_div_up(width, FRB_TILE) * frb_div_up(height, FRB_TILE);
    if (frb_gather_mode() && n > 0 && m_capacity > 0 && frb_use_tile_lists(n, tiles_b)) {
        FRB_STAGE("frb_composite_bwd", stream,
                  frb_composite_bwd_gather(n_views, width, height, (const int32_t*)(P + L.tile_order),
                                           (const int32_t*)(P + L.ranges), (const float*)(P + L.records), n,
                                           (const uint32_t*)(P + L.sorted_gids), background_host, FRB_ALPHA_MAX,
                                           (const float*)(P + L.state_T), (const int32_t*)(P + L.state_n), g_image,
                                           g_depth, g_alpha, grad2d, stream));
    } else
    FRB_STAGE("frb_composite_bwd", stream,
              frb_composite_bwd_sched(n_views, width, height, (const int32_t*)(P + L.tile_order),
                                      (const int32_t*)(P + L.ranges), (const float*)(P + L.sorted_records),
                                      (const uint32_t*)(P + L.sorted_gids), nullptr, 0.0f, background_host,
                                      (const float*)(P + L.state_T), (const int32_t*)(P + L.state_n), nullptr,
                                      g_image, g_depth, g_alpha, grad2d, nullptr, stream));
    FRB_STAGE("frb_project_bwd", stream,
              frb_project_bwd(n, n_views, positions, scales, rotations, camera_host, grad2d, g_positions, g_scales,
                              g_rotations, g_colors, g_opacities, stream));
    return 0;
}
