// SimplifiedRenderer: isotropic point splats with an integer pixel radius, composited with "over".
// Reference: SimplifiedRenderer.forward DR:1347-1458, Camera.project DR:54-85.
// The reference blends back to front (image = alpha * colour + (1 - alpha) * image, DR:1433-1436); that is the
// same sum as front-to-back compositing with transmittance, so the splats run through the tile compositor
// (composite.cu) with the clamp(alpha, 0, 1) of DR:1430 as alpha_max = 1 - 2^-24.  This file holds what differs:
// the projection / radius / rectangle rule, the depth map (front-most entry with alpha > 0.1, DR:1438-1442)
// and the chain of (u, v, depth) back to the positions.
#include "frb_common.cuh"

namespace {

struct SimpleProj {
    float pcx, pcy, pcz, zs, u, v, depth;
};

// Camera.project DR:54-85, fixed left-to-right fp32 order (bit-exact against the oracle restatement)
__device__ __forceinline__ SimpleProj simple_project(const float p[3], const FrbCamera& cam) {
    const float* V = cam.V;
    SimpleProj o;
    o.pcx = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[0], p[0]), FRB_MUL(V[1], p[1])), FRB_MUL(V[2], p[2])), V[3]);
    o.pcy = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[4], p[0]), FRB_MUL(V[5], p[1])), FRB_MUL(V[6], p[2])), V[7]);
    o.pcz = FRB_ADD(FRB_ADD(FRB_ADD(FRB_MUL(V[8], p[0]), FRB_MUL(V[9], p[1])), FRB_MUL(V[10], p[2])), V[11]);
    o.zs = FRB_MUL(frb_clamp_min(fabsf(o.pcz), cam.near_), frb_sign(FRB_ADD(o.pcz, 1e-8f)));   // DR:78
    o.u = FRB_ADD(FRB_DIV(FRB_MUL(cam.fx, o.pcx), -o.zs), cam.cx);                              // DR:81
    o.v = FRB_ADD(FRB_DIV(FRB_MUL(cam.fy, -o.pcy), -o.zs), cam.cy);                             // DR:82
    o.depth = -o.zs;                                                                            // DR:85
    return o;
}

__global__ void __launch_bounds__(256)
simple_project_fwd_kernel(int n, const __grid_constant__ FrbViewSet vs, const float* __restrict__ positions,
                          const float* __restrict__ scales, const float* __restrict__ colors,
                          const float* __restrict__ opacities, float4* __restrict__ records,
                          uint32_t* __restrict__ depth_bits, uint32_t* __restrict__ tiles_touched) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FrbCamera& cam = vs.cam[i / vs.n_per_view];
    const float p[3] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
    const SimpleProj o = simple_project(p, cam);
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, radius = 1;
    bool vis = o.depth > 0.0f;                                                      // DR:1398: "if d <= 0: continue"
    if (vis) {
        // DR:1402: int(max(scales.mean().item() * fx / d, 1)) - the product of two Python floats is fp64, the
        // division by the 0-dim fp32 tensor d is tensor.reciprocal() * scalar in fp32
        const float mean = FRB_DIV(FRB_ADD(FRB_ADD(scales[3 * i], scales[3 * i + 1]), scales[3 * i + 2]), 3.0f);
        const float prod = (float)((double)mean * (double)cam.fx);
        const float q = FRB_MUL(FRB_DIV(1.0f, o.depth), prod);
        radius = (1.0f > q) ? 1 : (int)fminf(q, 1.0e6f);
        radius = min(radius, 20);                                                   // DR:1403
        const double ud = (double)o.u, vd = (double)o.v;
        if (!(fabs(ud) < 1.0e9 && fabs(vd) < 1.0e9)) {
            vis = false;
        } else {
            const long long xi = (long long)ud, yi = (long long)vd;                 // int(): truncation towards zero
            const long long W = (long long)cam.width, H = (long long)cam.height;
            x0 = (int)max(0ll, xi - radius); x1 = (int)min(W, xi + radius + 1);     // DR:1407-1410
            y0 = (int)max(0ll, yi - radius); y1 = (int)min(H, yi + radius + 1);
            if (x0 >= x1 || y0 >= y1) vis = false;                                  // DR:1412
        }
    }
    uint32_t touched = 0;
    if (vis) {
        const int tx0 = x0 / FRB_TILE, tx1 = (x1 - 1) / FRB_TILE, ty0 = y0 / FRB_TILE, ty1 = (y1 - 1) / FRB_TILE;
        touched = (uint32_t)((tx1 - tx0 + 1) * (ty1 - ty0 + 1));
    } else {
        x0 = x1 = y0 = y1 = 0;
    }
    const float half = fmaxf(0.5f * (float)radius, 1.0f);                           // DR:1424: 2 * max(r/2, 1)^2
    const float A = -FRB_LOG2E / (2.0f * half * half);
    const uint32_t rect_lo = (uint32_t)x0 | ((uint32_t)y0 << 16);
    const uint32_t rect_hi = (uint32_t)x1 | ((uint32_t)y1 << 16) | 0x80008000u;
    records[3 * i + 0] = make_float4(o.u, o.v, A, 0.0f);
    records[3 * i + 1] = make_float4(A, opacities[i], __uint_as_float(rect_lo), __uint_as_float(rect_hi));
    records[3 * i + 2] = make_float4(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2], o.depth);
    depth_bits[i] = __float_as_uint(vis ? o.depth : 0.0f);
    tiles_touched[i] = touched;
}

__global__ void __launch_bounds__(256)
simple_project_bwd_kernel(int n, const __grid_constant__ FrbViewSet vs, const float* __restrict__ positions,
                          const float4* __restrict__ grad2d, float* __restrict__ g_positions,
                          float* __restrict__ g_colors, float* __restrict__ g_opacities) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 g0 = grad2d[3 * i + 0], g1 = grad2d[3 * i + 1], g2 = grad2d[3 * i + 2];
    g_colors[3 * i] = g2.x; g_colors[3 * i + 1] = g2.y; g_colors[3 * i + 2] = g2.z;
    g_opacities[i] = g1.y;
    float gp[3] = {0.f, 0.f, 0.f};
    const float g_u = g0.x, g_v = g0.y, g_d = g1.z;
    if (g_u != 0.f || g_v != 0.f || g_d != 0.f) {
        const FrbCamera& cam = vs.cam[i / vs.n_per_view];
        const float* V = cam.V;
        const float p[3] = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
        const SimpleProj o = simple_project(p, cam);
        const float iz = 1.0f / o.zs, iz2 = iz * iz;
        // u = -fx x / zs + cx ; v = fy y / zs + cy ; depth = -zs
        const float g_pcx = -g_u * cam.fx * iz;
        const float g_pcy = g_v * cam.fy * iz;
        const float g_zs = g_u * cam.fx * o.pcx * iz2 - g_v * cam.fy * o.pcy * iz2 - g_d;
        const float gate = (fabsf(o.pcz) >= cam.near_) ? 1.0f : 0.0f;               // clamp(min) passes inclusive
        const float g_pcz = g_zs * gate * frb_sign(o.pcz) * frb_sign(o.pcz + 1e-8f);
        gp[0] = V[0] * g_pcx + V[4] * g_pcy + V[8] * g_pcz;
        gp[1] = V[1] * g_pcx + V[5] * g_pcy + V[9] * g_pcz;
        gp[2] = V[2] * g_pcx + V[6] * g_pcy + V[10] * g_pcz;
    }
    g_positions[3 * i] = gp[0]; g_positions[3 * i + 1] = gp[1]; g_positions[3 * i + 2] = gp[2];
}

// depth map: the tile list is front to back, so the first entry with alpha > 0.1 has the smallest depth
__global__ void __launch_bounds__(256)
simple_depth_fwd_kernel(int width, int height, int tiles_x, int tiles_per_view, const int2* __restrict__ ranges,
                        const float4* __restrict__ sorted_records, const uint32_t* __restrict__ sorted_gids,
                        float* __restrict__ depth, int* __restrict__ hit) {
    const int tile = blockIdx.x;
    const int view = tile / tiles_per_view;
    const int t_in_view = tile - view * tiles_per_view;
    const int ty = t_in_view / tiles_x, tx = t_in_view - ty * tiles_x;
    const int px = tx * FRB_TILE + (threadIdx.x & (FRB_TILE - 1)), py = ty * FRB_TILE + (threadIdx.x / FRB_TILE);
    const bool in_image = px < width && py < height;
    const int2 range = ranges[tile];
    float d = 0.0f;
    int h = -1;
    bool done = !in_image;
    for (int e = range.x; e < range.y; ++e) {
        if (__syncthreads_and(done ? 1 : 0)) break;
        if (done) continue;
        const float4 r1 = sorted_records[3 * (size_t)e + 1], r2 = sorted_records[3 * (size_t)e + 2];
        const uint32_t lo = __float_as_uint(r1.z), hi = __float_as_uint(r1.w) & 0x7fff7fffu;
        if (px >= (int)(lo & 0xffff) && px < (int)(hi & 0xffff) && py >= (int)(lo >> 16) && py < (int)(hi >> 16)) {
            const float4 r0 = sorted_records[3 * (size_t)e + 0];
            const float dx = (float)px - r0.x, dy = (float)py - r0.y;
            const float a = fminf(fmaxf(frb_ex2(r0.z * (dx * dx + dy * dy)) * r1.y, 0.0f), 1.0f);
            if (a > 0.1f) {                                                         // DR:1439
                d = r2.w;
                h = (int)sorted_gids[e];
                done = true;
            }
        }
    }
    if (in_image) {
        const size_t pix = (size_t)view * width * height + (size_t)py * width + px;
        depth[pix] = d;                                                             // inf -> 0, DR:1453
        hit[pix] = h;
    }
}

__global__ void __launch_bounds__(256)
simple_depth_bwd_kernel(long long total, const int* __restrict__ hit, const float* __restrict__ g_depth,
                        float* __restrict__ grad2d) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int h = hit[t];
    const float g = g_depth[t];
    if (h >= 0 && g != 0.0f) atomicAdd(grad2d + (size_t)h * FRB_GRAD_FLOATS + 6, g);   // slot 6 = d depth
}

}  // namespace

extern "C" int frb_simple_project_fwd(int n, int n_views, const float* positions, const float* scales,
                                      const float* colors, const float* opacities, const float* camera_host,
                                      float* records, uint32_t* depth_bits, uint32_t* tiles_touched, void* stream) {
    FrbViewSet vs;
    int rc = frb_fill_views(n, n_views, camera_host, &vs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!positions || !scales || !colors || !opacities || !records || !depth_bits || !tiles_touched)
        return FRB_E_INVALID;
    simple_project_fwd_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, vs, positions, scales, colors, opacities, (float4*)records, depth_bits, tiles_touched);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_simple_project_bwd(int n, int n_views, const float* positions, const float* camera_host,
                                      const float* grad2d, float* g_positions, float* g_colors, float* g_opacities,
                                      void* stream) {
    FrbViewSet vs;
    int rc = frb_fill_views(n, n_views, camera_host, &vs);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!positions || !grad2d || !g_positions || !g_colors || !g_opacities) return FRB_E_INVALID;
    simple_project_bwd_kernel<<<frb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, vs, positions, (const float4*)grad2d, g_positions, g_colors, g_opacities);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_simple_depth_fwd(int n_views, int width, int height, const int32_t* ranges,
                                    const float* sorted_records, const uint32_t* sorted_gids, float* depth,
                                    int32_t* hit, void* stream) {
    if (n_views < 1 || n_views > FRB_MAX_VIEWS || width < 1 || height < 1) return FRB_E_INVALID;
    if (width > FRB_MAX_IMAGE_SIDE || height > FRB_MAX_IMAGE_SIDE) return FRB_E_TOO_LARGE;
    if (!ranges || !depth || !hit) return FRB_E_INVALID;
    const int tiles_x = frb_div_up(width, FRB_TILE), tpv = tiles_x * frb_div_up(height, FRB_TILE);
    simple_depth_fwd_kernel<<<n_views * tpv, 256, 0, (cudaStream_t)stream>>>(
        width, height, tiles_x, tpv, (const int2*)ranges, (const float4*)sorted_records, sorted_gids, depth, hit);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}

extern "C" int frb_simple_depth_bwd(int n_views, int width, int height, const int32_t* hit, const float* g_depth,
                                    float* grad2d, void* stream) {
    if (n_views < 1 || width < 1 || height < 1 || !hit || !g_depth || !grad2d) return FRB_E_INVALID;
    const long long total = (long long)n_views * width * height;
    simple_depth_bwd_kernel<<<frb_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(total, hit, g_depth, grad2d);
    frb_note_launches(1);
    FRB_LAUNCH_CHECK();
    return 0;
}
