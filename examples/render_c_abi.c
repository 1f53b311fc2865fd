/* Torch-free use of the C-ABI (include/fresnel_b200.h): forward + backward of the tile renderer from plain C.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/render_c_abi.c -o render_c_abi \
 *       fresnel_b200/csrc/libfresnel_b200.so -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/fresnel_b200/csrc -Wl,-rpath,/usr/local/cuda/lib64
 *   ./render_c_abi in.bin out.bin
 *
 * in.bin : int32 n, W, H ; float camera[20] ; float bg[3] ; then positions[n*3], scales[n*3], rotations[n*4],
 *          colors[n*3], opacities[n], g_image[3*H*W], g_depth[H*W]            (all float32, host order)
 * out.bin: image[3*H*W], depth[H*W], alpha[H*W], g_positions[n*3], g_scales[n*3], g_rotations[n*4],
 *          g_colors[n*3], g_opacities[n]
 * This is what a maintainer binding the library from C / C++ / cgo / JNI would write; tests/test_gpu_parity.py builds
 * and runs it and compares out.bin with the nn.Module path (bit-identical image: same kernels). */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fresnel_b200.h"

#define CK(e)                                                                              \
    do {                                                                                   \
        int _rc = (int)(e);                                                                \
        if (_rc != 0) {                                                                    \
            fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #e, _rc,         \
                    _rc > 0 ? cudaGetErrorString((cudaError_t)_rc) : frb_error_string(_rc)); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static float* upload(const float* host, size_t count) {
    float* d = NULL;
    if (cudaMalloc((void**)&d, sizeof(float) * (count ? count : 1)) != cudaSuccess) return NULL;
    if (count && cudaMemcpy(d, host, sizeof(float) * count, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    return d;
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[3];
    float cam[20], bg[3];
    if (fread(hdr, sizeof(int), 3, f) != 3 || fread(cam, sizeof(float), 20, f) != 20 ||
        fread(bg, sizeof(float), 3, f) != 3)
        return 2;
    const int n = hdr[0], W = hdr[1], H = hdr[2];
    const size_t hw = (size_t)W * H;
    const size_t n_in = (size_t)n * 14 + 4 * hw;
    float* in = (float*)malloc(sizeof(float) * n_in);
    if (!in || fread(in, sizeof(float), n_in, f) != n_in) return 2;
    fclose(f);
    const float *pos = in, *scl = pos + 3 * (size_t)n, *rot = scl + 3 * (size_t)n, *col = rot + 4 * (size_t)n,
                *opa = col + 3 * (size_t)n, *g_img = opa + n, *g_dep = g_img + 3 * hw;

    /* capacity: a rectangle is at most 2 * max_radius + 2 pixels wide */
    const float max_radius = 64.0f, t_eps = 0.0f;
    const int tiles = ((W + FRB_TILE - 1) / FRB_TILE) * ((H + FRB_TILE - 1) / FRB_TILE);
    const int span = (2 * 64 + 2 + FRB_TILE - 1) / FRB_TILE + 1;
    const int cap = n * (tiles < span * span ? tiles : span * span);
    FrbTileLayout L;
    CK(frb_tile_layout(n, 1, W, H, cap, &L));

    float *d_pos = upload(pos, 3 * (size_t)n), *d_scl = upload(scl, 3 * (size_t)n), *d_rot = upload(rot, 4 * (size_t)n),
          *d_col = upload(col, 3 * (size_t)n), *d_opa = upload(opa, n), *d_gi = upload(g_img, 3 * hw),
          *d_gd = upload(g_dep, hw);
    void *persist = NULL, *scratch = NULL;
    float *out = NULL, *grads = NULL;
    CK(cudaMalloc(&persist, L.persist_bytes ? L.persist_bytes : 256));
    CK(cudaMalloc(&scratch, L.scratch_bytes ? L.scratch_bytes : 256));
    CK(cudaMalloc((void**)&out, sizeof(float) * 5 * hw));
    CK(cudaMalloc((void**)&grads, sizeof(float) * 26 * (size_t)(n ? n : 1)));   /* grad2d 12 n | 14 n gradients */
    if (!d_pos || !d_scl || !d_rot || !d_col || !d_opa || !d_gi || !d_gd) return 3;
    cudaStream_t st;
    CK(cudaStreamCreate(&st));

    CK(frb_tile_render_fwd(n, 1, d_pos, d_scl, d_rot, d_col, d_opa, cam, max_radius, W, H, bg, t_eps, cap, persist,
                           scratch, out, out + 3 * hw, out + 4 * hw, st));
    float* g2 = grads;
    float *g_rot = grads + 12 * (size_t)n, *g_pos = g_rot + 4 * (size_t)n, *g_scl = g_pos + 3 * (size_t)n,
          *g_col = g_scl + 3 * (size_t)n, *g_opa = g_col + 3 * (size_t)n;
    CK(frb_tile_render_bwd(n, 1, d_pos, d_scl, d_rot, cam, W, H, bg, cap, persist, d_gi, d_gd, NULL, g2, g_pos, g_scl,
                           g_rot, g_col, g_opa, st));
    CK(cudaStreamSynchronize(st));

    const size_t n_out = 5 * hw + 14 * (size_t)n;
    float* host = (float*)malloc(sizeof(float) * n_out);
    CK(cudaMemcpy(host, out, sizeof(float) * 5 * hw, cudaMemcpyDeviceToHost));
    float* hg = host + 5 * hw;
    CK(cudaMemcpy(hg, g_pos, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hg + 3 * (size_t)n, g_scl, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hg + 6 * (size_t)n, g_rot, sizeof(float) * 4 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hg + 10 * (size_t)n, g_col, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hg + 13 * (size_t)n, g_opa, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    f = fopen(argv[2], "wb");
    if (!f || fwrite(host, sizeof(float), n_out, f) != n_out) return 2;
    fclose(f);
    double s = 0.0;
    for (size_t i = 0; i < 3 * hw; ++i) s += host[i];
    printf("render_c_abi: n=%d %dx%d image sum %.6f, library launches %llu\n", n, W, H, s, frb_launch_count());
    return 0;
}
