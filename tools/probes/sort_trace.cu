// Probe (not product code): where does a one-sweep radix pass over 100k depth keys spend its ~8 us?  Builds sort.cu
// with FRB_SORT_TRACE (thread 0 of every tile stamps %globaltimer at the phase boundaries) and prints, per pass, the
// spread of those stamps over the tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFRB_SORT_TRACE -I include -I fresnel_b200/csrc \
//        -o tools/probes/_sort_trace tools/probes/sort_trace.cu -lcuda
#include "../../fresnel_b200/csrc/sort.cu"

bool frb_pdl_enabled() { return true; }      // the two hooks sort.cu expects from pipeline.cu
void frb_note_launches(int) {}

#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 100000;
    std::vector<float> depth(n);
    std::mt19937 rng(0);
    std::normal_distribution<float> nd(2.0f, 0.5f);
    for (auto& d : depth) d = std::max(0.02f, nd(rng));
    uint32_t *d_bits, *d_order, *d_rank;
    void* ws;
    cudaMalloc(&d_bits, 4 * n); cudaMalloc(&d_order, 4 * n); cudaMalloc(&d_rank, 4 * n);
    cudaMalloc(&ws, frb_depth_order_workspace_bytes(n));
    cudaMemcpy(d_bits, depth.data(), 4 * n, cudaMemcpyHostToDevice);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 10; ++it) {
        cudaEventRecord(e0, st);
        int rc = frb_depth_order_range(n, d_bits, 0.01f, 100.0f, d_order, d_rank, ws, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        if (rc) { printf("rc=%d\n", rc); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    printf("n=%d depth order (memset + hist + passes), stream launches: best %.1f us\n", n, best * 1e3f);
    // the same chain replayed from a CUDA graph (no CPU launch cost between the nodes); the trace is the last replay's
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    frb_depth_order_range(n, d_bits, 0.01f, 100.0f, d_order, d_rank, ws, st);
    cudaStreamEndCapture(st, &graph);
    cudaGraphInstantiate(&exec, graph, 0);
    best = 1e9f;
    for (int it = 0; it < 10; ++it) {
        cudaEventRecord(e0, st);
        cudaGraphLaunch(exec, st);
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    printf("n=%d the same chain as one graph replay: best %.1f us\n", n, best * 1e3f);
    {
        // the cluster-resident sort (what frb_depth_order_range runs for n <= 131,072 unless FRB_CLUSTER_SORT=0)
        static unsigned long long ct[CS_MAX_CTAS][6][10];
        cudaMemcpyFromSymbol(ct, frb_cluster_trace, sizeof(ct));
        unsigned long long c0 = ~0ull;
        for (int c = 0; c < CS_MAX_CTAS; ++c) if (ct[c][0][0]) c0 = std::min(c0, ct[c][0][0]);
        if (c0 != ~0ull) {
            const char* cn[8] = {"pass start", "ranked", "counts sent", "cluster.sync", "bases", "staged", "sent", "cluster.sync"};
            for (int c = 0; c < CS_MAX_CTAS; c += 5) {
                printf(" cluster sort, CTA %d (ns after the first CTA started): load begin %lld, loaded %lld, first cluster.sync %lld\n",
                       c, (long long)(ct[c][0][0] - c0), (long long)(ct[c][0][1] - c0), (long long)(ct[c][0][2] - c0));
                for (int p = 1; p <= 4; ++p) {
                    printf("   pass %d:", p - 1);
                    for (int k = 0; k < 8; ++k) printf(" %s %lld;", cn[k], (long long)(ct[c][p][k] - c0));
                    printf("\n");
                }
                printf("   stored %lld\n", (long long)(ct[c][5][0] - c0));
            }
        }
    }
    static unsigned long long tr[4][1024][8];
    cudaMemcpyFromSymbol(tr, frb_sort_trace, sizeof(tr));
    const int tiles = std::min(1024, (n + 511) / 512);
    unsigned long long t0 = ~0ull;
    for (int t = 0; t < tiles; ++t) t0 = std::min(t0, tr[0][t][0]);
    const char* names[7] = {"entry", "ticket", "keys requested", "ranked", "published", "look-back done", "scattered"};
    for (int p = 0; p < 4; ++p) {
        printf(" pass %d (ns after the first block of pass 0 entered)\n", p);
        for (int s = 0; s < 7; ++s) {
            double mn = 1e18, mx = 0, sum = 0;
            for (int t = 1; t < tiles; ++t) {
                double v = (double)(long long)(tr[p][t][s] - t0);
                mn = std::min(mn, v); mx = std::max(mx, v); sum += v;
            }
            printf("  %-16s min %7.0f  mean %7.0f  max %7.0f\n", names[s], mn, sum / (tiles - 1), mx);
        }
        double rsum = 0, rmax = 0;
        for (int t = 1; t < tiles; ++t) { rsum += tr[p][t][7]; rmax = std::max(rmax, (double)tr[p][t][7]); }
        printf("  look-back rounds per tile: mean %.1f max %.0f (tiles %d)\n", rsum / (tiles - 1), rmax, tiles);
    }
    return 0;
}
