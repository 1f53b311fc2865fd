// C-ABI driver around the REFERENCE's GaussianCloud file functions (compiled from /root/reference by
// oracle/build_ref.sh into oracle/_ref/libref_cloud_io.so).  Test infrastructure: oracle/make_golden.py calls it to
// write tests/golden/cloud_97.ply with the reference's own save_ply and to read it back with its own load_ply.
// rows: n x 14 floats [position 3 | scale 3 | rotation wxyz 4 | colour 3 | opacity 1] (activated values).
#include <cstring>
#include <string>

#include "gaussian.hpp"

using fresnel::Gaussian3D;
using fresnel::GaussianCloud;

static GaussianCloud from_rows(const float* rows, int n) {
    GaussianCloud c;
    for (int i = 0; i < n; ++i) {
        const float* r = rows + 14 * i;
        c.add(Gaussian3D(glm::vec3(r[0], r[1], r[2]), glm::vec3(r[3], r[4], r[5]), glm::quat(r[6], r[7], r[8], r[9]),
                         glm::vec3(r[10], r[11], r[12]), r[13]));
    }
    return c;
}

static int to_rows(const GaussianCloud& c, float* rows, int cap) {
    const int n = (int)c.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const Gaussian3D& g = c[i];
        float* r = rows + 14 * i;
        r[0] = g.position.x; r[1] = g.position.y; r[2] = g.position.z;
        r[3] = g.scale.x; r[4] = g.scale.y; r[5] = g.scale.z;
        r[6] = g.rotation.w; r[7] = g.rotation.x; r[8] = g.rotation.y; r[9] = g.rotation.z;
        r[10] = g.color.r; r[11] = g.color.g; r[12] = g.color.b;
        r[13] = g.opacity;
    }
    return n;
}

extern "C" {
int ref_save_ply(const float* rows, int n, const char* path) { return from_rows(rows, n).save_ply(path) ? 0 : 1; }
int ref_save_binary(const float* rows, int n, const char* path) { return from_rows(rows, n).save_binary(path) ? 0 : 1; }
// returns the number of Gaussians in the file (rows receives at most cap of them), -1 on failure
int ref_load_ply(const char* path, float* rows, int cap) {
    GaussianCloud c;
    if (!c.load_ply(path)) return -1;
    return to_rows(c, rows, cap);
}
int ref_load_binary(const char* path, float* rows, int cap) {
    GaussianCloud c;
    if (!c.load_binary(path)) return -1;
    return to_rows(c, rows, cap);
}
}
