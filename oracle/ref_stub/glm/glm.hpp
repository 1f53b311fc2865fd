// Minimal stand-in for GLM 1.0.1 (absent from this image and un-vendored in the reference: CMake FetchContent),
// just enough for /root/reference/src/core/renderer/gaussian.hpp to compile so that the reference's OWN
// GaussianCloud::save_ply / load_ply / save_binary / load_binary (renderer.cpp) can be built into oracle/_ref by
// oracle/build_ref.sh and used to pin the file formats.  Test infrastructure only; none of the functions below take
// part in the file I/O code paths (those only read and write the x / y / z / w / r / g / b members).
#pragma once
#include <cmath>

namespace glm {

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
};

struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
};

struct mat3 {
    float m[3][3];      // column-major like GLM: m[col][row]
    mat3() : m{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}} {}
    mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2)
        : m{{a0, a1, a2}, {b0, b1, b2}, {c0, c1, c2}} {}
};

inline mat3 operator*(const mat3& a, const mat3& b) {
    mat3 r;
    for (int c = 0; c < 3; ++c)
        for (int row = 0; row < 3; ++row) {
            float s = 0;
            for (int k = 0; k < 3; ++k) s += a.m[k][row] * b.m[c][k];
            r.m[c][row] = s;
        }
    return r;
}

inline mat3 transpose(const mat3& a) {
    mat3 r;
    for (int c = 0; c < 3; ++c)
        for (int row = 0; row < 3; ++row) r.m[c][row] = a.m[row][c];
    return r;
}

template <typename T>
T identity();

}  // namespace glm
