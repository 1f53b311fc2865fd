// See ../glm.hpp: minimal stand-in for glm::quat (w, x, y, z members, identity, mat3_cast).
#pragma once
#include "../glm.hpp"

namespace glm {

struct quat {
    float w, x, y, z;
    quat() : w(1), x(0), y(0), z(0) {}
    quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_) {}
};

template <>
inline quat identity<quat>() { return quat(1, 0, 0, 0); }

inline mat3 mat3_cast(const quat& q) {
    const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
    const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    return mat3(1 - 2 * (yy + zz), 2 * (xy + wz), 2 * (xz - wy), 2 * (xy - wz), 1 - 2 * (xx + zz), 2 * (yz + wx),
                2 * (xz + wy), 2 * (yz - wx), 1 - 2 * (xx + yy));
}

}  // namespace glm
