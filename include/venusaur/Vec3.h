// Vec3.h -- minimal glm-free float3 for the host-side drop-in classes.  The reference uses glm::vec3
// (material.h:4, camera.h:3-4, Scene.h) which is not vendored; only .x/.y/.z (.r/.g/.b), + - * and
// normalize/cross/dot/length are needed by Scene, Sphere, Material and Camera.
#pragma once

#include <cmath>

namespace venusaur {

struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };

    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    // glm converts every constructor argument to float, whatever arithmetic produced it (Scene.h:25 mixes int,
    // double and float), so the parameters are double here and narrowed once.
    vec3(double x_, double y_, double z_) : x(static_cast<float>(x_)), y(static_cast<float>(y_)), z(static_cast<float>(z_)) {}

    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

inline vec3 operator+(const vec3& a, const vec3& b) { vec3 o; o.x = a.x + b.x; o.y = a.y + b.y; o.z = a.z + b.z; return o; }
inline vec3 operator-(const vec3& a, const vec3& b) { vec3 o; o.x = a.x - b.x; o.y = a.y - b.y; o.z = a.z - b.z; return o; }
inline vec3 operator*(const vec3& a, const vec3& b) { vec3 o; o.x = a.x * b.x; o.y = a.y * b.y; o.z = a.z * b.z; return o; }
inline vec3 operator*(const vec3& a, float s) { vec3 o; o.x = a.x * s; o.y = a.y * s; o.z = a.z * s; return o; }
inline vec3 operator*(float s, const vec3& a) { return a * s; }
inline vec3 operator-(const vec3& a) { vec3 o; o.x = -a.x; o.y = -a.y; o.z = -a.z; return o; }

// glm::dot / glm::length / glm::normalize / glm::cross, float formulas (detail/func_geometric.inl)
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 cross(const vec3& a, const vec3& b) {
    vec3 o;
    o.x = a.y * b.z - b.y * a.z;
    o.y = a.z * b.x - b.z * a.x;
    o.z = a.x * b.y - b.x * a.y;
    return o;
}

// float3 as returned by the reference's getters (sphere.h:13, material.h:20); layout-compatible with CUDA's float3.
struct float3_t { float x, y, z; };

// OptixAabb stand-in (Scene.h:83, sphere.h:17-28): six floats.
struct Aabb { float minX, minY, minZ, maxX, maxY, maxZ; };

}  // namespace venusaur
