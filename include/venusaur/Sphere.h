// Sphere.h -- same public interface as the reference's Core/sphere.h:4-35.
#pragma once

#include <cmath>

#include "Material.h"

namespace venusaur {

class Sphere {
public:
    Sphere(const vec3& center, float radius, const Material& material)
        : m_center(center), m_radius(radius), m_material(material) {}

    inline float3_t GetCenter() const { return float3_t{m_center.x, m_center.y, m_center.z}; }
    inline float GetRadius() const { return m_radius; }
    inline const Material& GetMaterial() const { return m_material; }
    inline void SetMaterial(const Material& material) { m_material = material; }

    // sphere.h:17-28 computes fabsf(radius) and then uses the signed radius, which inverts the box of an
    // RTIOW-style hollow sphere.  Here |r| is used; identical for every sphere Scene() makes.
    inline Aabb GetAABB() const {
        const float r = std::fabs(m_radius);
        return Aabb{m_center.x - r, m_center.y - r, m_center.z - r, m_center.x + r, m_center.y + r, m_center.z + r};
    }

private:
    vec3 m_center;
    float m_radius;
    Material m_material;
};

}  // namespace venusaur
