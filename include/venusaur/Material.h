// Material.h -- host-side material description; same public interface as the reference's Core/material.h:6-33.
#pragma once

#include "Vec3.h"

namespace venusaur {

class Material {
public:
    enum Type { Lambertian = 0, Metal, Dielectric };   // material.h:9-14; values match VN_LAMBERTIAN/METAL/DIELECTRIC

    Material(Type type, const vec3& albedo = vec3(0.0f), float fuzz = 0, float ir = 0)
        : m_type(type), m_albedo(albedo), m_fuzz(fuzz), m_ir(ir) {}

    inline Type GetType() const { return m_type; }
    inline float3_t GetAlbedo() const { return float3_t{m_albedo.r, m_albedo.g, m_albedo.b}; }
    inline float GetFuzz() const { return m_fuzz; }
    inline float GetIR() const { return m_ir; }

private:
    Type m_type;
    vec3 m_albedo;
    float m_fuzz;
    float m_ir;
};

}  // namespace venusaur
