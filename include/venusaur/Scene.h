// Scene.h -- the RTIOW "final scene"; same public members as the reference's Core/Scene.h:10-84
// (m_spheres, m_aabbs, m_indices) plus the synthetic generators of BASELINE.json's large configs.
//
// Determinism: the reference draws several random numbers inside single argument lists
// (Scene.h:25,33,42,106-112) and through std::uniform_real_distribution<float>, so its scene depends on the compiler
// and the standard library.  Here every draw is sequenced explicitly, left to right, and the float mapping is
// libstdc++'s generate_canonical<float,24> for a 32-bit engine: float(g()) * 2^-32, clamped below 1.  std::mt19937
// itself is fully specified by the standard (default seed 5489).  Result: 486 spheres on every platform.
#pragma once

#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "Sphere.h"
#include "../venusaur_b200.h"

namespace venusaur {

class Scene {
public:
    Scene() {
        Rng rng;
        add(Sphere(vec3(0.0, -1000.0, 0.0), 1000.0f, Material(Material::Lambertian, vec3(0.5, 0.5, 0.5))));
        for (int a = -11; a < 11; a++) {
            for (int b = -11; b < 11; b++) {
                const float choose_mat = rng.next();
                const float jx = rng.next();
                const float jz = rng.next();
                const vec3 center(a + 0.9 * jx, 0.2, b + 0.9 * jz);
                if (length(center - vec3(4, 0.2, 0)) > 0.9) {
                    if (choose_mat < 0.8) {
                        const vec3 c1 = rng.next3();
                        const vec3 c2 = rng.next3();
                        add(Sphere(center, 0.2f, Material(Material::Lambertian, c1 * c2)));
                    } else if (choose_mat < 0.95) {
                        const vec3 albedo = rng.next3(0.5f, 1.0f);
                        const float fuzz = rng.next(0.0f, 0.5f);
                        add(Sphere(center, 0.2f, Material(Material::Metal, albedo, fuzz)));
                    } else {
                        add(Sphere(center, 0.2f, Material(Material::Dielectric, vec3(0.0f), 0.0f, 1.5f)));
                    }
                }
            }
        }
        add(Sphere(vec3(0, 1, 0), 1.0f, Material(Material::Dielectric, vec3(0.0f), 0.0f, 1.5f)));
        add(Sphere(vec3(-4, 1, 0), 1.0f, Material(Material::Lambertian, vec3(0.4, 0.2, 0.1))));
        add(Sphere(vec3(4, 1, 0), 1.0f, Material(Material::Metal, vec3(0.7, 0.6, 0.5), 0.0f)));
    }

    // An empty scene to be filled with add() -- not in the reference, used by the synthetic configs.
    struct Empty {};
    explicit Scene(Empty) {}

    void add(const Sphere& s) {
        m_spheres.push_back(s);
        m_aabbs.push_back(s.GetAABB());
        m_indices.push_back(static_cast<uint32_t>(m_indices.size()));
    }

    // Flattens to the C ABI's sphere records (what Renderer::CreateSBT packs per sphere, Renderer.h:478-503).
    std::vector<vn_sphere> Flatten() const {
        std::vector<vn_sphere> out(m_spheres.size());
        for (size_t i = 0; i < m_spheres.size(); ++i) {
            const Sphere& s = m_spheres[i];
            const Material& m = s.GetMaterial();
            vn_sphere& o = out[i];
            o.cx = s.GetCenter().x; o.cy = s.GetCenter().y; o.cz = s.GetCenter().z; o.r = s.GetRadius();
            o.ax = m.GetAlbedo().x; o.ay = m.GetAlbedo().y; o.az = m.GetAlbedo().z;
            o.type = static_cast<uint32_t>(m.GetType());
            o.fuzz_or_ir = m.GetType() == Material::Dielectric ? m.GetIR() : m.GetFuzz();
        }
        return out;
    }

    std::vector<Sphere> m_spheres;
    std::vector<Aabb> m_aabbs;
    std::vector<uint32_t> m_indices;

private:
    struct Rng {
        std::mt19937 gen;
        float next() {
            float r = static_cast<float>(gen()) * 2.3283064365386963e-10f;   // 2^-32 (exact scaling)
            return r >= 1.0f ? std::nextafter(1.0f, 0.0f) : r;
        }
        float next(float lo, float hi) { return lo + (hi - lo) * next(); }
        vec3 next3() { vec3 v; v.x = next(); v.y = next(); v.z = next(); return v; }
        vec3 next3(float lo, float hi) { vec3 v; v.x = next(lo, hi); v.y = next(lo, hi); v.z = next(lo, hi); return v; }
    };
};

}  // namespace venusaur
