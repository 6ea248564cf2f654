// Camera.h -- thin-lens camera; same public interface as the reference's Core/camera.h:9-100 + Core/Camera.cpp.
// glm is replaced by the float formulas glm uses (normalize = v * (1/sqrt(dot)), cross, radians, tan).
#pragma once

#include <cmath>

#include "Vec3.h"

namespace venusaur {

class Camera {
public:
    Camera() { UpdateUVW(); }

    Camera(const vec3& origin, float vfov, float aspect, float aperture, float focalLength)
        : m_position(origin), m_vfov(vfov), m_aspect(aspect), m_aperture(aperture), m_focalLength(focalLength) {
        UpdateUVW();
    }

    void UVWFrame(vec3& U, vec3& V, vec3& W) {
        UpdateUVW();
        U = m_u; V = m_v; W = m_w;
    }

    inline float GetLensRadius() const { return m_aperture * 0.5f; }          // camera.h:20
    inline const vec3& GetPosition() const { return m_position; }
    inline void SetForward(vec3 direction) { m_forward = normalize(direction); }   // camera.h:24 (does not mark dirty)

    inline float GetFocalLength() { return m_focalLength; }
    inline void SetFocalLength(float length) {
        if (m_focalLength != length) { m_focalLength = length; m_changed = true; }
    }
    inline void SetAspect(float aspect) { if (m_aspect != aspect) { m_aspect = aspect; m_changed = true; } }

    inline void MoveForward(float speed) { m_position += m_forward * speed; m_changed = true; }
    inline void MoveRight(float speed) { m_position += normalize(m_u) * speed; m_changed = true; }
    inline void MoveUp(float speed) { m_position += normalize(m_v) * speed; m_changed = true; }

    // camera.h:56-78: forward = normalize(forward * quat(angle, axis)); v * q rotates by the inverse of q.
    void Pitch(float speed) { Rotate(speed, m_u); }
    void Yaw(float speed) { Rotate(speed, m_v); }
    void Roll(float speed) { Rotate(speed, m_w); }

    // Returns the dirty flag and clears it (camera.h:80-85); Renderer::Draw restarts accumulation when it is set.
    bool Changed() {
        const bool changed = m_changed;
        m_changed = false;
        return changed;
    }

private:
    void Rotate(float degrees, const vec3& axis) {
        const float angle = -degrees * 0.01745329251994329576923690768489f;
        const vec3 k = normalize(axis);
        const float c = std::cos(angle), s = std::sin(angle);
        const vec3 f = m_forward;
        m_forward = normalize(f * c + cross(k, f) * s + k * (dot(k, f) * (1.0f - c)));
        m_changed = true;
    }

    // Camera.cpp:24-37
    void UpdateUVW() {
        m_w = m_forward * m_focalLength;
        m_u = normalize(cross(m_w, vec3(0.0, 1.0, 0.0)));
        m_v = normalize(cross(m_u, m_w));
        const float theta = m_vfov * 0.01745329251994329576923690768489f;   // glm::radians
        const float h = std::tan(theta * 0.5f);
        const float viewportHeight = 2.0f * h;
        const float viewportWidth = m_aspect * viewportHeight;
        m_u *= m_focalLength * viewportWidth;
        m_v *= m_focalLength * viewportHeight;
    }

    vec3 m_position = vec3(0.0f);
    vec3 m_forward = vec3(0.0, 0.0, -1.0);
    float m_vfov = 45.0f;
    float m_aspect = 1.6f;
    float m_aperture = 0.0f;
    float m_focalLength = 1.0f;
    vec3 m_u, m_v, m_w;
    bool m_changed = true;
};

}  // namespace venusaur
