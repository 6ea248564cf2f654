// Renderer.h -- drop-in for the reference's Core/Renderer.h: the same three public methods
//   void Init(const Scene&, const std::string ptxSource)      Renderer.h:25
//   void Draw(Camera&, CUDAOutputBuffer<uchar4>&)              Renderer.h:35
//   void Cleanup()                                             Renderer.h:80
// implemented over the C ABI of libvenusaur_b200.so instead of OptiX.  Header-only, C++17, no CUDA toolkit needed.
//
// Deliberate differences from the reference:
//   Q1  the accumulation buffer is zero-initialised and blended as a running mean over the frames since the last
//       camera change; the reference blends the first frame with uninitialised memory at weight 1/2.  The RNG
//       stream ids are unchanged (subframe_index = 1, 2, 3, ...).  -DVENUSAUR_STRICT_ACCUM restores the literal
//       weights 1/(subframe_index+1).
//   Q3  no per-frame leaks (stream, params buffer, accum);   Q4  accum follows the output buffer's size.
//   max_depth is a run-time setting (default 4 = the constant at RayTracer.cu:172).
#pragma once

#include <cstdint>
#include <string>

#include "CUDAOutputBuffer.h"
#include <utility>
#include <vector>

#include "Camera.h"
#include "Exception.h"
#include "Scene.h"

namespace venusaur {

class Renderer {
public:
    Renderer() = default;
    ~Renderer() { if (m_handle) vn_destroy(m_handle); if (m_multi) vn_multi_destroy(m_multi); }
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    // ptxSource is accepted for source compatibility and ignored: the kernels are precompiled sm_100a SASS.
    void Init(const Scene& scene, const std::string /*ptxSource*/ = std::string()) {
        if (!m_devices.empty()) {                                          // SetDevices: one vn_multi handle drives them all
            if (!m_multi && vn_multi_create(m_devices.data(), static_cast<int>(m_devices.size()), &m_multi) != VN_OK)
                throw Exception(std::string("vn_multi_create failed: ") + vn_multi_last_error(nullptr));
            for (const auto& kv : m_options) VN_MULTI_CHECK(m_multi, vn_multi_set_option(m_multi, kv.first.c_str(), kv.second));
            const std::vector<vn_sphere> flat = scene.Flatten();
            VN_MULTI_CHECK(m_multi, vn_multi_set_spheres(m_multi, flat.data(), flat.size()));
            VN_MULTI_CHECK(m_multi, vn_multi_build_bvh(m_multi));
            m_subframe_index = 0;
            m_accumulated = 0;
            return;
        }
        if (!m_handle) {
            if (vn_create(m_device, &m_handle) != VN_OK) throw Exception(std::string("vn_create failed: ") + vn_last_error(nullptr));
        }
        for (const auto& kv : m_options) VN_CHECK(m_handle, vn_set_option(m_handle, kv.first.c_str(), kv.second));
        const std::vector<vn_sphere> flat = scene.Flatten();            // CreateSBT, Renderer.h:452-520
        VN_CHECK(m_handle, vn_set_spheres(m_handle, flat.data(), flat.size()));
        VN_CHECK(m_handle, vn_build_bvh(m_handle));                     // BuildAccelerationStructures, Renderer.h:160-255
        m_subframe_index = 0;
        m_accumulated = 0;
    }

    void Draw(Camera& camera, CUDAOutputBuffer<uchar4>& outputBuffer) {
        if (!m_handle && !m_multi) throw Exception("Renderer::Draw called before Init");
        const uint32_t w = static_cast<uint32_t>(outputBuffer.width()), h = static_cast<uint32_t>(outputBuffer.height());
        const bool resized = (w != m_width || h != m_height);
        if (m_multi) {
            // one Draw = m_subframesPerDraw subframes dealt to the devices + one fused peer reduce + tonemap into the buffer, which must
            // live on the first device of SetDevices (CUDAOutputBuffer(type, w, h, device))
            if (camera.Changed() || resized) { m_width = w; m_height = h; m_subframe_index = 0u; m_accumulated = 0u; }
            vn_params p{};
            FillParams(p, camera, outputBuffer, w, h);
            p.subframe_index = m_subframe_index + 1u;
            p.accum_count = m_accumulated;
            p.flags = m_flags;
            VN_MULTI_CHECK(m_multi, vn_multi_render(m_multi, &p, m_subframesPerDraw));
            m_subframe_index += m_subframesPerDraw;
            m_accumulated += m_subframesPerDraw;
            return;
        }
        if (camera.Changed() || resized) {                              // Renderer.h:37-45
            if (resized) { VN_CHECK(m_handle, vn_resize(m_handle, w, h)); m_width = w; m_height = h; }
            else VN_CHECK(m_handle, vn_reset_accum(m_handle));
            m_subframe_index = 0u;
            m_accumulated = 0u;
        }
        outputBuffer.setStream(static_cast<CUstream>(vn_stream(m_handle)));

        vn_params p{};
        FillParams(p, camera, outputBuffer, w, h);
        p.subframe_index = ++m_subframe_index;                          // Renderer.h:54: incremented before the launch
#ifdef VENUSAUR_STRICT_ACCUM
        p.accum_count = m_subframe_index;                               // RayTracer.cu:208-213, literally
#else
        p.accum_count = m_accumulated;
#endif
        p.flags = m_flags | VN_ASYNC;
        // (SetSubframesPerDraw(n): what n Draw calls leave in the buffers, without the frames in between -- one launch for all of them
        // where the scene allows, vn_render_subframes)
        const uint32_t n = m_devices.empty() && m_subframesPerDraw > 1u ? m_subframesPerDraw : 1u;
        if (n > 1u) VN_CHECK(m_handle, vn_render_subframes(m_handle, &p, n));
        else VN_CHECK(m_handle, vn_render(m_handle, &p));               // optixLaunch, Renderer.h:75
        outputBuffer.unmap();                                           // Renderer.h:76
        VN_CHECK(m_handle, vn_synchronize(m_handle));                   // CUDA_SYNC_CHECK, Renderer.h:77
        m_subframe_index += n - 1u;
        m_accumulated += n;
    }

    void Cleanup() {                                                    // Renderer.h:80-97
        if (m_handle) { vn_destroy(m_handle); m_handle = nullptr; }
        if (m_multi) { vn_multi_destroy(m_multi); m_multi = nullptr; }
        m_width = m_height = 0;
    }

    // ---- extensions (not in the reference)
    void SetDevice(int device) { m_device = device; }
    // Several devices behind the same Init / Draw / Cleanup (before Init).  One Draw then advances the progressive render by
    // `subframes_per_draw` subframes (default: one per device) -- the image a single device has after that many Draw calls, up to float
    // re-association (vn_multi_render, include/venusaur_b200.h).  The output buffer must live on devices[0].
    void SetDevices(const std::vector<int>& devices, uint32_t subframes_per_draw = 0) {
        if (m_handle || m_multi) throw Exception("Renderer::SetDevices must be called before Init");
        m_devices = devices;
        m_subframesPerDraw = subframes_per_draw ? subframes_per_draw : static_cast<uint32_t>(devices.size());
    }
    // One device: a Draw advances the progressive render by n subframes (n Draw calls of the reference, Renderer.h:35-78, without the
    // frames in between; the accumulation buffer and the final image are bit for bit the same).
    void SetSubframesPerDraw(uint32_t n) { if (m_devices.empty()) m_subframesPerDraw = n ? n : 1u; }
    void SetMaxDepth(uint32_t max_depth) { m_maxDepth = max_depth; }
    void SetSamplesPerPixel(uint32_t spp) { m_samplesPerPixel = spp; }
    void SetFlags(uint32_t flags) { m_flags = flags; }
    // a tuning knob of the library (vn_set_option, include/venusaur_b200.h): applied at once when the renderer is initialised, and
    // (again) by every Init; knobs that shape the BVH take effect with the next Init
    void SetOption(const std::string& name, double value) {
        bool known = false;
        for (auto& kv : m_options) if (kv.first == name) { kv.second = value; known = true; }
        if (!known) m_options.emplace_back(name, value);
        if (m_handle) VN_CHECK(m_handle, vn_set_option(m_handle, name.c_str(), value));
        if (m_multi) VN_MULTI_CHECK(m_multi, vn_multi_set_option(m_multi, name.c_str(), value));
    }
    uint32_t SubframeIndex() const { return m_subframe_index; }
    vn_handle Handle() const { return m_handle; }
    vn_stats Stats() const { vn_stats s{}; if (m_handle) vn_get_stats(m_handle, &s); else if (m_multi) vn_multi_get_stats(m_multi, &s); return s; }
    vn_multi_handle MultiHandle() const { return m_multi; }

private:
    void FillParams(vn_params& p, Camera& camera, CUDAOutputBuffer<uchar4>& outputBuffer, uint32_t w, uint32_t h) const {
        p.image = outputBuffer.map();
        p.width = w;
        p.height = h;
        p.samples_per_pixel = m_samplesPerPixel;                        // Renderer.h:53
        p.max_depth = m_maxDepth;
        const vec3 origin = camera.GetPosition();
        vec3 u, v, wv;
        camera.UVWFrame(u, v, wv);                                      // Renderer.h:55-61
        p.origin[0] = origin.x; p.origin[1] = origin.y; p.origin[2] = origin.z;
        p.u[0] = u.x; p.u[1] = u.y; p.u[2] = u.z;
        p.v[0] = v.x; p.v[1] = v.y; p.v[2] = v.z;
        p.w[0] = wv.x; p.w[1] = wv.y; p.w[2] = wv.z;
        p.lens_radius = camera.GetLensRadius();
    }

    vn_handle m_handle = nullptr;
    vn_multi_handle m_multi = nullptr;
    std::vector<int> m_devices;
    uint32_t m_subframesPerDraw = 1;
    std::vector<std::pair<std::string, double>> m_options;
    int m_device = 0;
    uint32_t m_width = 0, m_height = 0;
    uint32_t m_subframe_index = 0;
    uint32_t m_accumulated = 0;
    uint32_t m_samplesPerPixel = 16;     // Renderer.h:53,135
    uint32_t m_maxDepth = 4;             // RayTracer.cu:172
    uint32_t m_flags = 0;
};

}  // namespace venusaur

#ifndef VENUSAUR_NO_GLOBAL_NAMES
// The reference declares these classes in the global namespace (Core.cpp:21-31 uses them unqualified).
using venusaur::Camera;
using venusaur::CUDAOutputBuffer;
using venusaur::CUDAOutputBufferType;
using venusaur::Exception;
using venusaur::Material;
using venusaur::Renderer;
using venusaur::Scene;
using venusaur::Sphere;
#endif
