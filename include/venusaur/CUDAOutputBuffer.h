// CUDAOutputBuffer.h -- pixel buffer handed to Renderer::Draw; same public interface as the reference's
// Core/CUDAOutputBuffer.h:65-101 (map/unmap/width/height/resize/setStream/setDevice/getHostPointer).
// Backing modes: CUDA_DEVICE (what Core.cpp:252 always uses), CUDA_P2P (device memory; peers map it through
// vn_ipc_*), ZERO_COPY (mapped pinned host memory).  GL_INTEROP / getPBO need an OpenGL context, which a headless
// B200 node does not have: they throw.  Allocation goes through the C ABI so this header needs no CUDA toolkit.
#pragma once

#include <cstdint>
#include <vector>

#include "Exception.h"

#if !defined(__VECTOR_TYPES_H__) && !defined(VENUSAUR_HAVE_UCHAR4)
#define VENUSAUR_HAVE_UCHAR4
struct uchar4 { unsigned char x, y, z, w; };
#endif
#if !defined(__cuda_cuda_h__) && !defined(__DRIVER_TYPES_H__)
typedef struct CUstream_st* CUstream;
#endif

namespace venusaur {

enum class CUDAOutputBufferType { CUDA_DEVICE = 0, GL_INTEROP = 1, ZERO_COPY = 2, CUDA_P2P = 3 };

template <typename PIXEL_FORMAT>
class CUDAOutputBuffer {
public:
    // (type, width, height) as in the reference (CUDAOutputBuffer.h:68); device_idx is an extension -- the reference allocates on
    // device 0 in its constructor and setDevice() afterwards only affects later calls (CUDAOutputBuffer.h:90,104-110).  Here
    // setDevice() on a buffer that already holds pixels moves the allocation, so a buffer can never be mapped on one device and
    // freed or copied under another.
    CUDAOutputBuffer(CUDAOutputBufferType type, int32_t width, int32_t height, int32_t device_idx = 0) : m_type(type), m_device_idx(device_idx) {
        if (type == CUDAOutputBufferType::GL_INTEROP)
            throw Exception("CUDAOutputBuffer: GL_INTEROP needs an OpenGL context; use CUDA_DEVICE or ZERO_COPY");
        resize(width, height);
    }
    ~CUDAOutputBuffer() { release(); }
    CUDAOutputBuffer(const CUDAOutputBuffer&) = delete;
    CUDAOutputBuffer& operator=(const CUDAOutputBuffer&) = delete;

    void setDevice(int32_t device_idx) {
        if (device_idx == m_device_idx) return;
        const int32_t w = m_width, h = m_height;
        release();                       // frees under the device that owns the pixels
        m_device_idx = device_idx;
        m_width = m_height = 0;
        if (w > 0 && h > 0) resize(w, h);
    }
    void setStream(CUstream stream) { m_stream = stream; }

    void resize(int32_t width, int32_t height) {
        if (width < 1) width = 1;       // CUDAOutputBuffer.h:40-54 ensureMinimumSize
        if (height < 1) height = 1;
        if (m_width == width && m_height == height) return;
        release();
        m_width = width;
        m_height = height;
        const uint64_t bytes = static_cast<uint64_t>(width) * height * sizeof(PIXEL_FORMAT);
        void *dev = nullptr, *host = nullptr;
        check(vn_buffer_alloc(m_device_idx, bytes, m_type == CUDAOutputBufferType::ZERO_COPY, &dev, &host), "vn_buffer_alloc");
        m_device_pixels = static_cast<PIXEL_FORMAT*>(dev);
        m_host_zcopy_pixels = static_cast<PIXEL_FORMAT*>(host);
        if (!m_host_pixels.empty()) m_host_pixels.resize(static_cast<size_t>(width) * height);
    }

    PIXEL_FORMAT* map() { return m_device_pixels; }
    void unmap() { check(vn_stream_synchronize(m_device_idx, m_stream), "vn_stream_synchronize"); }   // CUDAOutputBuffer.h:264-281

    int32_t width() const { return m_width; }
    int32_t height() const { return m_height; }

    unsigned getPBO() { throw Exception("CUDAOutputBuffer::getPBO: no OpenGL in the B200 path; use getHostPointer()"); }
    void deletePBO() {}

    PIXEL_FORMAT* getHostPointer() {                                     // CUDAOutputBuffer.h:348-372
        if (m_type == CUDAOutputBufferType::ZERO_COPY) return m_host_zcopy_pixels;
        m_host_pixels.resize(static_cast<size_t>(m_width) * m_height);
        check(vn_buffer_copy_to_host(m_device_idx, m_host_pixels.data(), m_device_pixels,
                                     static_cast<uint64_t>(m_width) * m_height * sizeof(PIXEL_FORMAT)), "vn_buffer_copy_to_host");
        return m_host_pixels.data();
    }

private:
    static void check(int status, const char* what) {
        if (status != VN_OK) throw Exception(std::string(what) + " failed: " + vn_last_error(nullptr));
    }
    void release() {
        if (m_device_pixels || m_host_zcopy_pixels)
            vn_buffer_free(m_device_idx, m_device_pixels, m_host_zcopy_pixels, m_type == CUDAOutputBufferType::ZERO_COPY);
        m_device_pixels = nullptr;
        m_host_zcopy_pixels = nullptr;
    }

    CUDAOutputBufferType m_type;
    int32_t m_width = 0, m_height = 0;
    PIXEL_FORMAT* m_device_pixels = nullptr;
    PIXEL_FORMAT* m_host_zcopy_pixels = nullptr;
    std::vector<PIXEL_FORMAT> m_host_pixels;
    CUstream m_stream = nullptr;
    int32_t m_device_idx = 0;            // the device that owns m_device_pixels
};

}  // namespace venusaur
