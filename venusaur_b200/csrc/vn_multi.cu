// vn_multi.cu -- vn_multi_*: one host thread driving 1-8 devices behind the C ABI (include/venusaur_b200.h), built on the per-device
// handles of vn_api.cu.  Replaces nothing in the reference (it is single-GPU: cudaSetDevice(0), CUDAOutputBuffer.h:90,104); it is what
// lets the drop-in Renderer::Draw (Renderer.h:35-78) use a whole 8 x B200 box.
//
// Frame = n subframes of one view.  Device g renders subframes base + g, base + g + N, ... into its own partial-sum buffer
// (VN_ACCUM_SUM); an event marks the end of its last launch.  Every device then waits -- on its stream, not on the host -- for the
// events of all its peers and runs the fused reduce + tonemap kernel on its row slice: N peer loads per pixel over NVLink, one uchar4
// store into the image on devices[0], one float4 store of the sum into the result buffer on devices[0].  devices[0] finally waits for
// the N reduce events (and copies the image to the host if asked to).  The next frame's launches wait for the reduce events too, so
// nobody adds into a buffer a peer is still reading.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/venusaur_b200.h"

namespace {
constexpr int kMaxDevices = 8;
}

struct vn_multi_context {
    int n = 0;
    int device[kMaxDevices] = {};
    vn_handle h[kMaxDevices] = {};
    cudaStream_t stream[kMaxDevices] = {};
    cudaEvent_t rendered[kMaxDevices] = {};      // end of the device's last render launch of the frame
    cudaEvent_t reduced[kMaxDevices] = {};       // end of the device's reduce + tonemap kernel
    cudaEvent_t t0 = nullptr, t1 = nullptr;      // devices[0]: around the reduce phase (timing)
    bool reduce_outstanding = false;
    uint32_t width = 0, height = 0;
    uint32_t accumulated = 0;                    // subframes in the partial sums
    float4* result = nullptr;                    // devices[0]: sum over the devices (float4 per pixel)
    uint32_t* image_tmp = nullptr;               // devices[0]: staging for VN_IMAGE_HOST
    uint64_t result_pixels = 0;
    float ms_reduce = 0.0f;
    bool timing_valid = false;
    std::string last_error;
};

namespace {

std::string g_multi_error;

int mfail(vn_multi_context* m, int status, const std::string& msg) {
    if (m) m->last_error = msg;
    g_multi_error = msg;
    return status;
}

#define VNM_CUDA(m, call)                                                                                                   \
    do {                                                                                                                    \
        cudaError_t e_ = (call);                                                                                            \
        if (e_ != cudaSuccess)                                                                                              \
            return mfail((m), e_ == cudaErrorMemoryAllocation ? VN_ERR_OOM : VN_ERR_CUDA,                                   \
                         std::string("CUDA call (") + #call + ") failed with error: '" + cudaGetErrorString(e_) + "' (" +   \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                                              \
    } while (0)

#define VNM_DEV(m, i, call)                                                                                                 \
    do {                                                                                                                    \
        const int rc_ = (call);                                                                                             \
        if (rc_ != VN_OK) return mfail((m), rc_, std::string("device ") + std::to_string((m)->device[i]) + ": " + vn_last_error((m)->h[i])); \
    } while (0)

int ensure_frame_buffers(vn_multi_context* m, uint32_t width, uint32_t height, bool host_image) {
    const uint64_t pixels = (uint64_t)width * height;
    VNM_CUDA(m, cudaSetDevice(m->device[0]));
    if (pixels > m->result_pixels) {
        cudaFree(m->result); cudaFree(m->image_tmp);
        m->result = nullptr; m->image_tmp = nullptr; m->result_pixels = 0;
        VNM_CUDA(m, cudaMalloc(&m->result, pixels * sizeof(float4)));
        m->result_pixels = pixels;
    }
    if (host_image && !m->image_tmp) VNM_CUDA(m, cudaMalloc(&m->image_tmp, m->result_pixels * 4));
    return VN_OK;
}

}  // namespace

extern "C" {

const char* vn_multi_last_error(vn_multi_handle m) { return m ? m->last_error.c_str() : g_multi_error.c_str(); }

int vn_multi_create(const int* devices, int n, vn_multi_handle* out) {
    if (!out) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_create: out is NULL");
    *out = nullptr;
    if (!devices || n < 1 || n > kMaxDevices) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_create: 1..8 devices");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_create: the devices must be distinct");
    vn_multi_context* m = new vn_multi_context();
    m->n = n;
    auto bail = [&](int rc, const std::string& msg) { const std::string keep = msg; vn_multi_destroy(m); return mfail(nullptr, rc, keep); };
    for (int i = 0; i < n; i++) {
        m->device[i] = devices[i];
        const int rc = vn_create(devices[i], &m->h[i]);
        if (rc != VN_OK) return bail(rc, std::string("vn_multi_create: ") + vn_last_error(nullptr));
        m->stream[i] = static_cast<cudaStream_t>(vn_stream(m->h[i]));
    }
    for (int i = 0; i < n; i++) {
        if (cudaSetDevice(devices[i]) != cudaSuccess) return bail(VN_ERR_CUDA, "vn_multi_create: cudaSetDevice failed");
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (!can) return bail(VN_ERR_INVALID, "vn_multi_create: device " + std::to_string(devices[i]) + " cannot access device " + std::to_string(devices[j]) +
                                                      " as a peer (the fused reduce reads the peers' buffers over NVLink)");
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(VN_ERR_CUDA, std::string("vn_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        if (cudaEventCreateWithFlags(&m->rendered[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&m->reduced[i], cudaEventDisableTiming) != cudaSuccess)
            return bail(VN_ERR_CUDA, "vn_multi_create: cudaEventCreate failed");
    }
    cudaSetDevice(devices[0]);
    if (cudaEventCreate(&m->t0) != cudaSuccess || cudaEventCreate(&m->t1) != cudaSuccess) return bail(VN_ERR_CUDA, "vn_multi_create: cudaEventCreate failed");
    *out = m;
    return VN_OK;
}

void vn_multi_destroy(vn_multi_handle m) {
    if (!m) return;
    for (int i = 0; i < m->n; i++) {
        if (!m->h[i]) continue;
        cudaSetDevice(m->device[i]);
        if (m->stream[i]) cudaStreamSynchronize(m->stream[i]);
    }
    if (m->n > 0) {
        cudaSetDevice(m->device[0]);
        cudaFree(m->result); cudaFree(m->image_tmp);
        if (m->t0) cudaEventDestroy(m->t0);
        if (m->t1) cudaEventDestroy(m->t1);
    }
    for (int i = 0; i < m->n; i++) {
        if (m->h[i]) cudaSetDevice(m->device[i]);
        if (m->rendered[i]) cudaEventDestroy(m->rendered[i]);
        if (m->reduced[i]) cudaEventDestroy(m->reduced[i]);
        if (m->h[i]) vn_destroy(m->h[i]);
    }
    delete m;
}

int vn_multi_device_count(vn_multi_handle m) { return m ? m->n : 0; }
vn_handle vn_multi_device(vn_multi_handle m, int i) { return (m && i >= 0 && i < m->n) ? m->h[i] : nullptr; }
uint32_t vn_multi_subframes_accumulated(vn_multi_handle m) { return m ? m->accumulated : 0u; }

int vn_multi_set_option(vn_multi_handle m, const char* name, double value) {
    if (!m) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_set_option: NULL handle");
    for (int i = 0; i < m->n; i++) VNM_DEV(m, i, vn_set_option(m->h[i], name, value));
    return VN_OK;
}

int vn_multi_set_spheres(vn_multi_handle m, const vn_sphere* host_spheres, uint64_t n) {
    if (!m) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_set_spheres: NULL handle");
    for (int i = 0; i < m->n; i++) VNM_DEV(m, i, vn_set_spheres(m->h[i], host_spheres, n));
    return VN_OK;
}

int vn_multi_build_bvh(vn_multi_handle m) {
    if (!m) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_build_bvh: NULL handle");
    for (int i = 0; i < m->n; i++) VNM_DEV(m, i, vn_build_bvh(m->h[i]));      // replicated: the builder is deterministic
    return VN_OK;
}

int vn_multi_render(vn_multi_handle m, const vn_params* p, uint32_t n_subframes) {
    if (!m || !p) return mfail(m, VN_ERR_INVALID, "vn_multi_render: NULL argument");
    if (n_subframes < 1) return mfail(m, VN_ERR_INVALID, "vn_multi_render: n_subframes must be >= 1");
    if (p->row_begin != 0 || p->row_end != 0) return mfail(m, VN_ERR_INVALID, "vn_multi_render: row ranges are per-device (use the device handles)");
    if (p->accum_count != 0 && p->accum_count != m->accumulated)
        return mfail(m, VN_ERR_INVALID, "vn_multi_render: accum_count must be 0 (new accumulation) or the number of subframes accumulated so far (" + std::to_string(m->accumulated) + ")");
    const int N = m->n;
    const bool want_image = p->image && !(p->flags & VN_NO_TONEMAP);
    const bool host_image = want_image && (p->flags & VN_IMAGE_HOST);
    const bool resized = p->width != m->width || p->height != m->height;
    { const int rc = ensure_frame_buffers(m, p->width, p->height, host_image); if (rc != VN_OK) return rc; }
    // nobody adds into (or clears) a partial sum that a peer's reduce kernel of the previous frame may still be reading
    if (m->reduce_outstanding) {
        for (int g = 0; g < N; g++) {
            VNM_CUDA(m, cudaSetDevice(m->device[g]));
            for (int r = 0; r < N; r++) if (r != g) VNM_CUDA(m, cudaStreamWaitEvent(m->stream[g], m->reduced[r], 0));
        }
        m->reduce_outstanding = false;
    }
    if (resized || p->accum_count == 0) {
        for (int g = 0; g < N; g++) VNM_DEV(m, g, resized ? vn_resize(m->h[g], p->width, p->height) : vn_reset_accum(m->h[g]));
        m->width = p->width; m->height = p->height;
        m->accumulated = 0;
    }
    // sample-range sharding: device g renders subframes first + g, first + g + N, ...; launches go out round by round so that all
    // devices start together
    vn_params q = *p;
    q.image = nullptr;
    q.accum_count = 0;
    q.flags = (p->flags & (VN_FAST | VN_COUNTERS | VN_WAVEFRONT)) | VN_ACCUM_SUM | VN_NO_TONEMAP | VN_ASYNC;
    // (first one subframe per device -- a new view's first launch runs alone and the host waits for it before the next, which must not hold
    // up the other devices' first launches -- then each device's remaining subframes in one call: one launch where the scene allows)
    for (uint32_t k = 0; k < n_subframes && k < (uint32_t)N; k++) {
        q.subframe_index = p->subframe_index + k;
        VNM_DEV(m, (int)k, vn_render(m->h[k], &q));
    }
    for (uint32_t g = 0; g < (uint32_t)N && g + (uint32_t)N < n_subframes; g++) {
        const uint32_t rest = (n_subframes - g + (uint32_t)N - 1u) / (uint32_t)N - 1u;
        q.subframe_index = p->subframe_index + g + (uint32_t)N;
        VNM_DEV(m, (int)g, vn_render_subframes_strided(m->h[g], &q, rest, (uint32_t)N));
    }
    m->accumulated += n_subframes;
    void* peers[kMaxDevices];
    for (int g = 0; g < N; g++) {
        VNM_DEV(m, g, vn_accum_device_ptr(m->h[g], &peers[g]));
        VNM_CUDA(m, cudaSetDevice(m->device[g]));
        VNM_CUDA(m, cudaEventRecord(m->rendered[g], m->stream[g]));
    }
    // fused reduce + tonemap: device g owns rows [H g / N, H (g + 1) / N)
    const float scale = 1.0f / (float)m->accumulated;
    uint32_t* image = want_image ? (host_image ? m->image_tmp : static_cast<uint32_t*>(p->image)) : nullptr;
    VNM_CUDA(m, cudaSetDevice(m->device[0]));
    VNM_CUDA(m, cudaStreamWaitEvent(m->stream[0], m->rendered[0], 0));
    VNM_CUDA(m, cudaEventRecord(m->t0, m->stream[0]));
    for (int g = 0; g < N; g++) {
        VNM_CUDA(m, cudaSetDevice(m->device[g]));
        for (int r = 0; r < N; r++) if (r != g) VNM_CUDA(m, cudaStreamWaitEvent(m->stream[g], m->rendered[r], 0));
        const uint32_t r0 = (uint32_t)(((uint64_t)p->height * g) / N), r1 = (uint32_t)(((uint64_t)p->height * (g + 1)) / N);
        VNM_DEV(m, g, vn_reduce_tonemap_peers_to(m->h[g], peers, (uint32_t)N, scale, r0, r1, m->result, image, (p->flags & VN_FAST) | VN_ASYNC));
        VNM_CUDA(m, cudaEventRecord(m->reduced[g], m->stream[g]));
    }
    m->reduce_outstanding = true;
    VNM_CUDA(m, cudaSetDevice(m->device[0]));
    for (int r = 1; r < N; r++) VNM_CUDA(m, cudaStreamWaitEvent(m->stream[0], m->reduced[r], 0));
    VNM_CUDA(m, cudaEventRecord(m->t1, m->stream[0]));
    m->timing_valid = true;
    if (host_image) VNM_CUDA(m, cudaMemcpyAsync(p->image, m->image_tmp, (uint64_t)p->width * p->height * 4, cudaMemcpyDeviceToHost, m->stream[0]));
    if (!(p->flags & VN_ASYNC)) return vn_multi_synchronize(m);
    return VN_OK;
}

int vn_multi_synchronize(vn_multi_handle m) {
    if (!m) return mfail(nullptr, VN_ERR_INVALID, "vn_multi_synchronize: NULL handle");
    for (int g = m->n - 1; g >= 0; g--) VNM_DEV(m, g, vn_synchronize(m->h[g]));      // devices[0] last: its stream waits for everybody's reduce
    if (m->timing_valid) {
        VNM_CUDA(m, cudaSetDevice(m->device[0]));
        VNM_CUDA(m, cudaEventElapsedTime(&m->ms_reduce, m->t0, m->t1));
    }
    return VN_OK;
}

int vn_multi_read_accum(vn_multi_handle m, float* host_rgba) {
    if (!m || !host_rgba) return mfail(m, VN_ERR_INVALID, "vn_multi_read_accum: NULL argument");
    if (!m->result || !m->accumulated) return mfail(m, VN_ERR_INVALID, "vn_multi_read_accum: nothing rendered yet");
    { const int rc = vn_multi_synchronize(m); if (rc != VN_OK) return rc; }
    const uint64_t pixels = (uint64_t)m->width * m->height;
    VNM_CUDA(m, cudaSetDevice(m->device[0]));
    VNM_CUDA(m, cudaMemcpy(host_rgba, m->result, pixels * sizeof(float4), cudaMemcpyDeviceToHost));
    const float scale = 1.0f / (float)m->accumulated;                           // the factor the tonemap applied
    for (uint64_t i = 0; i < pixels; i++) { host_rgba[4 * i] *= scale; host_rgba[4 * i + 1] *= scale; host_rgba[4 * i + 2] *= scale; }
    return VN_OK;
}

int vn_multi_get_stats(vn_multi_handle m, vn_stats* out) {
    if (!m || !out) return mfail(m, VN_ERR_INVALID, "vn_multi_get_stats: NULL argument");
    memset(out, 0, sizeof(*out));
    for (int g = 0; g < m->n; g++) {
        vn_stats s;
        VNM_DEV(m, g, vn_get_stats(m->h[g], &s));
        out->segments += s.segments; out->paths += s.paths; out->node_visits += s.node_visits; out->sphere_tests += s.sphere_tests;
        out->segments_total += s.segments_total;
        out->kernel_launches += s.kernel_launches; out->kernel_launches_total += s.kernel_launches_total;
        if (s.ms_render > out->ms_render) out->ms_render = s.ms_render;
        if (s.ms_build > out->ms_build) out->ms_build = s.ms_build;
        if (s.ms_upload > out->ms_upload) out->ms_upload = s.ms_upload;
    }
    out->ms_trace = m->ms_reduce;
    return VN_OK;
}

}  // extern "C"
