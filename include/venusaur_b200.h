/*
 * venusaur_b200.h -- C ABI of libvenusaur_b200.so, the B200-native replacement for Venusaur's one
 * data-parallel hot path (the per-pixel Monte-Carlo loop of Core/RayTracer.cu and the three Renderer
 * methods that feed it).
 *
 * The reference has no FFI layer: Core/Core.cpp calls Renderer::{Init,Draw,Cleanup} (Renderer.h:25,35,80)
 * and Renderer talks to OptiX.  This header is the boundary a maintainer binds instead of OptiX; the C++17
 * drop-in classes in include/venusaur/ (Renderer, Scene, Camera, CUDAOutputBuffer, Exception) are a thin
 * header-only shim over it.  Every entry point names the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; no function throws; each returns VN_OK (0) or a negative
 * vn_status, and vn_last_error() returns the message (the C++ shim turns it into `throw Exception(msg)`,
 * mirroring CUDA_CHECK / OPTIX_CHECK in Exception.h:16-102).  A handle is bound to one CUDA device and is
 * not thread-safe (the reference is single-threaded, Core.cpp:358-430).  There is no CPU fallback: without
 * a CUDA device vn_create fails.
 */
#ifndef VENUSAUR_B200_H
#define VENUSAUR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VN_API __declspec(dllexport)
#else
#define VN_API __attribute__((visibility("default")))
#endif

typedef struct vn_context* vn_handle;

typedef enum vn_status {
    VN_OK = 0,
    VN_ERR_INVALID = -1,     /* bad argument / call order */
    VN_ERR_CUDA = -2,        /* a CUDA call or kernel failed; message has the call and the CUDA error string */
    VN_ERR_NO_DEVICE = -3,   /* no usable CUDA device (there is no CPU path) */
    VN_ERR_OOM = -4
} vn_status;

/* Material::Type, material.h:9-14 */
enum { VN_LAMBERTIAN = 0, VN_METAL = 1, VN_DIELECTRIC = 2 };

/* One sphere = SphereHitGroupData (RayTracer.h:40-45; center, radius, MaterialData{albedo,fuzz}|{ir}) plus the
 * material type that the reference encodes in the SBT program header (Renderer.h:487-503).  36 bytes. */
typedef struct vn_sphere {
    float cx, cy, cz, r;
    float ax, ay, az;       /* albedo: Lambertian, metal */
    float fuzz_or_ir;       /* metal: fuzz; dielectric: index of refraction */
    uint32_t type;          /* VN_LAMBERTIAN / VN_METAL / VN_DIELECTRIC */
} vn_sphere;

/* vn_params.flags */
enum {
    VN_EXACT          = 1u << 0, /* the default build, named for explicitness: IEEE kernels (no FMA contraction, IEEE div/sqrt),
                                    bit-identical to the host oracle -- the build that meets BASELINE.json's image tolerance */
    VN_IMAGE_HOST     = 1u << 1, /* params.image is HOST memory: the uchar4 frame is copied D2H before returning
                                    (what CUDAOutputBuffer::getHostPointer does, CUDAOutputBuffer.h:348-372) */
    VN_ACCUM_SUM      = 1u << 2, /* accum += frame mean (multi-GPU partial sums) instead of the running mean */
    VN_NO_TONEMAP     = 1u << 3, /* do not write params.image */
    VN_WAVEFRONT      = 1u << 4, /* use the queue-based wavefront kernels instead of the persistent path kernel */
    VN_COUNTERS       = 1u << 5, /* instrumented launch: also count BVH node visits and sphere tests */
    VN_ASYNC          = 1u << 6, /* do not synchronise the stream before returning (stats are then stale).  With VN_IMAGE_HOST the
                                    frame's D2H copy is pipelined on a second stream under the next vn_render's kernel: the host buffer
                                    is complete after vn_synchronize (or two vn_render calls later); alternate between two host buffers */
    VN_GRID           = 1u << 11, /* vn_trace_rays only: query the uniform grid + oversize list instead of the BVH */
    VN_FAST           = 1u << 7  /* relaxed-numerics build (FMA contraction, approximate rcp/rsqrt/sqrt, FP32 for the FP64
                                    fragments): a few % faster, PSNR > 60 dB vs the oracle but NOT within the 1e-3 per-pixel
                                    tolerance at 1024 spp (individual paths diverge); never the default */
};

/* Launch parameters = Params (RayTracer.h:3-17) minus the OptiX handle, plus what the reference hard-codes:
 * max_depth (RayTracer.cu:172, constant 4 there) and the blend weight it derives from subframe_index
 * (RayTracer.cu:208-213). */
typedef struct vn_params {
    void* image;                  /* uchar4[width*height], device pointer (or host with VN_IMAGE_HOST), may be NULL */
    uint32_t width, height;
    uint32_t samples_per_pixel;   /* Renderer.h:53 uses 16 */
    uint32_t subframe_index;      /* RNG stream id: seed = tea<4>(pixel, subframe_index), RayTracer.cu:169 */
    uint32_t max_depth;           /* max ray segments per path; reference = 4 */
    uint32_t accum_count;         /* frames already in accum: new = prev + (mean-prev)/(accum_count+1); 0 overwrites.
                                     The reference passes subframe_index here (RayTracer.cu:210). */
    float origin[3], u[3], v[3], w[3];
    float lens_radius;
    uint32_t row_begin, row_end;  /* render rows [row_begin,row_end) only; 0,0 = whole frame (tile sharding) */
    uint32_t flags;
} vn_params;

typedef struct vn_stats {
    uint64_t segments;        /* ray segments (closest-hit queries) traced by the last vn_render */
    uint64_t paths;
    uint64_t node_visits;     /* VN_COUNTERS only */
    uint64_t sphere_tests;    /* VN_COUNTERS only */
    uint64_t segments_total;  /* since vn_create / vn_reset_stats */
    uint32_t kernel_launches; /* kernels launched by the last vn_render */
    uint32_t kernel_launches_total;
    float ms_render;          /* CUDA-event time of the last vn_render's kernels (0 with VN_ASYNC) */
    float ms_trace;           /* the dominant trace kernel(s) alone */
    float ms_build;           /* last vn_build_bvh */
    float ms_upload;          /* last vn_set_spheres H2D */
} vn_stats;

typedef struct vn_bvh_info {
    uint64_t num_spheres;
    uint64_t num_nodes;       /* packed 32-byte nodes */
    uint32_t max_leaf_size;
    uint32_t scene_in_smem;   /* 1 when nodes+spheres are staged in shared memory by the trace kernels */
    float bounds_lo[3], bounds_hi[3];
} vn_bvh_info;

/* 32-byte BVH node as laid out in HBM (read by the kernels with two 128-bit loads).  Children of an internal
 * node are adjacent: child pair at [link, link+1]. */
typedef struct vn_node32 {
    float lo[3];
    uint32_t link;            /* internal: index of the left child (right = link+1); leaf: 0x80000000 | first<<3 | (count-1) */
    float hi[3];
    uint32_t aux;             /* number of spheres under this node */
} vn_node32;

/* ---- lifetime: Renderer::CreateContext / Cleanup (Renderer.h:144-158, 80-97) ---- */
VN_API int vn_create(int device, vn_handle* out);
VN_API void vn_destroy(vn_handle h);
VN_API const char* vn_last_error(vn_handle h);            /* h may be NULL: error of the last failed vn_create */
VN_API int vn_device_count(void);
VN_API const char* vn_version(void);

/* ---- scene: Renderer::CreateSBT (Renderer.h:452-520) + BuildAccelerationStructures (Renderer.h:160-255) ---- */
VN_API int vn_set_spheres(vn_handle h, const vn_sphere* host_spheres, uint64_t n);
/* Tuning knobs (defaults in brackets).  BVH: "leaf_size" [0 = auto], "aabb_pad" [0.01], "sah_max_prims" [4096: SAH splits up to this size],
 * "wide_max_prims" [16384: 4-wide nodes up to this size], "accel" [1 = BVH; 0 = auto: the uniform grid when the scene suits it, 2 = grid], "huge_factor" [50], "grid_max_per_cell" [16].  Path kernel: "wide_nodes" [1], "octant_nodes" [1], "leaf_vote" [0 = while-while; n: lanes waiting
 * at a leaf that trigger the warp's leaf turn], "wide_threads" [1024], "async_done" [26: k_render_async, a traversal burst ends when that many lanes of the
 * warp hold a finished ray; 0 = k_render_persistent, every round waits for its slowest ray], "async_node" [0 = no votes inside a phase; n: node steps while n
 * lanes stand on nodes], "async_leaf" [8], "warp_tiles" [1: in the phase form a warp takes a whole 8x4-pixel tile with one ticket and hands the pixels to its own
 * lanes; 0 = every lane takes single pixels from the global ticket], "tile_order" [1: once a view has been
 * rendered once, its 8x4-pixel tiles are handed out most expensive first (ray segments per tile, counted by that first launch); 0 = row-major, 2 = by the most
 * expensive pixel, 3 = max(sum / 8, most expensive pixel), 4 = like 1 with the tiles in which no path hit anything last], "wide_global" [0], "threads", "blocks_per_sm", "smem_scene_limit".
 * "lean" [1: k_render_lean -- 16-bit links, newest stack entry in a register, per-warp statistics -- for shared-memory scenes, and its
 * asynchronous form over pair nodes for scenes traversed from L2 / HBM; 0 = k_render_async / k_render_persistent], "global_done" [16: the burst
 * threshold of the L2 / HBM form], "global_ctas" [5: its CTAs of 256 threads per SM, 4..6], "hit_gate" [1: scenes traversed from L2 / HBM only count a root whose hit point lies inside the sphere's slightly
 * grown box; 0 = never; 2 = the pair-node kernels also on small scenes], "wavefront_slots".  Options that change the BVH invalidate it (call vn_build_bvh again). */
VN_API int vn_set_option(vn_handle h, const char* name, double value);
VN_API int vn_build_bvh(vn_handle h);
/* The spheres moved or changed material (same count, same order as in vn_set_spheres): uploads the records and refits the existing BVH --
 * same hierarchy, new boxes, 4-5 launches instead of a build (moving spheres).  The closest hit stays exact for any
 * motion; traversal cost grows as spheres leave their old neighbourhoods: call vn_build_bvh now and then.  Falls back to a full build
 * when the hierarchy is not at hand (very large scenes' wide nodes, the uniform grid).  vn_stats.ms_build reports the refit. */
VN_API int vn_update_spheres(vn_handle h, const vn_sphere* host_spheres, uint64_t n);
VN_API int vn_get_bvh_info(vn_handle h, vn_bvh_info* out);
VN_API int vn_read_bvh(vn_handle h, vn_node32* host_nodes, uint64_t cap_nodes, uint32_t* host_prim_order, uint64_t cap_prims);
/* Second closest-hit structure for small scenes (<= 16384 spheres of similar size plus at most 8 oversize ones): a uniform grid over the
 * Morton-sorted spheres + a list of oversize spheres every ray tests first.  Returns 1 and fills header104 = { float lo[3], inv_cell[3],
 * cell[3], hi[3]; uint32 res[3], n_cells, n_refs, n_big, big[8] }, start[n_cells + 1], refs[n_refs] (sorted sphere indices); 0 when the
 * scene has no grid ("accel" = 1, or it does not suit the structure). */
VN_API int vn_read_grid(vn_handle h, void* header104, uint16_t* host_start, uint64_t cap_start, uint16_t* host_refs, uint64_t cap_refs);
/* Spheres left out of the wide nodes because nearly every ray enters their box (radius > 50 x median; RTIOW: the ground): the path kernel
 * tests them before the traversal.  Returns their number (<= 8) and their sorted indices. */
VN_API int vn_read_huge(vn_handle h, uint32_t* idx8);
VN_API int vn_last_accel(vn_handle h);   /* what the last vn_render traversed: 1 pair nodes, 2 wide nodes, 3 wide nodes from L2/HBM, 4 grid */
/* 4-wide nodes derived from the pairs for scenes traversed out of shared memory (<= "wide_max_prims" spheres): 128 B per
 * node = 4 x { {lo.xyz, link}, {hi.xyz, count} }; link = wide-node index, a leaf link as above, or 0xFFFFFFFF (empty slot).
 * num_nodes_out = 0 when the scene has none. */
VN_API int vn_read_wide_bvh(vn_handle h, float* host_nodes, uint64_t cap_nodes, uint32_t* num_nodes_out, uint32_t* levels_out);

/* ---- frame: Renderer::Draw (Renderer.h:35-78) ---- */
VN_API int vn_resize(vn_handle h, uint32_t width, uint32_t height);   /* (re)allocates + zeroes accum (fixes Q3/Q4) */
VN_API int vn_reset_accum(vn_handle h);                               /* camera.Changed() path, Renderer.h:37-45 */
VN_API int vn_render(vn_handle h, const vn_params* p);                /* optixLaunch, Renderer.h:75 */
/* n calls of Renderer::Draw without looking at the frames in between (the reference's frame loop, Core.cpp:358-395, with the display
 * taken out; Renderer.h:50-54,75): subframes p->subframe_index .. + n - 1 on top of p->accum_count accumulated ones, the accumulation buffer
 * bit for bit what n vn_render calls leave, the image (if any) made once at the end.  Scenes rendered from shared memory take all of them
 * in ONE launch of the path kernel ("multi_subframes", <= 64 per launch): the launch drains once instead of n times. */
VN_API int vn_render_subframes(vn_handle h, const vn_params* p, uint32_t n);
/* ... subframes p->subframe_index + k * stride, k < n: the share of one device when a frame's subframes are dealt round-robin to `stride`
 * devices (sample-range sharding; with VN_ACCUM_SUM the buffer holds the sum of the subframe means). */
VN_API int vn_render_subframes_strided(vn_handle h, const vn_params* p, uint32_t n, uint32_t stride);
VN_API int vn_tonemap(vn_handle h, float scale, void* image, uint32_t flags); /* image = make_color(accum*scale), RayTracer.cu:16-47,216 */
VN_API int vn_synchronize(vn_handle h);                               /* CUDA_SYNC_CHECK, Renderer.h:77 */
VN_API int vn_get_stats(vn_handle h, vn_stats* out);
VN_API int vn_reset_stats(vn_handle h);
/* Launch timeline of the last VN_COUNTERS launch of the asynchronous path kernels, in out14[0..2]: ~(earliest CTA start), ~(time at which the
 * first lane found the tile tickets exhausted), latest warp end -- %globaltimer nanoseconds (tools/tail_probe.py turns them into the drain). */
VN_API int vn_read_sched_counters(vn_handle h, uint64_t* out14);
/* Drain histogram of the last VN_COUNTERS launch of k_render_lean: out[b] = lanes that ran out of work in the b-th 8.192 us bin after their
 * CTA started, out[1024 + b] = the ray segments of the last pixels those lanes finished (filled by the cost-collecting launch of a view). */
VN_API int vn_read_timeline(vn_handle h, uint32_t* out2048);
/* The same plus out[2048 + b] = work tiles fetched in bin b and out[3072 + b] = ray segments shaded in bin b: the launch's throughput over time. */
VN_API int vn_read_timeline_ex(vn_handle h, uint32_t* out4096);

/* ---- accumulation buffer: Params::accum (RayTracer.h:6), float4 per pixel ---- */
VN_API int vn_read_accum(vn_handle h, float* host_rgba);              /* D2H, width*height*4 floats */
VN_API int vn_write_accum(vn_handle h, const float* host_rgba);       /* H2D (resume a progressive render) */
VN_API int vn_accum_device_ptr(vn_handle h, void** dev_ptr);          /* for the cross-GPU reduce (NCCL / peer kernel) */
VN_API int vn_set_accum_external(vn_handle h, void* dev_ptr);         /* render into caller-owned float4[width*height]; NULL restores */
/* fused reduce + tonemap over peer-mapped accumulation buffers: image rows [row_begin,row_end) =
 * make_color(scale * sum_i peers[i]) ; also writes the sum into this handle's accum. */
VN_API int vn_reduce_tonemap_peers(vn_handle h, const void* const* peer_accum, uint32_t n_peers, float scale,
                                   uint32_t row_begin, uint32_t row_end, void* image, uint32_t flags);

/* the same kernel with the float4 sum written to `sum_out` (device pointer, may be a peer's memory or NULL) instead of this handle's
 * accum: the partial sums stay intact, so a progressive render can go on accumulating into them (vn_multi_render) */
VN_API int vn_reduce_tonemap_peers_to(vn_handle h, const void* const* peer_accum, uint32_t n_peers, float scale,
                                      uint32_t row_begin, uint32_t row_end, void* sum_out, void* image, uint32_t flags);

/* Device-side ordering between GPUs owned by DIFFERENT processes (one process per GPU, the launch bench.py is given): every handle
 * owns 62 epoch flags in device memory (vn_sync_flags returns their address; export it with vn_ipc_export).  vn_signal(h, i, e) makes
 * flag i = e once everything queued on the handle's stream so far is complete and visible to peers; vn_wait_flags holds the stream
 * until all the given flags (peers' memory) have reached e; vn_reduce_tonemap_peers_wait is the fused reduce + tonemap that first waits
 * for one flag per peer (wait_value != 0).  Epochs only grow.  No host barrier is needed around the once-per-frame reduce.  A wait
 * gives up after 4 s; vn_check_flags reports that as an error. */
VN_API int vn_sync_flags(vn_handle h, void** dev_ptr);
VN_API int vn_signal(vn_handle h, uint32_t index, uint32_t value);
VN_API int vn_wait_flags(vn_handle h, const void* const* flags, uint32_t n, uint32_t value);
VN_API int vn_check_flags(vn_handle h);
VN_API int vn_reduce_tonemap_peers_wait(vn_handle h, const void* const* peer_accum, uint32_t n_peers, float scale, uint32_t row_begin,
                                        uint32_t row_end, void* sum_out, void* image, const void* const* peer_flags, uint32_t wait_value,
                                        uint32_t flags);

/* ---- one host thread, several devices (the handle drives 1-8 devices so that Renderer::Draw, Renderer.h:35-78, scales on
 * an 8 x B200 box without any help from the caller).  Scene and BVH are replicated (every device builds its own LBVH: deterministic);
 * vn_multi_render deals the subframes of one view round-robin to the devices (sample-range sharding: the seed is a pure function of
 * pixel and subframe, RayTracer.cu:169), each device adds its subframes' means into its own partial-sum buffer, and ONE fused
 * reduce + tonemap kernel per device then loads its row slice of every peer's buffer over NVLink (cudaDeviceEnablePeerAccess) and
 * writes the frame mean's pixels straight into the image on devices[0].  Ordering is by CUDA events between the devices' streams;
 * the host never waits inside a frame.  The image equals a single device's running mean over the same subframes up to float
 * re-association (sum-then-divide vs the sequential lerp of RayTracer.cu:208-213). */
typedef struct vn_multi_context* vn_multi_handle;
VN_API int vn_multi_create(const int* devices, int n_devices, vn_multi_handle* out);   /* 1..8 distinct devices that are peers of each other */
VN_API void vn_multi_destroy(vn_multi_handle m);
VN_API const char* vn_multi_last_error(vn_multi_handle m);
VN_API int vn_multi_device_count(vn_multi_handle m);
VN_API vn_handle vn_multi_device(vn_multi_handle m, int i);   /* the i-th device's own handle (BVH read-back, per-device statistics) */
VN_API int vn_multi_set_option(vn_multi_handle m, const char* name, double value);
VN_API int vn_multi_set_spheres(vn_multi_handle m, const vn_sphere* host_spheres, uint64_t n);
VN_API int vn_multi_build_bvh(vn_multi_handle m);
/* Renders subframes p->subframe_index, +1, ..., + n_subframes - 1 (p->samples_per_pixel each) and combines ALL subframes accumulated
 * since the last reset into p->image (uchar4: device memory of devices[0], host memory with VN_IMAGE_HOST, or NULL).
 * p->accum_count = 0 starts a new accumulation (the camera.Changed() path); otherwise it must be the number of subframes accumulated
 * so far.  Returns without waiting for the devices when VN_ASYNC is set. */
VN_API int vn_multi_render(vn_multi_handle m, const vn_params* p, uint32_t n_subframes);
VN_API int vn_multi_synchronize(vn_multi_handle m);
VN_API int vn_multi_read_accum(vn_multi_handle m, float* host_rgba);   /* the frame mean: (sum over devices) / subframes accumulated */
/* sums over the devices; ms_render = the slowest device's last launch, ms_trace = device time of the last reduce + tonemap on devices[0] */
VN_API int vn_multi_get_stats(vn_multi_handle m, vn_stats* out);
VN_API uint32_t vn_multi_subframes_accumulated(vn_multi_handle m);

/* ---- plain device/host buffer helpers, so that the header-only C++ shim (CUDAOutputBuffer.h:173-281,348-372) needs no
 * CUDA toolkit of its own ---- */
VN_API int vn_buffer_alloc(int device, uint64_t bytes, int zero_copy_host, void** dev_ptr, void** host_ptr);
VN_API int vn_buffer_free(int device, void* dev_ptr, void* host_ptr, int zero_copy_host);
VN_API int vn_buffer_copy_to_host(int device, void* host_dst, const void* dev_src, uint64_t bytes);
VN_API int vn_stream_synchronize(int device, void* cuda_stream);
VN_API void* vn_stream(vn_handle h);                                  /* the handle's cudaStream_t */
/* CUDA IPC, to map another process's accumulation buffer for vn_reduce_tonemap_peers (one process per GPU) */
VN_API int vn_ipc_export(vn_handle h, void* dev_ptr, unsigned char handle_out[64]);     /* dev_ptr must be the base of its allocation */
/* for a pointer INSIDE an allocation (e.g. a tensor of a caching allocator): the allocation's handle + the pointer's offset in it;
 * the peer adds the offset to what vn_ipc_open returns */
VN_API int vn_ipc_export_at(vn_handle h, void* dev_ptr, unsigned char handle_out[64], uint64_t* offset_out);
VN_API int vn_ipc_open(vn_handle h, const unsigned char handle_in[64], void** dev_ptr);
VN_API int vn_ipc_close(vn_handle h, void* dev_ptr);

/* ---- host-side scene/camera helpers (Scene.h:13-80, Camera.cpp:24-37); no GPU needed ---- */
VN_API uint32_t vn_scene_rtiow_final(vn_sphere* out, uint32_t cap);
VN_API void vn_scene_random(vn_sphere* out, uint64_t n, uint32_t seed, float S, uint32_t mix);
VN_API void vn_camera_frame(const float lookfrom[3], const float forward[3], float vfov_deg, float aspect, float aperture,
                            float focal_length, float origin[3], float u[3], float v[3], float w[3], float* lens_radius);

/* ---- unit-level entry points used by the parity tests (each runs a CUDA kernel) ---- */
/* random.cuh:31-67 on the device: out[i] = tea<4>(v0[i], v1[i]); then n_draws LCG draws from that state. */
VN_API int vn_test_rng(vn_handle h, const uint32_t* v0, const uint32_t* v1, uint64_t n, uint32_t n_draws,
                       uint32_t* seeds_out, uint32_t* lcg_out, float* rnd_out);
/* closest hit (tmin 1e-3, tmax 1e16) through the BVH for arbitrary host rays: t (or -1) and ORIGINAL sphere index */
VN_API int vn_trace_rays(vn_handle h, const float* origins, const float* dirs, uint64_t n, float* t_out,
                         int32_t* prim_out, uint32_t flags);
/* the hand-written onesweep radix sort: sorts host (key,value) pairs on the device, keys ascending, stable */
VN_API int vn_sort_pairs(vn_handle h, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t key_bits);
/* 30-bit Morton codes of the uploaded spheres' centroids (LBVH builder stage 2) */
VN_API int vn_morton_codes(vn_handle h, uint32_t* codes_out, uint64_t cap);
/* make_color on the device (RayTracer.cu:16-47) */
VN_API int vn_test_make_color(vn_handle h, const float* rgb, uint64_t n, uint8_t* rgba_out, uint32_t flags);
/* one scatter event on the device: RayTracer.cu:272-440.  in: per-event ray dir(3), hit normal(3), front(1), seed;
 * out: new dir(3), scattered flag, seed after. */
VN_API int vn_test_scatter(vn_handle h, uint32_t material_type, const float albedo_fuzz_ir[4], const float* dirs,
                           const float* normals, const uint8_t* front, const uint32_t* seeds, uint64_t n,
                           float* dirs_out, uint8_t* scattered_out, uint32_t* seeds_out, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* VENUSAUR_B200_H */
