"""Python mirror of the reference's host interface for the hot path, over the C ABI.

Same names and meaning as Core/Renderer.h (Init / Draw / Cleanup), Core/Scene.h (m_spheres), Core/camera.h
(UVWFrame, GetLensRadius, GetPosition, SetForward, Changed) and Core/CUDAOutputBuffer.h (map / unmap / width /
height / resize / getHostPointer).  `Context` is the thin, explicit layer (one method per vn_* call) the tests and
bench.py use.  Every compute call runs CUDA kernels; a failing call raises Exception with vn_last_error().
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import (VN_ACCUM_SUM, VN_ASYNC, VN_COUNTERS, VN_DIELECTRIC, VN_EXACT, VN_FAST, VN_IMAGE_HOST, VN_GRID, VN_LAMBERTIAN,  # noqa: F401
                   VN_METAL, VN_NO_TONEMAP, VN_WAVEFRONT, vn_bvh_info, vn_node32, vn_params, vn_sphere, vn_stats)

SPHERE_DTYPE = np.dtype([("cx", "f4"), ("cy", "f4"), ("cz", "f4"), ("r", "f4"), ("ax", "f4"), ("ay", "f4"),
                         ("az", "f4"), ("fuzz_or_ir", "f4"), ("type", "u4")])
NODE_DTYPE = np.dtype([("lo", "f4", 3), ("link", "u4"), ("hi", "f4", 3), ("aux", "u4")])
assert SPHERE_DTYPE.itemsize == C.sizeof(vn_sphere) == 36 and NODE_DTYPE.itemsize == C.sizeof(vn_node32) == 32


class Exception(RuntimeError):  # noqa: A001 - same name as the reference's class (Exception.h:136-154)
    pass


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- host-side scene / camera (no GPU needed)
def rtiow_final_scene() -> np.ndarray:
    """Scene() of Scene.h:13-80 as a structured array of 36-byte sphere records (486 spheres)."""
    lib = L.load()
    n = lib.vn_scene_rtiow_final(None, 0)
    out = np.zeros(n, SPHERE_DTYPE)
    lib.vn_scene_rtiow_final(out.ctypes.data_as(C.POINTER(vn_sphere)), n)
    return out


def random_scene(n: int, seed: int, S: float, mix: int) -> np.ndarray:
    """Synthetic scenes of BASELINE.json configs 4/5."""
    lib = L.load()
    out = np.zeros(n, SPHERE_DTYPE)
    lib.vn_scene_random(out.ctypes.data_as(C.POINTER(vn_sphere)), n, seed, S, mix)
    return out


class Scene:
    """Scene.h:10-84: m_spheres (structured array), m_aabbs, m_indices."""

    def __init__(self, spheres: np.ndarray | None = None):
        self.m_spheres = rtiow_final_scene() if spheres is None else np.ascontiguousarray(spheres, SPHERE_DTYPE)
        s = self.m_spheres
        r = np.abs(s["r"])
        self.m_aabbs = np.stack([s["cx"] - r, s["cy"] - r, s["cz"] - r, s["cx"] + r, s["cy"] + r, s["cz"] + r], axis=1)
        self.m_indices = np.arange(len(s), dtype=np.uint32)


class Camera:
    """camera.h:9-100 / Camera.cpp: thin-lens camera; the basis math runs in the library's C++ Camera class."""

    def __init__(self, origin=(0.0, 0.0, 0.0), vfov=45.0, aspect=1.6, aperture=0.0, focalLength=1.0):
        self.m_position = tuple(float(x) for x in origin)
        self.m_forward = (0.0, 0.0, -1.0)
        self.m_vfov, self.m_aspect, self.m_aperture, self.m_focalLength = float(vfov), float(aspect), float(aperture), float(focalLength)
        self.m_changed = True

    def SetForward(self, direction):
        self.m_forward = tuple(float(x) for x in direction)

    def SetFocalLength(self, length):
        if self.m_focalLength != float(length):
            self.m_focalLength = float(length)
            self.m_changed = True

    def SetAspect(self, aspect):
        if self.m_aspect != float(aspect):
            self.m_aspect = float(aspect)
            self.m_changed = True

    def SetPosition(self, origin):
        self.m_position = tuple(float(x) for x in origin)
        self.m_changed = True

    def GetFocalLength(self):
        return self.m_focalLength

    def GetLensRadius(self):
        return float(np.float32(self.m_aperture) * np.float32(0.5))

    def GetPosition(self):
        return self.m_position

    def frame(self):
        """origin, U, V, W, lens_radius as float32 arrays (Camera::UVWFrame + GetPosition + GetLensRadius)."""
        lib = L.load()
        o, u, v, w = L.c_float3(), L.c_float3(), L.c_float3(), L.c_float3()
        lens = C.c_float()
        lib.vn_camera_frame(L.c_float3(*self.m_position), L.c_float3(*self.m_forward), self.m_vfov, self.m_aspect,
                            self.m_aperture, self.m_focalLength, o, u, v, w, C.byref(lens))
        f = lambda a: np.array(list(a), np.float32)  # noqa: E731
        return f(o), f(u), f(v), f(w), np.float32(lens.value)

    def UVWFrame(self):
        _, u, v, w, _ = self.frame()
        return u, v, w

    def Changed(self):
        c = self.m_changed
        self.m_changed = False
        return c


def rtiow_camera(width: int, height: int) -> Camera:
    """The camera of Core.cpp:21-29,355 with the aspect of the requested frame."""
    cam = Camera((13.0, 2.0, 3.0), 20.0, width / height, 0.1, 10.0)
    cam.SetForward((0.0 - 13.0, 0.0 - 2.0, 0.0 - 3.0))
    return cam


# ---------------------------------------------------------------- the C ABI, one method per entry point
class Context:
    def __init__(self, device: int = 0):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.vn_create(device, C.byref(h))
        if rc != 0:
            raise Exception("vn_create failed (%d): %s" % (rc, self.lib.vn_last_error(None).decode()))
        self.h = h
        self.device = device
        self.width = self.height = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.vn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except BaseException:  # noqa: BLE001
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise Exception("%s failed (%d): %s" % (what, rc, self.lib.vn_last_error(self.h).decode()))

    def set_option(self, name: str, value: float):
        self._check(self.lib.vn_set_option(self.h, name.encode(), float(value)), "vn_set_option")

    def set_spheres(self, spheres: np.ndarray):
        s = np.ascontiguousarray(spheres, SPHERE_DTYPE)
        self._check(self.lib.vn_set_spheres(self.h, s.ctypes.data_as(C.POINTER(vn_sphere)), len(s)), "vn_set_spheres")

    def build_bvh(self):
        self._check(self.lib.vn_build_bvh(self.h), "vn_build_bvh")

    def update_spheres(self, spheres: np.ndarray):
        """Moved / re-coloured spheres (same count and order): upload + refit of the existing BVH."""
        s = np.ascontiguousarray(spheres, SPHERE_DTYPE)
        self._check(self.lib.vn_update_spheres(self.h, s.ctypes.data_as(C.POINTER(vn_sphere)), len(s)), "vn_update_spheres")

    def bvh_info(self) -> vn_bvh_info:
        info = vn_bvh_info()
        self._check(self.lib.vn_get_bvh_info(self.h, C.byref(info)), "vn_get_bvh_info")
        return info

    def read_bvh(self):
        info = self.bvh_info()
        nodes = np.zeros(info.num_nodes, NODE_DTYPE)
        order = np.zeros(info.num_spheres, np.uint32)
        self._check(self.lib.vn_read_bvh(self.h, _ptr(nodes), len(nodes), _ptr(order), len(order)), "vn_read_bvh")
        return nodes, order

    def read_grid(self):
        """(header dict, start[n_cells + 1] uint16, refs[n_refs] uint16) of the uniform grid + oversize list, or None when the scene has none."""
        hdr = np.zeros(26, np.uint32)
        if self._check_count(self.lib.vn_read_grid(self.h, _ptr(hdr), None, 0, None, 0), "vn_read_grid") == 0:
            return None
        f = hdr.view(np.float32)
        h = {"lo": f[0:3].copy(), "inv_cell": f[3:6].copy(), "cell": f[6:9].copy(), "hi": f[9:12].copy(), "res": hdr[12:15].copy(),
             "n_cells": int(hdr[15]), "n_refs": int(hdr[16]), "n_big": int(hdr[17]), "big": hdr[18:18 + int(hdr[17])].copy(), "raw": hdr.copy()}
        start = np.zeros(h["n_cells"] + 1, np.uint16)
        refs = np.zeros(max(h["n_refs"], 1), np.uint16)
        self._check_count(self.lib.vn_read_grid(self.h, _ptr(hdr), _ptr(start), len(start), _ptr(refs), len(refs)), "vn_read_grid")
        return h, start, refs[:h["n_refs"]]

    def read_huge(self):
        idx = (C.c_uint32 * 8)()
        n = self._check_count(self.lib.vn_read_huge(self.h, idx), "vn_read_huge")
        return [int(idx[i]) for i in range(n)]

    def last_accel(self) -> int:
        return int(self.lib.vn_last_accel(self.h))

    def _check_count(self, rc, what):
        if rc < 0:
            self._check(rc, what)
        return rc

    def read_wide_bvh(self):
        """(nodes[num_wide, 4, 2, 4] float32 -- child c = {lo.xyz, link}, {hi.xyz, count} --, levels); empty when the scene has no 4-wide nodes."""
        n, lev = C.c_uint32(), C.c_uint32()
        self._check(self.lib.vn_read_wide_bvh(self.h, None, 0, C.byref(n), C.byref(lev)), "vn_read_wide_bvh")
        nodes = np.zeros((n.value, 4, 2, 4), np.float32)
        if n.value:
            self._check(self.lib.vn_read_wide_bvh(self.h, _ptr(nodes), n.value, C.byref(n), C.byref(lev)), "vn_read_wide_bvh")
        return nodes, lev.value

    def morton_codes(self) -> np.ndarray:
        codes = np.zeros(self.bvh_info().num_spheres, np.uint32)
        self._check(self.lib.vn_morton_codes(self.h, _ptr(codes), len(codes)), "vn_morton_codes")
        return codes

    def resize(self, width: int, height: int):
        self._check(self.lib.vn_resize(self.h, width, height), "vn_resize")
        self.width, self.height = width, height

    def reset_accum(self):
        self._check(self.lib.vn_reset_accum(self.h), "vn_reset_accum")

    def make_params(self, camera: Camera, width: int, height: int, spp: int, subframe_index: int, max_depth: int,
                    accum_count: int = 0, image=None, flags: int = 0, rows=(0, 0)) -> vn_params:
        o, u, v, w, lens = camera.frame()
        p = vn_params()
        p.image = image
        p.width, p.height = width, height
        p.samples_per_pixel, p.subframe_index, p.max_depth, p.accum_count = spp, subframe_index, max_depth, accum_count
        p.origin, p.u, p.v, p.w = L.c_float3(*o), L.c_float3(*u), L.c_float3(*v), L.c_float3(*w)
        p.lens_radius = float(lens)
        p.row_begin, p.row_end = rows
        p.flags = flags
        return p

    def render(self, params: vn_params):
        self._check(self.lib.vn_render(self.h, C.byref(params)), "vn_render")
        self.width, self.height = params.width, params.height      # vn_render (re)sizes accum to the frame

    def render_subframes(self, params: vn_params, n: int, stride: int = 1):
        """Subframes params.subframe_index + k * stride, k < n, as n render() calls would leave them, in as few launches as the scene allows
        (vn_render_subframes[_strided]; Renderer::Draw called n times, Renderer.h:35-78)."""
        self._check(self.lib.vn_render_subframes_strided(self.h, C.byref(params), n, stride), "vn_render_subframes")
        self.width, self.height = params.width, params.height

    def tonemap(self, scale: float, image, flags: int = 0):
        self._check(self.lib.vn_tonemap(self.h, scale, image, flags), "vn_tonemap")

    def synchronize(self):
        self._check(self.lib.vn_synchronize(self.h), "vn_synchronize")

    def launch_timeline(self):
        """(start, tickets exhausted, end) of the last VN_COUNTERS launch in %globaltimer nanoseconds (vn_read_sched_counters)."""
        raw = (C.c_uint64 * 14)()
        self._check(self.lib.vn_read_sched_counters(self.h, raw), "vn_read_sched_counters")
        M = (1 << 64) - 1
        return M - raw[0], M - raw[1], raw[2]

    def timeline(self):
        """(lanes retired, segments of their last pixels) per 8.192 us bin of the last VN_COUNTERS launch (vn_read_timeline)."""
        raw = (C.c_uint32 * 2048)()
        self._check(self.lib.vn_read_timeline(self.h, raw), "vn_read_timeline")
        a = np.frombuffer(raw, np.uint32).copy()
        return a[:1024], a[1024:]

    def timeline_ex(self):
        """(lanes retired, segments of their last pixels, tiles fetched, segments shaded) per 8.192 us bin of the last VN_COUNTERS launch."""
        raw = (C.c_uint32 * 4096)()
        self._check(self.lib.vn_read_timeline_ex(self.h, raw), "vn_read_timeline_ex")
        a = np.frombuffer(raw, np.uint32).copy()
        return a[:1024], a[1024:2048], a[2048:3072], a[3072:]

    def stats(self) -> vn_stats:
        s = vn_stats()
        self._check(self.lib.vn_get_stats(self.h, C.byref(s)), "vn_get_stats")
        return s

    def reset_stats(self):
        self._check(self.lib.vn_reset_stats(self.h), "vn_reset_stats")

    def read_accum(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 4), np.float32)
        self._check(self.lib.vn_read_accum(self.h, _ptr(out)), "vn_read_accum")
        return out

    def write_accum(self, rgba: np.ndarray):
        a = np.ascontiguousarray(rgba, np.float32)
        assert a.size == self.width * self.height * 4
        self._check(self.lib.vn_write_accum(self.h, _ptr(a)), "vn_write_accum")

    def accum_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.vn_accum_device_ptr(self.h, C.byref(p)), "vn_accum_device_ptr")
        return p.value

    def set_accum_external(self, dev_ptr):
        self._check(self.lib.vn_set_accum_external(self.h, dev_ptr), "vn_set_accum_external")

    def reduce_tonemap_peers(self, peer_ptrs, scale: float, rows, image, flags: int = 0):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        self._check(self.lib.vn_reduce_tonemap_peers(self.h, arr, len(peer_ptrs), scale, rows[0], rows[1], image, flags),
                    "vn_reduce_tonemap_peers")

    def reduce_tonemap_peers_wait(self, peer_ptrs, scale: float, rows, sum_out, image, peer_flags, wait_value: int, flags: int = 0):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        fl = (C.c_void_p * len(peer_flags))(*peer_flags)
        self._check(self.lib.vn_reduce_tonemap_peers_wait(self.h, arr, len(peer_ptrs), scale, rows[0], rows[1], sum_out, image, fl, wait_value, flags),
                    "vn_reduce_tonemap_peers_wait")

    def sync_flags(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.vn_sync_flags(self.h, C.byref(p)), "vn_sync_flags")
        return p.value

    def signal(self, index: int, value: int):
        self._check(self.lib.vn_signal(self.h, index, value), "vn_signal")

    def wait_flags(self, flag_ptrs, value: int):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._check(self.lib.vn_wait_flags(self.h, arr, len(flag_ptrs), value), "vn_wait_flags")

    def check_flags(self):
        self._check(self.lib.vn_check_flags(self.h), "vn_check_flags")

    IPC_BYTES = 72

    def ipc_export(self, dev_ptr) -> bytes:
        """72 bytes for the peer's ipc_open: the CUDA IPC handle of the ALLOCATION that holds dev_ptr + dev_ptr's offset inside it (a
        tensor of torch's caching allocator usually sits in the middle of a segment it shares with other tensors)."""
        buf = (C.c_ubyte * 64)()
        off = C.c_uint64()
        self._check(self.lib.vn_ipc_export_at(self.h, dev_ptr, buf, C.byref(off)), "vn_ipc_export_at")
        return bytes(buf) + int(off.value).to_bytes(8, "little")

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_ubyte * 64).from_buffer_copy(handle[:64])
        off = int.from_bytes(handle[64:72], "little") if len(handle) >= 72 else 0
        key = bytes(handle[:64])
        opened = self.__dict__.setdefault("_ipc_opened", {})
        if key not in opened:              # one allocation may be opened only once per process
            p = C.c_void_p()
            self._check(self.lib.vn_ipc_open(self.h, buf, C.byref(p)), "vn_ipc_open")
            opened[key] = p.value
        return opened[key] + off

    def ipc_close(self, dev_ptr):
        self._check(self.lib.vn_ipc_close(self.h, dev_ptr), "vn_ipc_close")

    # ---- unit-level entry points
    def test_rng(self, v0, v1, n_draws: int):
        v0 = np.ascontiguousarray(v0, np.uint32)
        v1 = np.ascontiguousarray(v1, np.uint32)
        n = len(v0)
        seeds = np.zeros(n, np.uint32)
        lcg = np.zeros((n, n_draws), np.uint32)
        rnd = np.zeros((n, n_draws), np.float32)
        self._check(self.lib.vn_test_rng(self.h, _ptr(v0), _ptr(v1), n, n_draws, _ptr(seeds), _ptr(lcg), _ptr(rnd)), "vn_test_rng")
        return seeds, lcg, rnd

    def trace_rays(self, origins, dirs, flags: int = 0):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        n = len(o)
        t = np.zeros(n, np.float32)
        prim = np.zeros(n, np.int32)
        self._check(self.lib.vn_trace_rays(self.h, _ptr(o), _ptr(d), n, _ptr(t), _ptr(prim), flags), "vn_trace_rays")
        return t, prim

    def sort_pairs(self, keys, values, key_bits: int = 32):
        k = np.array(keys, np.uint32)
        v = np.array(values, np.uint32)
        self._check(self.lib.vn_sort_pairs(self.h, _ptr(k), _ptr(v), len(k), key_bits), "vn_sort_pairs")
        return k, v

    def make_color(self, rgb, flags: int = 0):
        a = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
        out = np.zeros((len(a), 4), np.uint8)
        self._check(self.lib.vn_test_make_color(self.h, _ptr(a), len(a), _ptr(out), flags), "vn_test_make_color")
        return out

    def scatter(self, material_type: int, albedo_fuzz_ir, dirs, normals, front, seeds, flags: int = 0):
        d = np.ascontiguousarray(dirs, np.float32)
        nr = np.ascontiguousarray(normals, np.float32)
        fr = np.ascontiguousarray(front, np.uint8)
        sd = np.ascontiguousarray(seeds, np.uint32)
        n = len(d)
        dout = np.zeros((n, 3), np.float32)
        ok = np.zeros(n, np.uint8)
        sout = np.zeros(n, np.uint32)
        m = (C.c_float * 4)(*[float(x) for x in albedo_fuzz_ir])
        self._check(self.lib.vn_test_scatter(self.h, material_type, m, _ptr(d), _ptr(nr), _ptr(fr), _ptr(sd), n,
                                             _ptr(dout), _ptr(ok), _ptr(sout), flags), "vn_test_scatter")
        return dout, ok, sout


class MultiContext:
    """vn_multi_*: one host thread, several devices (include/venusaur_b200.h).  Scene and BVH replicated, the subframes of a frame
    dealt round-robin to the devices, one fused peer reduce + tonemap per frame into the image on devices[0]."""

    def __init__(self, devices):
        self.lib = L.load()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = self.lib.vn_multi_create(arr, len(self.devices), C.byref(h))
        if rc != 0:
            raise Exception("vn_multi_create failed (%d): %s" % (rc, self.lib.vn_multi_last_error(None).decode()))
        self.h = h
        self.width = self.height = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.vn_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except BaseException:  # noqa: BLE001
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise Exception("%s failed (%d): %s" % (what, rc, self.lib.vn_multi_last_error(self.h).decode()))

    def set_option(self, name: str, value: float):
        self._check(self.lib.vn_multi_set_option(self.h, name.encode(), float(value)), "vn_multi_set_option")

    def set_spheres(self, spheres: np.ndarray):
        s = np.ascontiguousarray(spheres, SPHERE_DTYPE)
        self._check(self.lib.vn_multi_set_spheres(self.h, s.ctypes.data_as(C.POINTER(vn_sphere)), len(s)), "vn_multi_set_spheres")

    def build_bvh(self):
        self._check(self.lib.vn_multi_build_bvh(self.h), "vn_multi_build_bvh")

    make_params = Context.make_params

    def render(self, params: vn_params, n_subframes: int):
        self._check(self.lib.vn_multi_render(self.h, C.byref(params), n_subframes), "vn_multi_render")
        self.width, self.height = params.width, params.height

    def synchronize(self):
        self._check(self.lib.vn_multi_synchronize(self.h), "vn_multi_synchronize")

    def read_accum(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 4), np.float32)
        self._check(self.lib.vn_multi_read_accum(self.h, _ptr(out)), "vn_multi_read_accum")
        return out

    def stats(self) -> vn_stats:
        s = vn_stats()
        self._check(self.lib.vn_multi_get_stats(self.h, C.byref(s)), "vn_multi_get_stats")
        return s

    def subframes_accumulated(self) -> int:
        return int(self.lib.vn_multi_subframes_accumulated(self.h))


# ---------------------------------------------------------------- drop-in classes
class CUDAOutputBuffer:
    """CUDAOutputBuffer<uchar4> (CUDAOutputBuffer.h:65-101) in CUDA_DEVICE mode."""

    CUDA_DEVICE, GL_INTEROP, ZERO_COPY, CUDA_P2P = 0, 1, 2, 3

    def __init__(self, type_: int, width: int, height: int, device: int = 0):
        if type_ == self.GL_INTEROP:
            raise Exception("CUDAOutputBuffer: GL_INTEROP needs an OpenGL context; use CUDA_DEVICE or ZERO_COPY")
        self.lib = L.load()
        self.m_type, self.m_device_idx = type_, device
        self.m_width = self.m_height = 0
        self._dev = self._host = None
        self._stream = None
        self.resize(width, height)

    def _free(self):
        if self._dev or self._host:
            self.lib.vn_buffer_free(self.m_device_idx, self._dev, self._host, int(self.m_type == self.ZERO_COPY))
        self._dev = self._host = None

    def __del__(self):
        try:
            self._free()
        except BaseException:  # noqa: BLE001
            pass

    def setStream(self, stream):
        self._stream = stream

    def setDevice(self, device_idx: int):
        if device_idx == self.m_device_idx:
            return
        w, h = self.m_width, self.m_height
        self._free()                       # under the device that owns the pixels
        self.m_device_idx = device_idx
        self.m_width = self.m_height = 0
        if w and h:
            self.resize(w, h)

    def resize(self, width: int, height: int):
        width, height = max(1, width), max(1, height)
        if (width, height) == (self.m_width, self.m_height):
            return
        self._free()
        dev, host = C.c_void_p(), C.c_void_p()
        rc = self.lib.vn_buffer_alloc(self.m_device_idx, width * height * 4, int(self.m_type == self.ZERO_COPY), C.byref(dev), C.byref(host))
        if rc != 0:
            raise Exception("vn_buffer_alloc failed: " + self.lib.vn_last_error(None).decode())
        self._dev, self._host = dev, host
        self.m_width, self.m_height = width, height

    def map(self):
        return self._dev

    def unmap(self):
        rc = self.lib.vn_stream_synchronize(self.m_device_idx, self._stream)
        if rc != 0:
            raise Exception("vn_stream_synchronize failed: " + self.lib.vn_last_error(None).decode())

    def width(self):
        return self.m_width

    def height(self):
        return self.m_height

    def getHostPointer(self) -> np.ndarray:
        out = np.zeros((self.m_height, self.m_width, 4), np.uint8)
        rc = self.lib.vn_buffer_copy_to_host(self.m_device_idx, _ptr(out), self._dev, out.nbytes)
        if rc != 0:
            raise Exception("vn_buffer_copy_to_host failed: " + self.lib.vn_last_error(None).decode())
        return out


class Renderer:
    """Renderer.h:22-97: Init(scene, ptxSource) / Draw(camera, outputBuffer) / Cleanup()."""

    def __init__(self, device: int = 0):
        self.m_device = device
        self.m_devices: list[int] | None = None      # SetDevices: several devices behind the same Init / Draw / Cleanup
        self.m_subframes_per_draw = 1
        self.mctx: MultiContext | None = None
        self.ctx: Context | None = None
        self.m_samplesPerPixel = 16     # Renderer.h:53
        self.m_maxDepth = 4             # RayTracer.cu:172
        self.m_flags = 0
        self.m_subframe_index = 0
        self.m_accumulated = 0
        self._size = (0, 0)
        self.strict_accum = False       # True: literal blend weights of RayTracer.cu:208-213

    def SetDevices(self, devices, subframes_per_draw: int = 0):
        """Extension (not in the reference): render on several devices.  One Draw then advances the progressive render by
        `subframes_per_draw` subframes (default: one per device), i.e. it delivers the image a single device has after that many
        Draw calls -- up to float re-association, see vn_multi_render."""
        if self.ctx is not None or self.mctx is not None:
            raise Exception("Renderer::SetDevices must be called before Init")
        self.m_devices = [int(d) for d in devices]
        self.m_subframes_per_draw = int(subframes_per_draw) if subframes_per_draw else len(self.m_devices)

    def SetSubframesPerDraw(self, n: int):
        """Extension: on one device a Draw advances the progressive render by n subframes -- the buffers n Draw calls of the reference
        leave (Renderer.h:35-78), without the frames in between; one launch for all of them where the scene allows."""
        if self.m_devices is None:
            self.m_subframes_per_draw = max(1, int(n))

    def Init(self, scene: Scene, ptxSource: str = ""):
        if self.m_devices is not None:
            if self.mctx is None:
                self.mctx = MultiContext(self.m_devices)
            self.mctx.set_spheres(scene.m_spheres)
            self.mctx.build_bvh()
            self.m_subframe_index = self.m_accumulated = 0
            return
        if self.ctx is None:
            self.ctx = Context(self.m_device)
        self.ctx.set_spheres(scene.m_spheres)
        self.ctx.build_bvh()
        self.m_subframe_index = self.m_accumulated = 0

    def _draw_multi(self, camera: Camera, outputBuffer: CUDAOutputBuffer):
        size = (outputBuffer.width(), outputBuffer.height())
        if camera.Changed() or size != self._size:
            self._size = size
            self.m_subframe_index = self.m_accumulated = 0
        k = self.m_subframes_per_draw
        p = self.mctx.make_params(camera, size[0], size[1], self.m_samplesPerPixel, self.m_subframe_index + 1, self.m_maxDepth,
                                  accum_count=self.m_accumulated, image=outputBuffer.map(), flags=self.m_flags)
        self.mctx.render(p, k)                       # synchronous: the frame is in the buffer (on devices[0]) when it returns
        self.m_subframe_index += k
        self.m_accumulated += k

    def Draw(self, camera: Camera, outputBuffer: CUDAOutputBuffer):
        if self.mctx is not None:
            return self._draw_multi(camera, outputBuffer)
        if self.ctx is None:
            raise Exception("Renderer::Draw called before Init")
        size = (outputBuffer.width(), outputBuffer.height())
        resized = size != self._size
        if camera.Changed() or resized:
            if resized:
                self.ctx.resize(*size)
                self._size = size
            else:
                self.ctx.reset_accum()
            self.m_subframe_index = self.m_accumulated = 0
        outputBuffer.setStream(self.ctx.lib.vn_stream(self.ctx.h))
        self.m_subframe_index += 1
        count = self.m_subframe_index if self.strict_accum else self.m_accumulated
        p = self.ctx.make_params(camera, size[0], size[1], self.m_samplesPerPixel, self.m_subframe_index, self.m_maxDepth,
                                 accum_count=count, image=outputBuffer.map(), flags=self.m_flags | VN_ASYNC)
        n = self.m_subframes_per_draw if self.m_devices is None else 1
        if n > 1:                                    # SetSubframesPerDraw: n Draw calls without the frames in between (vn_render_subframes)
            self.ctx.render_subframes(p, n)
        else:
            self.ctx.render(p)
        outputBuffer.unmap()
        self.ctx.synchronize()
        self.m_subframe_index += n - 1
        self.m_accumulated += n

    def Cleanup(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None
        if self.mctx is not None:
            self.mctx.close()
            self.mctx = None
        self._size = (0, 0)
