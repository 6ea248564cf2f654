#!/usr/bin/env python
"""bench.py -- headline benchmark: Mrays/s (primary + secondary ray segments) on the RTIOW final scene at 1920x1080,
max depth 50, through the B200 kernels (BASELINE.json: metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference path's CPU restatement on the host cores)

A step = one Renderer::Draw (Renderer.h:35-78): one 16-spp subframe of the whole frame, subframe_index = 1, 2, 3, ...
(K = 64 steps is the 1024-spp workload of configs[1]).  Rank r of N renders subframes r+1, r+1+N, ... of the same frame
(sample-range sharding, scene + BVH replicated, no data-path collective) and the partial sums are combined ONCE, after
the last step, by a fused peer-memory reduce+tonemap kernel over NVLink (or an NCCL reduce with --reduce nccl).

Prints one JSON line (rank 0).  `value` = all segments traced by all ranks / max-over-ranks device time of the timed
region, inputs resident in HBM.  `e2e` = the same metric through the public API with HOST buffers: every step passes the
launch parameters from host memory and copies the uchar4 frame back to pinned host memory inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (width, height, spp per subframe, max_depth, scene)
    "c2": (1920, 1080, 16, 50, "rtiow"),
    "c1": (400, 225, 10, 50, "rtiow"),
    "c3": (3840, 2160, 16, 50, "rtiow"),
    "c4": (1920, 1080, 16, 50, "random1m"),
    "c5": (1920, 1080, 16, 64, "random16m"),
}
SCENES = {   # name: (count, seed, half-extent S, material mix) -- SURVEY 8(d) C4 / C5
    "random1m": (1_000_000, 0x5EED0001, 100.0, 0),
    "random16m": (16_000_000, 0x5EED0002, 250.0, 1),
}
METRIC = "Mrays/sec (primary+secondary) at 1080p RTIOW final scene, 1/2/4/8 B200"


def describe(name, steps):
    w, h, spp, depth, scene = WORKLOADS[name]
    scene_txt = "RTIOW final scene (486 spheres)" if scene == "rtiow" else "synthetic %dM random spheres%s" % (SCENES[scene][0] // 1_000_000, " (50%% dielectric)" if SCENES[scene][3] else "")
    return "%s %dx%d, %d spp/subframe x %d subframes, max depth %d" % (scene_txt, w, h, spp, steps, depth)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 <= t <= t1 + 0.1:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(x.split(",")[1]) for _, x in self.rows[-3:] if len(x.split(",")) > 2] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU baseline (oracle)
def sample_pixels(width, height, tiles=64, tile=32, seed=1234):
    import numpy as np
    rng = np.random.RandomState(seed)
    px = []
    for _ in range(tiles):
        x0 = int(rng.randint(0, max(1, width - tile)))
        y0 = int(rng.randint(0, max(1, height - tile)))
        ys, xs = np.mgrid[y0:min(height, y0 + tile), x0:min(width, x0 + tile)]
        px.append((ys * width + xs).ravel())
    return np.concatenate(px).astype(np.uint32)


def cpu_reference_rate(workload, subframes, threads=0, first_subframe=1):
    """Times the oracle (CPU restatement of the reference path, multi-threaded) on a bounded sample of the workload:
    64 seeded 32x32 tiles x spp x `subframes`.  Returns (Mrays/s, cores, sample description, seconds)."""
    import oracle_lib as ol
    width, height, spp, depth, scene = WORKLOADS[workload]
    spheres = ol.rtiow_final_scene() if scene == "rtiow" else ol.random_scene(*SCENES[scene])
    orc = ol.Oracle(spheres)
    S2 = 2.0 * SCENES[scene][2] if scene != "rtiow" else 0.0
    cam = ol.rtiow_camera(width, height) if scene == "rtiow" else ol.camera((0, 0, S2), (0, 0, -1.0), 40.0, width / height, 0.0, S2)
    px = sample_pixels(width, height)
    cores = threads or (os.cpu_count() or 1)
    segs, t0 = 0, time.perf_counter()
    for k in range(subframes):
        p = orc.params(cam, width, height, spp, first_subframe + k, depth, atten=ol.ATTEN_UNWIND, closest=ol.CLOSEST_BVH, threads=cores)
        _, st = orc.render_mean(p, pixels=px)
        segs += st.segments
    dt = time.perf_counter() - t0
    sample = "64 seeded 32x32-pixel tiles of the %dx%d frame x %d spp x %d subframes (%d paths)" % (width, height, spp, subframes, len(px) * spp * subframes)
    return segs / dt / 1e6, cores, sample, dt, segs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    cpu_reference_rate(args.workload, max(1, warmup))
    t0 = time.perf_counter()
    rate, cores, sample, dt, segs = cpu_reference_rate(args.workload, steps, first_subframe=1)
    out = {
        "impl": "reference",
        "metric": METRIC, "value": rate, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3 / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(args.workload, steps), "what": "CPU restatement of the reference path (oracle/, pinned to the reference's "
                   "own RayTracer.cu compiled for the host), own CPU BVH, all host threads; each step = one 16-spp subframe of a bounded sample"},
        "cpu_baseline": {"value": rate, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------- GPU arm
def _profile_numbers():
    """ncu numbers of the headline kernel, copied from the newest committed summary (profiles/rNN_trace_kernel.json): they are
    quoted in the roofline object with their source, never passed off as measured in this run."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_trace_kernel.json")))
    for f in reversed(files):
        try:
            return json.load(open(f)), os.path.relpath(f, ROOT)
        except (OSError, ValueError):
            continue
    return {}, None


def run_ours(args):
    import numpy as np
    import torch
    import venusaur_b200 as vb
    from venusaur_b200 import sharding
    from venusaur_b200 import VN_ACCUM_SUM, VN_ASYNC, VN_COUNTERS, VN_FAST, VN_IMAGE_HOST, VN_NO_TONEMAP, VN_WAVEFRONT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; venusaur_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    width, height, spp, depth, scene_name = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    ctx = vb.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
    if args.leaf_size:
        ctx.set_option("leaf_size", args.leaf_size)
    if args.threads:
        ctx.set_option("threads", args.threads)
    if args.blocks_per_sm:
        ctx.set_option("blocks_per_sm", args.blocks_per_sm)
    spheres = vb.rtiow_final_scene() if scene_name == "rtiow" else vb.random_scene(*SCENES[scene_name])
    ctx.set_spheres(spheres)
    ctx.build_bvh()                       # first build of the handle: includes the two arena allocations (cudaMalloc) of the builder
    build_first_ms = ctx.stats().ms_build
    ctx.build_bvh()                       # steady state: what a rebuild costs (moving spheres, Renderer::Init on a warm handle)
    build_ms = ctx.stats().ms_build
    info = ctx.bvh_info()
    if scene_name == "rtiow":
        cam = vb.rtiow_camera(width, height)
    else:
        S2 = 2.0 * SCENES[scene_name][2]
        cam = vb.Camera((0.0, 0.0, S2), 40.0, width / height, 0.0, S2)
        cam.SetForward((0.0, 0.0, -1.0))
    ctx.resize(width, height)
    kflag = {"wavefront": VN_WAVEFRONT, "persistent": 0}[args.kernel] | (VN_FAST if args.fast else 0)

    stream = torch.cuda.ExternalStream(ctx.lib.vn_stream(ctx.h), device=dev)      # the stream the kernels are launched on
    copy_stream = torch.cuda.Stream(device=dev)
    images = [torch.zeros((height, width, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
    image = images[0]
    accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)      # torch-owned so NCCL can reduce it
    ctx.set_accum_external(accum.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                 # > 126 MB L2
    host_images = [torch.empty((height, width, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

    def subframe_of(step):          # rank r renders subframes r+1, r+1+N, ... (1-based stream ids, Renderer.h:54)
        return sharding.subframes_for_rank(rank, world, step + 1)[step]

    def step_params(step, e2e=False):
        if world > 1:
            return ctx.make_params(cam, width, height, spp, subframe_of(step), depth, flags=kflag | VN_ACCUM_SUM | VN_NO_TONEMAP | VN_ASYNC)
        if e2e:        # frame k's copy to pinned host memory runs under frame k+1's kernel: two host buffers, alternating
            return ctx.make_params(cam, width, height, spp, subframe_of(step), depth, accum_count=step, image=host_images[step & 1].data_ptr(),
                                   flags=kflag | VN_IMAGE_HOST | VN_ASYNC)
        return ctx.make_params(cam, width, height, spp, subframe_of(step), depth, accum_count=step, image=image.data_ptr(), flags=kflag | VN_ASYNC)

    def barrier():
        # the library renders on its own non-blocking stream, which NCCL never waits on: drain it BEFORE the process-group barrier,
        # so that passing the barrier means every rank's kernels (e.g. its partial sums) are complete
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # peer mapping for the fused reduce+tonemap (one process per GPU => CUDA IPC handles, exchanged over the NCCL group): every rank's
    # partial-sum buffer and epoch flags, and rank 0's two image buffers
    peer_ptrs, peer_images, peer_flag0, peer_flag1 = None, None, None, None
    epoch = [0]
    if world > 1 and args.reduce == "peer":
        my_flags = ctx.sync_flags()
        mine = torch.tensor(list(ctx.ipc_export(accum.data_ptr())) + list(ctx.ipc_export(images[0].data_ptr())) + list(ctx.ipc_export(images[1].data_ptr()))
                            + list(ctx.ipc_export(my_flags)), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        peer_ptrs, flag_base = [], []
        B = ctx.IPC_BYTES                                 # handle of the allocation + offset of the tensor inside it
        for r in range(world):
            hb = bytes(allh[r].cpu().tolist())
            peer_ptrs.append(accum.data_ptr() if r == rank else ctx.ipc_open(hb[:B]))
            flag_base.append(my_flags if r == rank else ctx.ipc_open(hb[3 * B:4 * B]))
            if r == 0:
                peer_images = [images[i].data_ptr() for i in range(2)] if rank == 0 else [ctx.ipc_open(hb[B:2 * B]), ctx.ipc_open(hb[2 * B:3 * B])]
        peer_flag0 = [b for b in flag_base]              # flag 0: "my partial sums are complete"
        peer_flag1 = [b + 4 for b in flag_base]          # flag 1: "my reduce kernel is done (it no longer reads anybody's sums)"

    def finish_frame(n_sub, keep_sums=False, buf=0):
        """Combine the per-rank partial sums once per frame and tonemap (north star: one reduce per frame).  Peer path: no host
        barrier -- epoch flags in device memory order the GPUs (vn_signal / vn_reduce_tonemap_peers_wait / vn_wait_flags)."""
        if world == 1:
            return
        scale = 1.0 / float(n_sub)
        rows = sharding.row_slice(rank, world, height)
        if peer_ptrs is not None:
            epoch[0] += 1
            e = epoch[0]
            ctx.signal(0, e)                                      # behind my renders: my partial sums are complete
            ctx.reduce_tonemap_peers_wait(peer_ptrs, scale, rows, None if keep_sums else accum.data_ptr(), peer_images[buf], peer_flag0, e, VN_ASYNC)
            ctx.signal(1, e)                                      # my row slice of the image is written, I read nobody's sums any more
            ctx.wait_flags(peer_flag1, e)                         # everybody is: the frame on rank 0 is complete, the sums may change again
        else:
            ctx.synchronize()
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                ctx.tonemap(scale, images[buf].data_ptr(), 0)

    # ---- instrumented pass (untimed): V_node / V_sphere per segment for the roofline model
    ctx.render(ctx.make_params(cam, width, height, spp, 1, depth, flags=VN_COUNTERS | VN_NO_TONEMAP | (VN_FAST if args.fast else 0)))
    cst = ctx.stats()
    accel = ctx.last_accel() if args.kernel == "persistent" else 0      # 1 pair nodes, 2 wide nodes, 3 wide nodes from L2/HBM, 4 grid
    v_node = cst.node_visits / max(1, cst.segments)
    v_sphere = cst.sphere_tests / max(1, cst.segments)

    # ---- warm-up
    ctx.reset_accum()
    for s in range(W):
        ctx.render(step_params(s))
    finish_frame(max(1, W) * world)
    barrier()

    # ---- timed region: EXACTLY K steps, device-timed on the launching stream
    ctx.reset_accum()
    barrier()
    ctx.reset_stats()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ev_fin = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    barrier()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for s in range(K):
            flush.fill_(s & 0xFF)                                 # L2 flush between timed iterations (not timed)
            evs[s][0].record(stream)
            ctx.render(step_params(s))
            evs[s][1].record(stream)
        ev_fin[0].record(stream)
    finish_frame(K * world)
    with torch.cuda.stream(stream):
        ev_fin[1].record(stream)
    barrier()
    if peer_ptrs is not None:
        ctx.check_flags()
    t_wall1 = time.perf_counter()
    clk = clocks.stop(t_wall0, t_wall1)
    st = ctx.stats()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    fin_ms = ev_fin[0].elapsed_time(ev_fin[1]) if world > 1 else 0.0
    my_ms = sum(step_ms) + fin_ms
    my_segs = st.segments_total
    launches = st.kernel_launches_total

    # ---- end to end through the public API with HOST buffers, the same thing at every N: K frames, each frame = one subframe per GPU
    # with the launch parameters passed from host memory, (N > 1: one fused reduce + tonemap into rank 0's image,) and the uchar4 frame
    # copied to pinned host memory; frame k's copy runs under frame k+1's kernels (two buffers)
    e2e_ms, e2e_segs = None, 0
    if world == 1:
        for s in range(min(W, 3)):                               # untimed: first use of the copy stream and its staging buffers
            ctx.render(step_params(s, e2e=True))
        ctx.reset_accum()
        ctx.synchronize()
        ctx.reset_stats()
        t0 = time.perf_counter()
        for s in range(K):
            ctx.render(step_params(s, e2e=True))                 # params from host memory; uchar4 frame D2H into pinned memory
        ctx.synchronize()                                        # every frame has landed in host memory
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_segs = ctx.stats().segments_total
        assert int(host_images[(K - 1) & 1][..., 3].min()) == 255, "the last frame did not reach host memory"
    elif peer_ptrs is None:
        # --reduce nccl: the NCCL reduce sums in place, so the partial sums cannot go on accumulating: K steps + ONE reduce + the final frame
        ctx.reset_accum()
        barrier()
        ctx.reset_stats()
        t0 = time.perf_counter()
        for s in range(K):
            ctx.render(step_params(s))
        finish_frame(K * world)
        if rank == 0:
            host_images[0].copy_(images[0], non_blocking=False)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_segs = ctx.stats().segments_total
    else:
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        ctx.reset_accum()
        barrier()
        ctx.reset_stats()
        t0 = time.perf_counter()
        for s in range(K):
            if rank == 0 and s >= 2:
                stream.wait_event(copied[s & 1])                 # the image buffer of this frame was last read by the copy of two frames ago
            ctx.render(step_params(s))
            finish_frame((s + 1) * world, keep_sums=True, buf=s & 1)
            if rank == 0:
                ready = torch.cuda.Event()
                ready.record(stream)
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):
                    host_images[s & 1].copy_(images[s & 1], non_blocking=True)
                    copied[s & 1].record(copy_stream)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if peer_ptrs is not None:
            ctx.check_flags()
        e2e_segs = ctx.stats().segments_total
        if rank == 0:
            assert int(host_images[(K - 1) & 1][..., 3].min()) == 255, "the last frame did not reach host memory"

    # ---- the same K subframes as ONE call (vn_render_subframes: Renderer::Draw K times without looking at the frames in between).  Scenes rendered
    # from shared memory take up to 64 subframes per launch -- the launch drains once instead of once per subframe; the accumulation buffer and
    # the final frame are bit for bit those of the K single calls (tests/test_gpu_parity.py::test_subframes_in_one_launch...).  Reported next to
    # `value`, which stays one launch per step.
    grouped = None
    if world == 1 and args.kernel == "persistent" and K > 1:
        ctx.reset_accum()
        ctx.synchronize()
        ctx.reset_stats()
        ev_g = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        with torch.cuda.stream(stream):
            flush.fill_(7)
            ev_g[0].record(stream)
            ctx.render_subframes(ctx.make_params(cam, width, height, spp, subframe_of(0), depth, accum_count=0, image=image.data_ptr(), flags=kflag | VN_ASYNC), K)
            ev_g[1].record(stream)
        ctx.synchronize()
        gst = ctx.stats()
        grouped = {"subframes": K, "ms": ev_g[0].elapsed_time(ev_g[1]), "segments": gst.segments_total, "launches": gst.kernel_launches_total}

    # ---- strong scaling: ONE fixed frame of `--strong-subframes` subframes (default 64 = the 1024-spp frame of configs[1]) split over the
    # ranks, cold tile order included (the first launch of the view counts the tile costs), one reduce at the end
    strong = None
    if args.strong_subframes > 0:
        S = args.strong_subframes
        mine_sub = list(range(rank + 1, S + 1, world))
        ctx.set_option("tile_order", float(dict(kv.split("=", 1) for kv in args.opt if "=" in kv).get("tile_order", 1)))   # forget the view's tile costs
        ctx.reset_accum()
        barrier()
        ctx.reset_stats()
        ev_s = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        with torch.cuda.stream(stream):
            ev_s[0].record(stream)
            # this rank's subframes (rank + 1, rank + 1 + N, ...) as ONE call: a launch of their own for the view's first subframe (tile costs),
            # then one launch for the rest where the scene is rendered from shared memory (vn_render_subframes_strided)
            if world > 1 and mine_sub:
                ctx.render_subframes(ctx.make_params(cam, width, height, spp, mine_sub[0], depth, flags=kflag | VN_ACCUM_SUM | VN_NO_TONEMAP | VN_ASYNC), len(mine_sub), stride=world)
            elif mine_sub:
                ctx.render_subframes(ctx.make_params(cam, width, height, spp, 1, depth, accum_count=0, image=image.data_ptr(), flags=kflag | VN_ASYNC), S)
        finish_frame(S)
        with torch.cuda.stream(stream):
            ev_s[1].record(stream)
        barrier()
        if peer_ptrs is not None:
            ctx.check_flags()
        strong = {"subframes": S, "ms": ev_s[0].elapsed_time(ev_s[1]), "segments": ctx.stats().segments_total}

    # ---- multi-GPU parity: the reduced strong-scaling frame against rank 0's own running mean over the same subframes
    parity = None
    if world > 1 and strong is not None:
        rows = sharding.row_slice(rank, world, height)
        full = [torch.zeros_like(accum) for _ in range(world)]
        dist.all_gather(full, accum)
        if rank == 0:
            S = strong["subframes"]
            got = torch.zeros_like(accum)
            if args.reduce == "peer":
                for r in range(world):
                    a, b = sharding.row_slice(r, world, height)
                    got[a:b] = full[r][a:b]
            else:
                got = accum.clone()
            got_mean = (got[..., :3] / float(S)).cpu().numpy()
            got_img = images[0].cpu().numpy()
            ref = vb.Context(local_rank)
            ref.set_spheres(spheres)
            ref.build_bvh()
            ref_img = np.zeros((height, width, 4), np.uint8)
            for k in range(S):
                ref.render(ref.make_params(cam, width, height, spp, k + 1, depth, accum_count=k, image=ref_img.ctypes.data,
                                           flags=kflag | (VN_IMAGE_HOST if k == S - 1 else VN_NO_TONEMAP)))
            want = ref.read_accum()[..., :3]
            ref.close()
            rel = np.abs(got_mean - want) / np.maximum(np.abs(want), 1e-6)
            d = np.abs(got_img.astype(np.int32) - ref_img.astype(np.int32))
            parity = {"max_rel": float(rel.max()), "max_code": int(d.max()), "pixels_off_by_one": float((d.max(axis=-1) > 0).mean()),
                      "what": "the %d-GPU frame (%d subframes, sample-range sharding + one %s reduce) against one GPU's running mean over the same subframes: "
                              "float re-association only" % (world, S, args.reduce)}

    # ---- max over ranks / sums
    tot_segs, max_ms, e2e_max_ms, e2e_tot = my_segs, my_ms, e2e_ms, e2e_segs
    tot_launch = launches
    strong_ms, strong_segs = (strong["ms"], strong["segments"]) if strong else (0.0, 0)
    if dist is not None:
        t = torch.tensor([float(my_segs), float(e2e_segs), float(launches), float(strong_segs)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        m = torch.tensor([my_ms, e2e_ms, strong_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        tot_segs, e2e_tot, tot_launch, strong_segs = int(t[0].item()), int(t[1].item()), int(t[2].item()), int(t[3].item())
        max_ms, e2e_max_ms, strong_ms = float(m[0].item()), float(m[1].item()), float(m[2].item())

    if rank == 0:
        value = tot_segs / (max_ms * 1e-3) / 1e6
        e2e_value = e2e_tot / (e2e_max_ms * 1e-3) / 1e6
        props = torch.cuda.get_device_properties(dev)
        n_sm = props.multi_processor_count
        # which loop structure the wide-node path kernel ran with (library defaults unless --opt overrides them)
        _o = dict(kv.split("=", 1) for kv in args.opt if "=" in kv)
        _ad, _an = int(float(_o.get("async_done", 26))), int(float(_o.get("async_node", 0)))
        _to = int(float(_o.get("tile_order", 1)))
        _lean = int(float(_o.get("lean", 1)))
        tile_order = (None if args.kernel != "persistent" else
                      "row-major tickets" if _to == 0 else
                      "8x4-pixel tiles handed out most expensive first (ray segments per tile counted by the first launch of the view, i.e. during warm-up; "
                      "the strong_scaling frame includes that collecting launch)")
        schedule = None
        if args.kernel == "persistent" and accel == 2:
            schedule = ("k_render_persistent: every round waits for its slowest ray" if _ad == 0 else
                        "%s, %s: a traversal burst ends when %d lanes hold a finished ray%s" % ("k_render_lean" if (_lean and _an == 0 and int(float(_o.get("warp_tiles", 1)))) else "k_render_async",
                        "phase form" if _an == 0 else "voted turns", _ad, "; warps own whole 8x4 tiles" if (_an == 0 and int(float(_o.get("warp_tiles", 1)))) else ""))
            if _lean and _ad and float(_o.get("split_tail", 0.5)) > 0:
                schedule += "; a frame = two launches of the kernel: the tiles that saw nothing but sky when the view's costs were collected go to a second launch on another stream that overlaps the first one's drain"
        elif args.kernel == "persistent" and accel == 1 and not info.scene_in_smem:
            schedule = ("k_render_lean<global>: asynchronous shading, voted node / leaf turns, a burst ends when %d lanes hold a finished ray" % int(float(_o.get("global_done", 16)))
                        if (_lean and _ad) else "k_render_persistent: every round waits for its slowest ray")
            if _lean and _ad and float(_o.get("steal", 1)) > 0:
                schedule += "; once the tile tickets are exhausted idle lanes take single samples of the pixels other lanes of their warp still hold (lean_drain)"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        # issue-slot roofline: thread-instructions per segment from the budget model of SURVEY 8(d)
        # with V_node / V_sphere measured by the instrumented kernel on this very workload.
        # Coefficients: pair nodes = SURVEY's budget; 4-wide nodes = SASS counts of the shipped kernel (node step 57 instructions,
        # one sphere test ~100 with its leaf / pop overhead and IEEE sqrt + div, shade + RNG + camera + ray set-up ~230 per segment).
        wide = accel == 2
        if accel == 4:      # uniform grid + oversize list: V_node = cell steps (25 instructions each incl. the vote), 45 per sphere test,
            i_seg = 25.0 * v_node + 45.0 * v_sphere + 300.0      # + shade / RNG / camera / ray-box clip and DDA set-up
        else:
            i_seg = (57.0 * v_node + 100.0 * v_sphere + 230.0) if wide else (40.0 * v_node + 30.0 * v_sphere + 150.0)
        f_clk = (clk.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        peak_tinst = 32 * 4 * n_sm * f_clk / 1e12
        per_gpu_rate = (my_segs / (sum(step_ms) * 1e-3))
        achieved_tinst = per_gpu_rate * i_seg / 1e12
        prof, prof_file = _profile_numbers()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes of one launch: accum read (blend) + write + uchar4 image per pixel; the scene lives in smem
        algo_bytes = width * height * (16 + 16 + 4)
        mean_step_ms = sum(step_ms) / len(step_ms)
        n_prims = int(info.num_spheres)
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(args.workload, K), "parallelism": "subframe(sample-range) sharding x%d, scene+BVH replicated" % world,
                       "kernel": args.kernel, "schedule": schedule, "tile_order": tile_order, "accel": {0: "n/a", 1: "BVH pair nodes", 2: "BVH 4-wide octant-sorted nodes (shared memory)", 3: "BVH 4-wide nodes (L2/HBM)", 4: "uniform grid + oversize list (shared memory)"}[accel], "build": ("VN_FAST (relaxed numerics; not within the image tolerance)" if args.fast else "default IEEE build, bit-identical to the oracle"), "l2_flush": "256 MiB fill between timed steps",
                       "scene_in_smem": bool(info.scene_in_smem), "bvh_nodes": int(info.num_nodes), "leaf_size": int(info.max_leaf_size),
                       "bvh_build_ms": build_ms, "reduce": (args.reduce if world > 1 else "none"), "reduce_ms": fin_ms},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(vb._lib.vn_params) * world,
                    "d2h_bytes_per_step": width * height * 4,
                    "what": ("K frames through the public API with host buffers; a frame = one 16-spp subframe per GPU with launch params from host memory%s + the uchar4 "
                             "frame copied to pinned host memory (the copy of frame k overlaps the kernels of frame k+1)"
                             % ("" if world == 1 else ", one fused peer reduce + tonemap into rank 0's image (device-side epoch flags, no host barrier)"))},
            "gpu_launches": int(tot_launch),
            "segments": int(tot_segs), "segments_per_path": tot_segs / float(width * height * spp * K * world),
            "roofline": {"bound": "issue", "achieved": achieved_tinst, "peak": peak_tinst, "unit": "Tinst/s (thread instructions)",
                         "frac": achieved_tinst / peak_tinst, "traffic": prof.get("dram_bytes_per_launch"),
                         "model": "I_seg = %s = %.0f thread-instructions/segment (V_node=%.2f, V_sphere=%.2f measured in this run); "
                                  "peak = 32 lanes x 4 schedulers x %d SMs x %.0f MHz (median SM clock during the run)"
                                  % ("25*V_cell + 45*V_sphere + 300 (uniform grid + oversize list)" if accel == 4 else "57*V_node + 100*V_sphere + 230 (4-wide nodes)" if wide else "40*V_node + 30*V_sphere + 150 (pair nodes)", i_seg, v_node, v_sphere, n_sm, f_clk / 1e6),
                         "ncu": {"source": prof_file, "note": "copied from the committed ncu summary of the same kernel, not measured in this run",
                                 "thread_inst_per_segment": prof.get("thread_inst_per_segment"), "issue_slot_utilisation": prof.get("issue_slot_utilisation"),
                                 "avg_active_lanes": prof.get("avg_active_threads_per_inst")}},
            "roofline_hbm": {"bound": "hbm", "achieved": algo_bytes / (mean_step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": algo_bytes / (mean_step_ms * 1e-3) / 1e9 / hbm_peak, "traffic": prof.get("dram_bytes_per_launch"),
                             "note": "accum RMW + uchar4 store only (36 B/pixel/launch); BVH + spheres are staged in shared memory, so this "
                                     "path is not HBM-bound (of measured %s)" % ("peak" if peaks else "fallback")},
            # north-star kernel (1): the LBVH builder.  Algorithmic bytes per sphere: 2 x 36 B (bounds + Morton passes read the records),
            # 68 B for the 30-bit key-index sort (4 onesweep passes of 16 B + the histogram read), gather 36 B in + 65 B out (geom, material,
            # type, two leaf-box float4), Karras 4 B in + 16 B out, refit 2 x 64 B, pack 2 x 32 B out + 64 B in = 493 B
            "bvh_build": {"ms": build_ms, "ms_first_build_incl_allocation": build_first_ms, "spheres": n_prims, "mprims_per_s": n_prims / max(build_ms, 1e-6) / 1e3,
                          "algorithmic_bytes_per_sphere": 493, "achieved_gbs": n_prims * 493.0 / max(build_ms, 1e-6) / 1e6, "peak_gbs": hbm_peak,
                          "frac": n_prims * 493.0 / max(build_ms, 1e-6) / 1e6 / hbm_peak,
                          "note": "event time of the whole build (15-18 launches, one host sync for the node count); for small scenes the launch chain, not bandwidth, is the time"},
        }
        if strong is not None:
            out["strong_scaling"] = {"subframes": strong["subframes"], "ms": strong_ms, "value": strong_segs / (strong_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                                     "what": "ONE fixed frame of %d x %d spp split over the %d GPUs (sample ranges), device time from the first launch to the end of the reduce, "
                                             "max over ranks; includes the cold first launch of the view (row-major tiles, tile costs collected); each rank's remaining subframes are one vn_render_subframes_strided call (one launch)" % (strong["subframes"], spp, world)}
        if parity is not None:
            out["parity"] = parity
        if grouped is not None:
            out["subframes_in_one_call"] = {"subframes": grouped["subframes"], "ms": grouped["ms"], "ms_per_subframe": grouped["ms"] / grouped["subframes"],
                                            "value": grouped["segments"] / (grouped["ms"] * 1e-3) / 1e6, "unit": "Mrays/s", "gpu_launches": grouped["launches"],
                                            "what": "the same %d subframes through ONE vn_render_subframes call (up to 64 subframes per launch of the path kernel for scenes "
                                                    "rendered from shared memory, one drain per launch; tonemap once at the end); device time of the call; accumulation buffer and "
                                                    "frame bit-identical to the %d single calls timed in `value`" % (grouped["subframes"], grouped["subframes"])}
        if not info.scene_in_smem:
            # scenes traversed from L2/HBM (C4 / C5): SURVEY 8(d) byte model, one pair visit = 2 x 32-byte nodes, one sphere = 32 bytes.
            # configs[3] (1 M spheres, 57 MB of nodes + spheres) is L2-resident: ncu measures 92-94 % L2 hits and < 10 GB/s of DRAM traffic, so its
            # roofline is anchored on L2 bandwidth, with the ncu-measured L2 -> L1 throughput as the peak; configs[4] (16 M spheres,
            # 0.9 GB) is anchored on HBM with ncu's dram__bytes as `traffic`.  Both kernels are in fact bound by SIMT divergence and latency.
            # The shipped kernel steps through QUANTISED pairs (option "qnodes", the default): one 32-byte record per pair visit, so the
            # algorithm's bytes are 32 per pair visit; with SURVEY 8(d)'s two packed 32-byte nodes per visit (qnodes=0) it is 64.
            quantised = float(dict(kv.split("=", 1) for kv in args.opt if "=" in kv).get("qnodes", 1)) != 0.0
            node_bytes = 32.0 if quantised else 64.0
            b_seg = node_bytes * v_node + 32.0 * v_sphere
            out["roofline_issue"] = out["roofline"]
            sub = prof.get("c4" if n_prims <= 2_000_000 else "c5", {})         # the ncu summary of THIS kernel on this scene (profiles/r02_trace_kernel.json)
            out["roofline_issue"]["traffic"] = sub.get("dram_bytes_per_launch")
            out["roofline_issue"]["ncu"] = {"source": prof_file, "note": "copied from the committed ncu summary of the same kernel on the same scene, not measured in this run",
                                            "thread_inst_per_segment": sub.get("thread_inst_per_segment"), "issue_slot_utilisation": sub.get("issue_slot_utilisation"),
                                            "avg_active_lanes": sub.get("avg_active_threads_per_inst"), "l2_sector_hit_rate_pct": sub.get("lts__t_sector_hit_rate.pct"),
                                            "dram_bytes_per_launch": sub.get("dram_bytes_per_launch"), "l2_bytes_from_sms_per_launch": sub.get("l2_bytes_from_sms_per_launch")}
            l2_resident = n_prims <= 2_000_000
            # ncu's L2 peak for reads that come from the SMs: lts__t_sectors_srcunit_tex.peak_sustained = 2 sectors / cycle / slice; the capture in
            # profiles/r02_c4_trace_kernel_ncu.txt reads 78.3 sectors/ns = 10.85 % of it, i.e. a peak of 23.1 TB/s
            peak_bw = 23100.0 if l2_resident else hbm_peak
            out["roofline"] = {"bound": "l2" if l2_resident else "hbm", "achieved": per_gpu_rate * b_seg / 1e9, "peak": peak_bw, "unit": "GB/s",
                               "frac": per_gpu_rate * b_seg / 1e9 / peak_bw, "traffic": prof.get("c4_l2_bytes_per_launch" if l2_resident else "c5_dram_bytes_per_launch"),
                               "model": "B_seg = %.0f*V_node + 32*V_sphere = %.0f algorithmic bytes/segment (%s; SURVEY 8d's figure for two packed 32-byte nodes per visit would be %.0f) (V_node=%.2f pair visits, V_sphere=%.2f measured in this run); %s"
                                        % (node_bytes, b_seg, "one 32-byte quantised pair per visit" if quantised else "two packed 32-byte nodes per visit", 64.0 * v_node + 32.0 * v_sphere, v_node, v_sphere, "served by L2 (ncu: 83-94 %% L2 hits); `traffic` = ncu's L2 sectors from the SMs x 32 B per launch; peak = ncu's lts__t_sectors_srcunit_tex peak (2 sectors/cycle/slice = 23.1 TB/s), see profiles/README.md" if l2_resident
                                           else "served by HBM (ncu: 45-53 %% L2 hits); `traffic` = ncu's dram__bytes per launch; peak = measured HBM copy bandwidth")}
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample, dt, _ = cpu_reference_rate(args.workload, 160 if scene_name == "rtiow" else 8)
            out["cpu_baseline"] = {"value": rate, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample, "seconds": dt}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "wavefront"])
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--leaf-size", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-subframes", type=int, default=-1, help="subframes of the fixed strong-scaling frame (0 = skip; default 64 for the RTIOW workloads, 0 otherwise)")
    ap.add_argument("--fast", action="store_true", help="opt-in relaxed-numerics kernels (VN_FAST)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (vn_set_option), repeatable")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)
    if args.strong_subframes < 0:
        args.strong_subframes = 64 if WORKLOADS[args.workload][4] == "rtiow" and args.kernel == "persistent" else 0
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
