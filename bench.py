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
def run_ours(args):
    import numpy as np
    import torch
    import venusaur_b200 as vb
    from venusaur_b200 import sharding
    from venusaur_b200 import VN_ACCUM_SUM, VN_ASYNC, VN_COUNTERS, VN_FAST, VN_IMAGE_HOST, VN_NO_TONEMAP, VN_POOL, VN_SLOTS, VN_WAVEFRONT

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
    ctx.build_bvh()
    info = ctx.bvh_info()
    build_ms = ctx.stats().ms_build
    if scene_name == "rtiow":
        cam = vb.rtiow_camera(width, height)
    else:
        S2 = 2.0 * SCENES[scene_name][2]
        cam = vb.Camera((0.0, 0.0, S2), 40.0, width / height, 0.0, S2)
        cam.SetForward((0.0, 0.0, -1.0))
    ctx.resize(width, height)
    kflag = {"wavefront": VN_WAVEFRONT, "pool": VN_POOL, "persistent": 0, "slots": VN_SLOTS}[args.kernel] | (VN_FAST if args.fast else 0)
    for opt in ("pool_slots", "pool_threads", "pool_service", "pool_leaf_batch"):
        if getattr(args, opt):
            ctx.set_option(opt, getattr(args, opt))

    stream = torch.cuda.ExternalStream(ctx.lib.vn_stream(ctx.h), device=dev)      # the stream the kernels are launched on
    image = torch.zeros((height, width, 4), dtype=torch.uint8, device=dev)
    accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)      # torch-owned so NCCL can reduce it
    ctx.set_accum_external(accum.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                 # > 126 MB L2
    host_images = [torch.empty((height, width, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    host_image = host_images[0]

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

    # peer mapping for the fused reduce+tonemap (one process per GPU => CUDA IPC handles, exchanged over the NCCL group)
    peer_ptrs, peer_image = None, None
    if world > 1 and args.reduce == "peer":
        mine = torch.tensor(list(ctx.ipc_export(accum.data_ptr())) + list(ctx.ipc_export(image.data_ptr())), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        peer_ptrs = []
        for r in range(world):
            hb = bytes(allh[r].cpu().tolist())
            peer_ptrs.append(accum.data_ptr() if r == rank else ctx.ipc_open(hb[:64]))
            if r == 0:
                peer_image = image.data_ptr() if rank == 0 else ctx.ipc_open(hb[64:])

    def finish_frame(n_sub):
        """Combine the per-rank partial sums once per frame and tonemap (north star: one reduce per frame)."""
        if world == 1:
            return
        scale = 1.0 / float(n_sub)
        rows = sharding.row_slice(rank, world, height)
        if peer_ptrs is not None:
            barrier()                                             # all partial sums complete before peers read them
            ctx.reduce_tonemap_peers(peer_ptrs, scale, rows, peer_image, 0)      # my row slice: load N peers, store into rank 0's image
            barrier()
        else:
            ctx.synchronize()
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                ctx.tonemap(scale, image.data_ptr(), 0)

    # ---- instrumented pass (untimed): V_node / V_sphere per segment for the roofline model
    ctx.render(ctx.make_params(cam, width, height, spp, 1, depth, flags=VN_COUNTERS | VN_NO_TONEMAP | (VN_FAST if args.fast else 0) | (VN_SLOTS if args.kernel == "slots" else 0)))
    cst = ctx.stats()
    sched = ctx.sched_counters() if args.kernel == "slots" else None
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
    seg_steps = []
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
    t_wall1 = time.perf_counter()
    clk = clocks.stop(t_wall0, t_wall1)
    st = ctx.stats()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    fin_ms = ev_fin[0].elapsed_time(ev_fin[1]) if world > 1 else 0.0
    my_ms = sum(step_ms) + fin_ms
    my_segs = st.segments_total
    launches = st.kernel_launches_total

    # ---- end-to-end through the public API with host buffers (single GPU: Renderer::Draw + getHostPointer per step)
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
    else:
        # multi-GPU e2e: the same K steps + reduce, plus rank 0 reading the final frame back to pinned host memory
        ctx.reset_accum()
        barrier()
        ctx.reset_stats()
        t0 = time.perf_counter()
        for s in range(K):
            ctx.render(step_params(s))
        finish_frame(K * world)
        if rank == 0:
            host_image.copy_(image, non_blocking=False)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        e2e_segs = ctx.stats().segments_total

    # ---- max over ranks / sums
    tot_segs, max_ms, e2e_max_ms, e2e_tot = my_segs, my_ms, e2e_ms, e2e_segs
    tot_launch = launches
    if dist is not None:
        t = torch.tensor([float(my_segs), float(e2e_segs), float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        m = torch.tensor([my_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        tot_segs, e2e_tot, tot_launch = int(t[0].item()), int(t[1].item()), int(t[2].item())
        max_ms, e2e_max_ms = float(m[0].item()), float(m[1].item())

    if rank == 0:
        value = tot_segs / (max_ms * 1e-3) / 1e6
        e2e_value = e2e_tot / (e2e_max_ms * 1e-3) / 1e6
        props = torch.cuda.get_device_properties(dev)
        n_sm = props.multi_processor_count
        # which loop structure the wide-node path kernel ran with (library defaults unless --opt overrides them)
        _o = dict(kv.split("=", 1) for kv in args.opt if "=" in kv)
        _ad, _an = int(float(_o.get("async_done", 26))), int(float(_o.get("async_node", 0)))
        _to = int(float(_o.get("tile_order", 1)))
        tile_order = (None if args.kernel != "persistent" else
                      "row-major tickets" if _to == 0 else
                      "8x4-pixel tiles handed out most expensive first (ray segments per tile counted by the first launch of the view, i.e. during warm-up)")
        schedule = None
        if args.kernel == "persistent" and accel == 2:
            schedule = ("k_render_persistent: every round waits for its slowest ray" if _ad == 0 else
                        "k_render_async, %s: a traversal burst ends when %d lanes hold a finished ray%s" % ("phase form" if _an == 0 else "voted turns", _ad,
                        "; warps own whole 8x4 tiles" if (_an == 0 and int(float(_o.get("warp_tiles", 1)))) else ""))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        # issue-slot roofline: thread-instructions per segment from the budget model of SURVEY 8(d)
        # with V_node / V_sphere measured by the instrumented kernel on this very workload.
        # Coefficients: pair nodes = SURVEY's budget; 4-wide nodes = SASS counts of the shipped kernel (node step 53 instructions,
        # one sphere test ~100 with its leaf / pop overhead and IEEE sqrt + div, shade + RNG + camera + ray set-up ~230 per segment;
        # profiles/r01s3_path_kernel_ncu.txt: 694 thread-instructions per segment measured, this model gives 697).
        wide = accel == 2
        if accel == 4:      # uniform grid + oversize list: V_node = cell steps (25 instructions each incl. the vote), 45 per sphere test,
            i_seg = 25.0 * v_node + 45.0 * v_sphere + 300.0      # + shade / RNG / camera / ray-box clip and DDA set-up
        else:
            i_seg = (53.0 * v_node + 100.0 * v_sphere + 230.0) if wide else (40.0 * v_node + 30.0 * v_sphere + 150.0)
        f_clk = (clk.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        peak_tinst = 32 * 4 * n_sm * f_clk / 1e12
        per_gpu_rate = (my_segs / (sum(step_ms) * 1e-3))
        achieved_tinst = per_gpu_rate * i_seg / 1e12
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01s3_trace_kernel.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes of one launch: accum read (blend) + write + uchar4 image per pixel; the scene lives in smem
        algo_bytes = width * height * (16 + 16 + 4)
        mean_step_ms = sum(step_ms) / len(step_ms)
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(args.workload, K), "parallelism": "subframe(sample-range) sharding x%d, scene+BVH replicated" % world,
                       "kernel": args.kernel, "schedule": schedule, "tile_order": tile_order, "accel": {0: "n/a", 1: "BVH pair nodes", 2: "BVH 4-wide octant-sorted nodes (shared memory)", 3: "BVH 4-wide nodes (L2/HBM)", 4: "uniform grid + oversize list (shared memory)"}[accel], "build": ("VN_FAST (relaxed numerics; not within the image tolerance)" if args.fast else "default IEEE build, bit-identical to the oracle"), "l2_flush": "256 MiB fill between timed steps",
                       "scene_in_smem": bool(info.scene_in_smem), "bvh_nodes": int(info.num_nodes), "leaf_size": int(info.max_leaf_size),
                       "bvh_build_ms": build_ms, "reduce": (args.reduce if world > 1 else "none"), "reduce_ms": fin_ms,
                       "sched": ({k: [v[0], round(v[1], 2)] for k, v in sched.items()} if sched else None)},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(vb._lib.vn_params),
                    "d2h_bytes_per_step": (width * height * 4 if world == 1 else width * height * 4 // max(1, K)),
                    "what": "vn_render with launch params from host memory + uchar4 frame copied to pinned host memory every step (the copy of frame k overlaps the kernel of frame k+1)" if world == 1
                            else "K steps + one reduce + final frame D2H on rank 0"},
            "gpu_launches": int(tot_launch),
            "segments": int(tot_segs), "segments_per_path": tot_segs / float(width * height * spp * K * world),
            "roofline": {"bound": "issue", "achieved": achieved_tinst, "peak": peak_tinst, "unit": "Tinst/s (thread instructions)",
                         "frac": achieved_tinst / peak_tinst, "traffic": prof.get("dram_bytes_per_launch"),
                         "model": "I_seg = %s = %.0f thread-instructions/segment (V_node=%.2f, V_sphere=%.2f measured); "
                                  "peak = 32 lanes x 4 schedulers x %d SMs x %.0f MHz (median SM clock during the run)"
                                  % ("25*V_cell + 45*V_sphere + 300 (uniform grid + oversize list)" if accel == 4 else "53*V_node + 100*V_sphere + 230 (4-wide nodes)" if wide else "40*V_node + 30*V_sphere + 150 (pair nodes)", i_seg, v_node, v_sphere, n_sm, f_clk / 1e6),
                         "measured_inst_per_segment": prof.get("thread_inst_per_segment"),
                         "issue_slot_utilisation_ncu": prof.get("issue_slot_utilisation")},
            "roofline_hbm": {"bound": "hbm", "achieved": algo_bytes / (mean_step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": algo_bytes / (mean_step_ms * 1e-3) / 1e9 / hbm_peak, "traffic": prof.get("dram_bytes_per_launch"),
                             "note": "accum RMW + uchar4 store only (36 B/pixel/launch); BVH + spheres are staged in shared memory, so this "
                                     "path is not HBM-bound (of measured %s)" % ("peak" if peaks else "fallback")},
        }
        if not info.scene_in_smem:
            # scenes traversed from L2/HBM (C4 / C5): SURVEY 8(d) byte model, one pair visit = 2 x 32-byte nodes, one sphere = 32 bytes
            b_seg = 64.0 * v_node + 32.0 * v_sphere
            out["roofline_issue"] = out["roofline"]
            out["roofline"] = {"bound": "hbm", "achieved": per_gpu_rate * b_seg / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": per_gpu_rate * b_seg / 1e9 / hbm_peak, "traffic": None,
                               "model": "B_seg = 64*V_node + 32*V_sphere = %.0f bytes/segment (V_node=%.2f pair visits, V_sphere=%.2f measured); node fetches "
                                        "are random 64-byte pairs served by L2 (1 M spheres: 93 %% L2 hits) or HBM (16 M: 58 %%)" % (b_seg, v_node, v_sphere)}
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
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "wavefront", "pool", "slots"])
    for opt in ("pool-slots", "pool-threads", "pool-service", "pool-leaf-batch"):
        ap.add_argument("--" + opt, type=int, default=0)
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--leaf-size", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast", action="store_true", help="opt-in relaxed-numerics kernels (VN_FAST)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (vn_set_option), repeatable")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
