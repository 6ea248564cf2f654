"""Several GPUs behind the drop-in API: vn_multi_* (one host thread, N devices, events between the streams, one fused
peer reduce + tonemap per frame) and the one-process-per-GPU path bench.py is launched with (CUDA IPC + device-side epoch flags).
The frame must equal a single GPU's running mean over the same subframes up to float re-association (sum-then-divide vs the sequential
lerp of RayTracer.cu:208-213): accumulation buffer within 2e-6 relative, uchar4 image within one code value.

Tests that need more than one GPU skip on a single-GPU box (run them with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`);
the single-device cases exercise the same code (N = 1 peers) everywhere."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import venusaur_b200 as vb
from venusaur_b200 import VN_ACCUM_SUM, VN_ASYNC, VN_IMAGE_HOST, VN_NO_TONEMAP, sharding

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, DEPTH = 400, 225, 16, 50


def _n_gpus():
    return int(vb.load().vn_device_count())


def _single_gpu_reference(n_subframes, first=1):
    ref = vb.Context(0)
    ref.set_spheres(vb.rtiow_final_scene())
    ref.build_bvh()
    cam = vb.rtiow_camera(W, H)
    img = np.zeros((H, W, 4), np.uint8)
    for k in range(n_subframes):
        ref.render(ref.make_params(cam, W, H, SPP, first + k, DEPTH, accum_count=k, image=img.ctypes.data, flags=VN_IMAGE_HOST))
    acc, st = ref.read_accum(), ref.stats()
    ref.close()
    return acc, img, st.segments_total


def _check(acc, img, want, wimg, what):
    rel = np.abs(acc[..., :3] - want[..., :3]) / np.maximum(np.abs(want[..., :3]), 1e-6)
    d = np.abs(img.astype(np.int32) - wimg.astype(np.int32))
    print("%s: accum max rel err %.3g, image max code diff %d on %.4f %% of pixels" % (what, rel.max(), d.max(), 100.0 * (d.max(axis=-1) > 0).mean()))
    assert rel.max() < 2e-6 and d.max() <= 1


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_vn_multi_frame_equals_single_gpu_running_mean(n_dev):
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    m = vb.MultiContext(list(range(n_dev)))
    m.set_spheres(vb.rtiow_final_scene())
    m.build_bvh()
    cam = vb.rtiow_camera(W, H)
    S = 6
    want, wimg, wsegs = _single_gpu_reference(S)
    # one call, six subframes dealt to the devices, image to host memory
    img = np.zeros((H, W, 4), np.uint8)
    m.render(m.make_params(cam, W, H, SPP, 1, DEPTH, image=img.ctypes.data, flags=VN_IMAGE_HOST), S)
    assert m.subframes_accumulated() == S
    _check(m.read_accum(), img, want, wimg, "vn_multi %d devices, one call" % n_dev)
    assert m.stats().segments_total == wsegs                     # the same rays were traced, wherever they ran
    # progressive: the same six subframes in three calls (accum_count continues), asynchronous, device image on devices[0]
    lib = vb.load()
    dev, host = C.c_void_p(), C.c_void_p()
    assert lib.vn_buffer_alloc(0, W * H * 4, 0, C.byref(dev), C.byref(host)) == 0
    for call in range(3):
        m.render(m.make_params(cam, W, H, SPP, 1 + 2 * call, DEPTH, accum_count=2 * call, image=dev, flags=VN_ASYNC), 2)
    m.synchronize()
    img2 = np.zeros((H, W, 4), np.uint8)
    assert lib.vn_buffer_copy_to_host(0, img2.ctypes.data_as(C.c_void_p), dev, img2.nbytes) == 0
    _check(m.read_accum(), img2, want, wimg, "vn_multi %d devices, three progressive calls" % n_dev)
    lib.vn_buffer_free(0, dev, host, 0)
    # a wrong accum_count is refused (it would silently mix two accumulations)
    with pytest.raises(vb.Exception):
        m.render(m.make_params(cam, W, H, SPP, 7, DEPTH, accum_count=3, flags=VN_NO_TONEMAP), 1)
    # accum_count = 0 starts over
    m.render(m.make_params(cam, W, H, SPP, 1, DEPTH, image=img.ctypes.data, flags=VN_IMAGE_HOST), 1)
    w1, wi1, _ = _single_gpu_reference(1)
    _check(m.read_accum(), img, w1, wi1, "vn_multi %d devices, restart" % n_dev)
    m.close()


@pytest.mark.parametrize("n_dev", [1, 2])
def test_renderer_set_devices_python_and_cpp(n_dev, oracle_mod):
    """Renderer::SetDevices: the reference's three calls (Init / Draw / Cleanup), several devices underneath.  Two Draws of n_dev
    subframes each = the single-GPU image after 2 n_dev Draws."""
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    r = vb.Renderer()
    r.SetDevices(list(range(n_dev)))
    r.m_maxDepth = DEPTH
    r.Init(vb.Scene())
    cam = vb.rtiow_camera(W, H)
    buf = vb.CUDAOutputBuffer(vb.CUDAOutputBuffer.CUDA_DEVICE, W, H, 0)
    for _ in range(2):
        r.Draw(cam, buf)
    img = buf.getHostPointer()
    want, wimg, _ = _single_gpu_reference(2 * n_dev)
    assert np.abs(img.astype(np.int32) - wimg.astype(np.int32)).max() <= 1
    r.Cleanup()
    # the C++ header-only shim
    src = r'''
    #include <cstdio>
    #include "venusaur/Renderer.h"
    int main(int argc, char** argv) {
        const int n_dev = atoi(argv[2]);
        std::vector<int> devices;
        for (int i = 0; i < n_dev; i++) devices.push_back(i);
        Scene scene;
        Camera camera(venusaur::vec3(13, 2, 3), 20.0f, 400.0f / 225.0f, 0.1f, 10.0f);
        camera.SetForward(venusaur::vec3(0 - 13, 0 - 2, 0 - 3));
        Renderer renderer;
        renderer.SetDevices(devices);
        renderer.SetMaxDepth(50);
        renderer.Init(scene, "");
        CUDAOutputBuffer<uchar4> buf(CUDAOutputBufferType::CUDA_DEVICE, 400, 225, 0);
        for (int i = 0; i < 2; ++i) renderer.Draw(camera, buf);
        uchar4* px = buf.getHostPointer();
        FILE* f = fopen(argv[1], "wb");
        fwrite(px, 4, 400 * 225, f);
        fclose(f);
        printf("%u\n", renderer.SubframeIndex());
        renderer.Cleanup();
        return 0;
    }'''
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    cpp, exe, raw = os.path.join(out, "dropin_multi.cpp"), os.path.join(out, "dropin_multi"), os.path.join(out, "dropin_multi.raw")
    open(cpp, "w").write(src)
    libdir = os.path.dirname(vb.lib_path())
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), cpp, "-o", exe, "-L" + libdir, "-lvenusaur_b200", "-Wl,-rpath," + libdir], check=True)
    res = subprocess.run([exe, raw, str(n_dev)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert int(res.stdout.strip()) == 2 * n_dev
    cimg = np.fromfile(raw, np.uint8).reshape(H, W, 4)
    assert np.abs(cimg.astype(np.int32) - wimg.astype(np.int32)).max() <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import venusaur_b200 as vb
from venusaur_b200 import VN_ACCUM_SUM, VN_ASYNC, VN_NO_TONEMAP, sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
W, H, spp, depth, K = 400, 225, 16, 50, 3
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
ctx = vb.Context(rank)
ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
cam = vb.rtiow_camera(W, H)
ctx.resize(W, H)
accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
image = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
ctx.set_accum_external(accum.data_ptr())
flags = ctx.sync_flags()
mine = torch.tensor(list(ctx.ipc_export(accum.data_ptr())) + list(ctx.ipc_export(image.data_ptr())) + list(ctx.ipc_export(flags)), dtype=torch.uint8, device=dev)
allh = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allh, mine)
ptrs, fl, img0 = [], [], None
B = ctx.IPC_BYTES
for r in range(world):
    hb = bytes(allh[r].cpu().tolist())
    ptrs.append(accum.data_ptr() if r == rank else ctx.ipc_open(hb[:B]))
    fl.append(flags if r == rank else ctx.ipc_open(hb[2 * B:3 * B]))
    if r == 0:
        img0 = image.data_ptr() if rank == 0 else ctx.ipc_open(hb[B:2 * B])
ok = True
for frame in range(2):                                   # two frames: the epoch flags are reused with a growing epoch
    ctx.reset_accum()
    for sub in sharding.subframes_for_rank(rank, world, K):
        ctx.render(ctx.make_params(cam, W, H, spp, sub + 100 * frame, depth, flags=VN_ACCUM_SUM | VN_NO_TONEMAP | VN_ASYNC))
    e = frame + 1
    # NO host barrier: device-side epoch flags order the GPUs
    ctx.signal(0, e)
    ctx.reduce_tonemap_peers_wait(ptrs, 1.0 / (K * world), sharding.row_slice(rank, world, H), accum.data_ptr(), img0, fl, e, VN_ASYNC)
    ctx.signal(1, e)
    ctx.wait_flags([f + 4 for f in fl], e)
    ctx.synchronize()
    ctx.check_flags()
    full = [torch.zeros_like(accum) for _ in range(world)]
    dist.all_gather(full, accum)
    if rank == 0:
        got = torch.zeros_like(accum)
        for r in range(world):
            a, b = sharding.row_slice(r, world, H)
            got[a:b] = full[r][a:b]
        got = (got[..., :3] / float(K * world)).cpu().numpy()
        ref = vb.Context(0)
        ref.set_spheres(vb.rtiow_final_scene()); ref.build_bvh(); ref.resize(W, H)
        ref_img = np.zeros((H, W, 4), np.uint8)
        for k in range(K * world):
            ref.render(ref.make_params(cam, W, H, spp, k + 1 + 100 * frame, depth, accum_count=k, image=ref_img.ctypes.data, flags=vb.VN_IMAGE_HOST))
        want = ref.read_accum()[..., :3]
        ref.close()
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
        d = np.abs(image.cpu().numpy().astype(np.int32) - ref_img.astype(np.int32))
        print("ipc peer reduce N=%d frame %d: accum max rel err %.3g, image max code diff %d" % (world, frame, rel.max(), d.max()), flush=True)
        ok = ok and bool(rel.max() < 2e-6) and bool(d.max() <= 1)
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
ctx.close()
sys.exit(0 if flag.item() == 1 else 1)
'''


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.timeout(600)
def test_one_process_per_gpu_peer_reduce_with_epoch_flags(world, tmp_path):
    """The path bench.py takes under torchrun: every rank renders its subframes into its own partial sums, maps the peers' buffers and
    epoch flags through CUDA IPC and runs the fused reduce + tonemap without any host barrier."""
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=500)[0] for p in procs]
    print(outs[0])
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
