"""Dumps per-ray BVH step sequences of the RTIOW workload (host build of the product traversal code) for the SIMT simulators\n(tools/simt_sim_phases.cpp, tools/simt_sim_slots.cpp, tools/simt_model.py).  Writes /tmp/simt/seq_<variant>.bin."""
import sys, os, ctypes as C, numpy as np, subprocess
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_lib as ol
os.makedirs("/tmp/simt", exist_ok=True)
subprocess.run(["g++","-O2","-std=c++17","-ffp-contract=off","-fPIC","-shared","-w","-I/root/repo/venusaur_b200/csrc","-o","/tmp/simt/libhh.so","/root/repo/tests/host_harness.cpp"],check=True)
hh=C.CDLL('/tmp/simt/libhh.so')
hh.hh_step_sequences.restype=C.c_uint64
hh.hh_step_sequences.argtypes=[C.c_void_p,C.c_uint32,C.c_uint32,C.c_float,C.c_void_p,C.c_void_p,C.c_uint64]
hh.hh_set_sah_max.argtypes=[C.c_uint32]
hh.hh_set_wide.argtypes=[C.c_int]
class HP(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in "width height spp subframe max_depth".split()] + \
               [("origin", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3), ("lens", C.c_float)]
W,H,spp=480,270,4
cam=ol.rtiow_camera(W,H)
p=HP(); p.width,p.height,p.spp,p.subframe,p.max_depth=W,H,spp,1,50
p.origin,p.u,p.v,p.w,p.lens=(C.c_float*3)(*cam[0]),(C.c_float*3)(*cam[1]),(C.c_float*3)(*cam[2]),(C.c_float*3)(*cam[3]),float(cam[4])
rt=np.ascontiguousarray(ol.rtiow_final_scene())
for name,sah,leaf,wide in [("karras2",0,2,0),("wkarras2",0,2,1),("wsah1",4096,1,1),("wsah2",4096,2,1),("wsah3",4096,3,1),("wsah4",4096,4,1)]:
    hh.hh_set_sah_max(sah); hh.hh_set_wide(wide)
    n=hh.hh_step_sequences(rt.ctypes.data,len(rt),leaf,0.01,C.byref(p),None,0)
    buf=np.zeros(n,np.uint8)
    hh.hh_step_sequences(rt.ctypes.data,len(rt),leaf,0.01,C.byref(p),buf.ctypes.data,n)
    buf.tofile('/tmp/simt/seq_%s.bin'%name)
    nseg=(buf==255).sum(); print(name,'tokens',n,'segments',nseg,'paths',(buf==253).sum(),'node/seg %.2f'%((buf==0).sum()/nseg),'leafvisits/seg %.2f'%(((buf>=1)&(buf<=8)).sum()/nseg),'spheres/seg %.2f'%(buf[(buf>=1)&(buf<=8)].sum()/nseg))
