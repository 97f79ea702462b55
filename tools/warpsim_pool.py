"""Model of a CTA-level triangle pool against the production scheduler of k_trace (tests/hostemu/warpsim.cpp: emu_warpsim_pool).
usage: python tools/warpsim_pool.py <scene> <W> <H> <segment>   ->  profiles/r02x_warpsim_pool.txt"""
import sys, os, ctypes as C, importlib, numpy as np
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/tests"); sys.path.insert(0, ROOT+"/tools")
import warpsim as ws
ptc=ws.ptc; st=ws.st
from conftest import HostEmu
name=sys.argv[1]; W=int(sys.argv[2]); H=int(sys.argv[3]); seg=int(sys.argv[4])
emu=HostEmu(ptc); L=emu.L
class PoolParams(C.Structure): _fields_=[(n,C.c_uint32) for n in ("warps_per_cta","ctas","cap","refill_vote","node_min","p0","p1","p2")]
class PoolOut(C.Structure): _fields_=[(n,C.c_uint64) for n in ("it_refill","it_tri","it_node","ln_tri","ln_node","node_visits","tri_tests","rays","mismatches","it_decide","blocked","_pad")]
L.emu_warpsim.restype=C.c_int
L.emu_warpsim.argtypes=[C.c_void_p,C.c_void_p,C.c_uint64,C.POINTER(ws.SimParams),C.POINTER(ws.SimOut),C.c_void_p]
L.emu_warpsim_pool.restype=C.c_int
L.emu_warpsim_pool.argtypes=[C.c_void_p,C.c_void_p,C.c_uint64,C.POINTER(PoolParams),C.POINTER(PoolOut)]
L.emu_collect_rays.restype=C.c_uint64
L.emu_collect_rays.argtypes=[C.c_void_p,C.c_void_p,C.c_uint32,C.c_void_p,C.c_uint64]
scene,cam = st.BUILDERS[name]() if name!="sponza" else st.sponza(tex_size=64)
cfg=ptc.make_config(W,H); es=emu.scene(scene,cfg)
un=st.make_camera(cam,W,H).uniforms(frame=1)
buf=np.zeros((W*H,8),np.float32)
n=int(L.emu_collect_rays(es._h,C.byref(un),seg,buf.ctypes.data_as(C.c_void_p),W*H))
rays=buf[:n]
P=ws.SimParams(nwarps=64,policy=0,tri_vote=8,refill_vote=16,node_inner=12,tri_inner=8); O=ws.SimOut()
L.emu_warpsim(es._h,rays.ctypes.data_as(C.c_void_p),n,C.byref(P),C.byref(O),None)
cost=O.it_outer*22+O.it_node*54+O.it_tri*75+O.it_refill*150+O.it_inner_vote*5
print(f"current  cost/ray {cost/n:6.1f}  node it {O.it_node/n*32:5.1f} util {O.ln_node/max(1,O.it_node)/32:4.2f}  tri it {O.it_tri/n*32:5.1f} util {O.ln_tri/max(1,O.it_tri)/32:4.2f} refill {O.it_refill/n*32:4.1f} nodes/ray {O.node_visits/n:5.1f} tris {O.tri_tests/n:4.2f}")
for G in (2,4,8):
  for cap in (4,8,16):
    for nm in (8,16):
      PP=PoolParams(warps_per_cta=G,ctas=64//G,cap=cap,refill_vote=16,node_min=nm); PO=PoolOut()
      L.emu_warpsim_pool(es._h,rays.ctypes.data_as(C.c_void_p),n,C.byref(PP),C.byref(PO))
      c2=PO.it_decide*12+PO.it_node*(54+10)+PO.it_tri*(75+25)+PO.it_refill*150
      print(f"pool G{G} cap{cap:2d} nmin{nm:2d} cost/ray {c2/n:6.1f} ({c2/cost:5.3f}) node it {PO.it_node/n*32:5.1f} util {PO.ln_node/max(1,PO.it_node)/32:4.2f} tri it {PO.it_tri/n*32:5.1f} util {PO.ln_tri/max(1,PO.it_tri)/32:4.2f} refill {PO.it_refill/n*32:4.1f} decide {PO.it_decide/n*32:5.1f} nodes/ray {PO.node_visits/n:5.1f} tris {PO.tri_tests/n:4.2f} mism {PO.mismatches}")
