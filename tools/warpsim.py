"""CPU model of k_trace's warp scheduling (tests/hostemu/warpsim.cpp): compare vote policies and ray orders by the
warp-instruction estimate  outer*C_OUTER + node*C_NODE + tri*C_TRI + refill*C_REFILL + inner_votes*C_INNER  (constants
counted in the SASS of k_trace).  usage: python tools/warpsim.py [scene] [W] [H] [segments...]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ptc = importlib.import_module("wgpu-pathtracer_b200")
st = importlib.import_module("wgpu-pathtracer_b200.standins")
from conftest import HostEmu

C_OUTER, C_NODE, C_TRI, C_REFILL, C_INNER = 22, 40, 62, 150, 5


class SimParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("tri_vote", "refill_vote", "policy", "nwarps", "node_steps", "node_inner", "tri_inner", "_pad")]


class SimOut(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("it_refill", "it_tri", "it_node", "ln_refill", "ln_tri", "ln_node", "node_visits", "tri_tests", "rays",
                                          "mismatches", "it_outer", "it_inner_vote")]


def tile_order(W, H, tw=8, th=4):
    ys, xs = np.mgrid[0:H, 0:W]
    key = ((ys // th) * ((W + tw - 1) // tw) + xs // tw) * (tw * th) + (ys % th) * tw + xs % tw
    return np.argsort(key.ravel(), kind="stable")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "sponza"
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 480
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 272
    segs = [int(a) for a in sys.argv[4:]] or [0, 2]
    emu = HostEmu(ptc)
    L = emu.L
    L.emu_warpsim.restype = C.c_int
    L.emu_warpsim.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SimParams), C.POINTER(SimOut), C.c_void_p]
    L.emu_collect_rays.restype = C.c_uint64
    L.emu_collect_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
    scene, cam = st.BUILDERS[name]() if name != "sponza" else st.sponza(tex_size=64)
    cfg = ptc.make_config(W, H)
    es = emu.scene(scene, cfg)
    un = st.make_camera(cam, W, H).uniforms(frame=1)
    policies = [("rv%d" % rv, dict(policy=0, tri_vote=8, refill_vote=rv, node_inner=12, tri_inner=8)) for rv in (4, 8, 12, 16, 20)]
    for seg in segs:
        buf = np.zeros((W * H, 8), np.float32)
        n = int(L.emu_collect_rays(es._h, C.byref(un), seg, buf.ctypes.data_as(C.c_void_p), W * H))
        rays = buf[:n]
        # pixel ids are not returned; for segment 0 the ray order is the pixel order, so the tile permutation applies directly
        orders = [("scanline", rays)]
        if seg == 0:
            orders.append(("tile8x4", np.ascontiguousarray(rays[tile_order(W, H)])))
        else:
            octant = (rays[:, 4] < 0).astype(np.int64) * 4 + (rays[:, 5] < 0) * 2 + (rays[:, 6] < 0)
            orders.append(("binned by direction octant", np.ascontiguousarray(rays[np.argsort(octant, kind="stable")])))
            # a spatial sort of the origins (Morton order on a 2^7 grid), then octant
            lo, hi = rays[:, :3].min(0), rays[:, :3].max(0)
            q = np.clip(((rays[:, :3] - lo) / np.maximum(hi - lo, 1e-9) * 128).astype(np.int64), 0, 127)
            m = np.zeros(len(rays), np.int64)
            for b in range(7):
                for a in range(3):
                    m |= ((q[:, a] >> b) & 1) << (3 * b + a)
            orders.append(("Morton(origin) sort", np.ascontiguousarray(rays[np.argsort(m, kind="stable")])))
            orders.append(("octant, then Morton(origin)", np.ascontiguousarray(rays[np.lexsort((m, octant))])))
        for oname, r in orders:
            print(f"== {name} {W}x{H} segment {seg}: {n} rays, order {oname}")
            base = None
            for pname, kw in policies:
                P = SimParams(nwarps=64, **kw)
                O = SimOut()
                L.emu_warpsim(es._h, r.ctypes.data_as(C.c_void_p), n, C.byref(P), C.byref(O), None)
                cost = O.it_outer * C_OUTER + O.it_node * C_NODE + O.it_tri * C_TRI + O.it_refill * C_REFILL + O.it_inner_vote * C_INNER
                cost60 = cost - O.it_refill * (C_REFILL - 60)
                base = base or cost
                print(f"  {pname:22s} cost/ray {cost / n:7.1f} ({cost / base:5.3f})  outer {O.it_outer / n * 32:6.1f} node it {O.it_node / n * 32:6.1f} util {O.ln_node / max(1, O.it_node) / 32:4.2f}"
                      f" | refill=60: {cost60 / n:6.1f} |  tri it {O.it_tri / n * 32:5.1f} util {O.ln_tri / max(1, O.it_tri) / 32:4.2f}  refill {O.it_refill / n * 32:4.1f}  nodes/ray {O.node_visits / n:5.1f} tris/ray {O.tri_tests / n:4.2f} mism {O.mismatches}")


if __name__ == "__main__":
    main()
