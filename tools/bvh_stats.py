"""Traversal statistics of the PRODUCT's BVH on the CPU (tests/hostemu build of the device code): nodes visited and
triangles tested per ray for one frame of a stand-in scene, next to the oracle's BVH2 counters on the same rays.
usage: python tools/bvh_stats.py [scene=sponza] [width=480] [height=270] [bvh_build=0]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "sponza"
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 480
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 270
    pkg = importlib.import_module("wgpu-pathtracer_b200")
    st = importlib.import_module("wgpu-pathtracer_b200.standins")
    import conftest
    import binding as oracle
    scene, cam = st.BUILDERS[scene_name]()
    build = int(sys.argv[4]) if len(sys.argv) > 4 else 0        # ptc_bvh_build: 0 SAH (host), 1 LBVH, 2 PLOC
    cfg = pkg.make_config(W, H, max_bounces=4, bvh_build=build)
    camera = st.make_camera(cam, W, H)
    emu = conftest.HostEmu(pkg).scene(scene, cfg)
    u = camera.uniforms(frame=1)
    r = emu.render_frame(u)
    o = oracle.OracleScene(scene, cfg).render_frame(u, None)
    print(f"{scene_name} {W}x{H}: bvh {emu.bvh_info()}")
    print(f"  product: rays {r['rays']}  nodes/ray {r['nodes'] / r['rays']:.3f}  tris/ray {r['tris'] / r['rays']:.3f}")
    print(f"  oracle : rays {o['rays']}  nodes/ray {o['nodes'] / o['rays']:.3f}  tris/ray {o['tris'] / o['rays']:.3f}")


if __name__ == "__main__":
    main()
