"""GPU-side diagnostics (run under gpurun): device/host primitive agreement, per-bounce divergence, per-frame stage times."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
ptc = importlib.import_module("wgpu-pathtracer_b200")
st = importlib.import_module("wgpu-pathtracer_b200.standins")
import binding as oracle
T = ptc.types


def math_check():
    rng = np.random.RandomState(0)
    n = 1 << 20
    sets = {"uniform": (rng.uniform(-8, 8, n), rng.uniform(-8, 8, n)), "unit": (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)),
            "bits": (rng.randint(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32), rng.uniform(0.1, 3, n))}
    for name, (a, b) in sets.items():
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        for fn in range(13):
            if name == "bits" and fn != 8:
                continue
            out = np.zeros(n, np.float32)
            rc = ptc.lib().ptc_debug_math(0, fn, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n)
            assert rc == 0, ptc.lib().ptc_last_error()
            ref = oracle.math_fn(fn, a, b)
            bad = ~((out == ref) | (np.isnan(out) & np.isnan(ref)))
            print(f"math {name} fn={fn}: mismatches {int(bad.sum())}", (a[bad][:3], b[bad][:3], out[bad][:3], ref[bad][:3]) if bad.any() else "")


def bounce_sweep(scene_name, W, H, frames=1, depths=(0, 1, 2, 4, 8)):
    scene, cam = st.BUILDERS[scene_name]()
    camera = st.make_camera(cam, W, H)
    for mb in depths:
        cfg = ptc.make_config(W, H, max_bounces=mb)
        step = ptc.RTStep.create(scene, cfg)
        orc = oracle.OracleScene(scene, cfg)
        ref = None
        for f in range(1, frames + 1):
            u = camera.uniforms(frame=f)
            step.update(u); step.render()
            ref = orc.render_frame(u, None if ref is None else ref["color"])
        got = step.read(T.PLANE_COLOR)
        bad = np.argwhere((got != ref["color"]).any(axis=-1) & ~(np.isnan(got).any(axis=-1) & np.isnan(ref["color"]).any(axis=-1)))
        print(f"{scene_name} {W}x{H} frames={frames} max_bounces={mb}: differing pixels {len(bad)}; aux equal:",
              [bool(np.array_equal(step.read(p), ref[k], equal_nan=True)) for p, k in ((T.PLANE_DEPTH, "depth"), (T.PLANE_NORMAL, "normal"), (T.PLANE_ALBEDO, "albedo"), (T.PLANE_HIT_ID, "hit_id"))],
              "rays", step.stats()["rays_traced"], ref["rays"] if frames == 1 else "")
        for y, x in bad[:4]:
            print("   px", x, y, "gpu", got[y, x], "oracle", ref["color"][y, x])
        step.destroy()


def frame_times(scene_name, W, H, frames=6):
    scene, cam = st.BUILDERS[scene_name]()
    camera = st.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    for f in range(1, frames + 1):
        step.set_profiling(True)
        step.update(camera.uniforms(frame=f)); step.render(); step.poll(True)
        p = step.profile()
        print(f"{scene_name} frame {f}: trace {p['trace']['ms']:.3f} ms shade {p['shade']['ms']:.3f} ms rays {step.stats()['rays_traced']}")
    step.destroy()


def back_to_back(scene_name, W, H, frames=3, use_torch=False):
    """frames queued without host synchronisation (the bench pattern), compared with the oracle afterwards"""
    scene, cam = st.BUILDERS[scene_name]()
    camera = st.make_camera(cam, W, H)
    cfg = ptc.make_config(W, H)
    step = ptc.RTStep.create(scene, cfg)
    if use_torch:
        import torch
        out = torch.zeros(H, W, 4, device="cuda:0", dtype=torch.float32)
        step.set_output(out.data_ptr())
    orc = oracle.OracleScene(scene, cfg)
    us = [camera.uniforms(frame=f) for f in range(1, frames + 1)]
    step.set_profiling(True)
    for u in us:
        step.update(u); step.render()
    step.poll(True)
    p = step.profile()
    print(f"{scene_name} back-to-back x{frames} torch={use_torch}: trace {p['trace']['ms']:.3f} ms shade {p['shade']['ms']:.3f} ms rays {step.stats()['rays_traced']}")
    got = step.read(T.PLANE_COLOR)
    ref = None
    rays = 0
    for u in us:
        ref = orc.render_frame(u, None if ref is None else ref["color"]); rays += ref["rays"]
    bad = np.argwhere((got != ref["color"]).any(axis=-1))
    print("   oracle rays", rays, "differing pixels", len(bad), "gpu nan", int(np.isnan(got).sum()), "oracle nan", int(np.isnan(ref["color"]).sum()),
          "hit ids equal", bool(np.array_equal(step.read(T.PLANE_HIT_ID), ref["hit_id"])))
    for y, x in bad[:6]:
        print("   px", x, y, "gpu", got[y, x], "oracle", ref["color"][y, x])
    step.destroy()


if __name__ == "__main__":
    which = sys.argv[1:] or ["math", "sweep", "times"]
    if "math" in which:
        math_check()
    if "sweep" in which:
        bounce_sweep("cornell", 512, 512, frames=1)
        bounce_sweep("cornell", 512, 512, frames=4, depths=(8,))
    if "sponza" in which:
        bounce_sweep("sponza", int(os.environ.get("DBG_W", 480)), int(os.environ.get("DBG_H", 270)), frames=1, depths=tuple(int(x) for x in os.environ.get("DBG_DEPTHS", "0,1,2,4").split(",")))
    if "b2b" in which:
        back_to_back("sponza", 1920, 1080, 3)
        back_to_back("sponza", 1920, 1080, 3, use_torch=True)
        back_to_back("cornell", 512, 512, 4)
    if "times" in which:
        frame_times("sponza", 1920, 1080)
