"""Small end-to-end run for compute-sanitizer (memcheck / initcheck): reference mode, light sampling, device LBVH build,
blit + reprojection + reads, on frames that use several frame parts.  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ptc = importlib.import_module("wgpu-pathtracer_b200"); st = importlib.import_module("wgpu-pathtracer_b200.standins")
T = ptc.types
for name, kw in (("cornell", {}), ("lit_room", dict(light_sampling=T.LIGHTS_MIS, bvh_build=T.BVH_LBVH_DEVICE))):
    scene, cam = st.BUILDERS[name]()
    W, H = 211, 157
    step = ptc.RTStep.create(scene, ptc.make_config(W, H, max_bounces=4, **kw))
    camera = st.make_camera(cam, W, H)
    prev = None
    for f in (1, 2, 3):
        u = camera.uniforms(frame=f)
        step.update(u); step.render(); step.blit()
        if prev is not None:
            step.rayproject(prev, u)
        prev = u
    step.poll(True)
    for p in (T.PLANE_COLOR, T.PLANE_DEPTH, T.PLANE_NORMAL, T.PLANE_ALBEDO, T.PLANE_HIT_ID, T.PLANE_DISPLAY):
        step.read(p)
    rays = np.zeros((1000, 8), np.float32); rays[:, 4] = 1.0
    step.trace_rays(rays); step.trace_rays_any(rays)
    print(name, "ok", step.stats())
    step.destroy()
