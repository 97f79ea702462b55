"""Host time per ptc_update + ptc_render with an empty GPU queue (is the frame loop launch-bound?): ~0.10 ms per 1.53 ms frame on B200.
usage (GPU box): python tools/host_cost.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
ptc = importlib.import_module("wgpu-pathtracer_b200"); st = importlib.import_module("wgpu-pathtracer_b200.standins")
scene, cam = st.sponza(); W, H = 1920, 1080
camera = st.make_camera(cam, W, H)
step = ptc.RTStep.create(scene, ptc.make_config(W, H))
for f in range(1, 9): step.update(camera.uniforms(frame=f)); step.render()
step.poll(True)
for rep in range(3):
    ts = []
    for i in range(12):
        u = camera.uniforms(frame=10 + i)
        a = time.perf_counter(); step.update(u); step.render(); ts.append((time.perf_counter() - a) * 1e3)
    step.poll(True)
    print("host ms per update+render, GPU queue empty at start:", " ".join("%.3f" % t for t in ts))
