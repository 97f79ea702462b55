"""Where the e2e loop's time goes: ms per step of the bench's e2e sequence with its pieces added one at a time
(render only; + blit; + asynchronous read of the DISPLAY plane), and the host time spent inside each call."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
ptc = importlib.import_module("wgpu-pathtracer_b200"); st = importlib.import_module("wgpu-pathtracer_b200.standins")
T = ptc.types
scene, cam = st.sponza(); W, H = 1920, 1080
camera = st.make_camera(cam, W, H)
step = ptc.RTStep.create(scene, ptc.make_config(W, H))
pinned = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()
step.update(camera.uniforms(frame=1)); step.render(); step.blit(); step.read_async(T.PLANE_DISPLAY, pinned.data_ptr()); step.poll(True); step.read_wait()
K = 256
for mode in ("render", "render+blit", "render+blit+read", "render", "render+blit+read"):
    res = []
    for rep in range(3):
        step.poll(True); step.read_wait()
        host = np.zeros(3)
        t0 = time.perf_counter()
        for i in range(K):
            a = time.perf_counter()
            step.update(camera.uniforms(frame=1 + i)); step.render(); b = time.perf_counter()
            if "blit" in mode: step.blit()
            c = time.perf_counter()
            if "read" in mode: step.read_async(T.PLANE_DISPLAY, pinned.data_ptr())
            d = time.perf_counter()
            host += (b - a, c - b, d - c)
        step.poll(True); step.read_wait()
        res.append(((time.perf_counter() - t0) / K * 1e3, host / K * 1e3))
    ms = np.median([r[0] for r in res])
    h = res[0][1]
    print(f"{mode:18s} {ms:.4f} ms/step   host ms per call: update+render {h[0]:.4f} blit {h[1]:.4f} read_async {h[2]:.4f}")
