"""Diagnose e2e variance: per-step host-side call times of the bench's e2e loop, repeated."""
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
for rep in range(6):
    ts = []
    t0 = time.perf_counter()
    for i in range(64):
        a = time.perf_counter()
        u = camera.uniforms(frame=1 + i); step.update(u); b = time.perf_counter()
        step.render(); c = time.perf_counter()
        step.blit(); step.read_async(T.PLANE_DISPLAY, pinned.data_ptr()); d = time.perf_counter()
        ts.append((b - a, c - b, d - c))
    step.poll(True); step.read_wait()
    dt = time.perf_counter() - t0
    ts = np.array(ts) * 1e3
    print(f"rep {rep}: {dt/64*1e3:.3f} ms/step; host ms mean update {ts[:,0].mean():.3f} render {ts[:,1].mean():.3f} blit+read {ts[:,2].mean():.3f}; max {ts.max(axis=0)}")
