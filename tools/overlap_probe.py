"""Feasibility probe: aggregate Mrays/s of P concurrent steps at 1/P of the frame height versus one full-frame step
(do persistent kernels of independent streams fill each other's tails?).  usage: python tools/overlap_probe.py [P=2]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ptc = importlib.import_module("wgpu-pathtracer_b200"); st = importlib.import_module("wgpu-pathtracer_b200.standins")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "sponza"
scene, cam = st.BUILDERS[name]()
W, H, K = 1920, 1080, 64


def run(parts):
    steps = [ptc.RTStep.create(scene, ptc.make_config(W, H // parts)) for _ in range(parts)]
    cams = [st.make_camera(cam, W, H // parts) for _ in range(parts)]
    for f in range(1, 5):
        for s, c in zip(steps, cams):
            s.update(c.uniforms(frame=f)); s.render()
    for s in steps: s.poll(True)
    r0 = sum(s.stats()["rays_traced"] for s in steps)
    torch.cuda.synchronize(); t = time.perf_counter()
    for f in range(5, 5 + K):
        for s, c in zip(steps, cams):
            s.update(c.uniforms(frame=f)); s.render()
    for s in steps: s.poll(True)
    dt = time.perf_counter() - t
    rays = sum(s.stats()["rays_traced"] for s in steps) - r0
    for s in steps: s.destroy()
    return rays / dt / 1e6, dt / K * 1e3


for parts in (1, P, 1, P):
    v, ms = run(parts)
    print(f"{name}: {parts} concurrent step(s) of {W}x{H // parts}: {v:.1f} Mrays/s, {ms:.3f} ms per full frame")
