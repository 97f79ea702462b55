"""Renders a few frames of a bench workload and prints the rays traced — the process tools/trace_counters.sh runs under ncu
(every k_trace* launch of the process is captured, so instructions / rays needs no launch bookkeeping)."""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sponza_1080p")
ap.add_argument("--frames", type=int, default=3)
a = ap.parse_args()
pkg, st, w, scene, camera = bench.build_workload(a.workload)
cfg = pkg.make_config(w["width"], w["height"], max_bounces=w["max_bounces"], light_sampling=w.get("light_sampling", 0))
step = pkg.RTStep.create(scene, cfg)
for f in range(1, a.frames + 1):
    step.update(camera.uniforms(frame=f)); step.render()
step.poll(True)
s = step.stats()
print(json.dumps({"workload": a.workload, "frames": a.frames, "rays": s["rays_traced"], "frame_parts": s["frame_parts"]}))
