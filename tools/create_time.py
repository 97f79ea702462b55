"""create_time.py [workload] — host + device time of ptc_create for one stand-in scene (scene preparation, BVH build, uploads),
per builder: what a caller waits for before the first frame.  Prints one JSON line."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ptc = importlib.import_module("wgpu-pathtracer_b200")
standins = importlib.import_module("wgpu-pathtracer_b200.standins")

name = sys.argv[1] if len(sys.argv) > 1 else "san_miguel"
t = time.time()
scene, cam = standins.BUILDERS[name]()
out = {"workload": name, "triangles": len(scene.indices) // 3, "standin_s": round(time.time() - t, 2), "host_threads": os.cpu_count(), "builders": {}}
for label, kind in (("sah_host", ptc.types.BVH_SAH_HOST), ("lbvh_device", ptc.types.BVH_LBVH_DEVICE), ("ploc_device", ptc.types.BVH_PLOC_DEVICE)):
    cfg = ptc.make_config(256, 256, bvh_build=kind)
    t = time.time()
    step = ptc.RTStep.create(scene, cfg, device=0)
    dt = time.time() - t
    info = step.bvh_info()
    out["builders"][label] = {"ptc_create_s": round(dt, 3), "build_ms": info["build_ms"], "nodes": info["nodes"], "max_depth": info["max_depth"]}
    step.destroy()
print(json.dumps(out))
