#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on the BASELINE.json workload.

Workload (N=1): configs[2] of BASELINE.json, the scene the metric is quoted on — "sponza 1920x1080 ... Disney BRDF",
as the procedural stand-in of wgpu-pathtracer_b200/standins.py (the reference's .obj files are Git-LFS pointers
and /root/reference does not exist on the GPU box).  One step = one RTStep::render = one frame = 1 sample for every
pixel = raygen + 5 x (closest-hit traversal + shade/accumulate) kernels, MAX_BOUNCES = 4 as the reference.

value    Mrays/s (trace_ray invocations per second, whole job), inputs resident in HBM, device-timed with CUDA events on
         the step's stream, max over ranks.
e2e      the same through the reference-facing call sequence with HOST buffers: every step uploads its 112-byte Uniforms
         (ptc_update), renders (ptc_render), tonemaps to the 8-bit display image (ptc_blit, the reference's BlitStep) and
         reads that image back into pinned host memory (ptc_read_async, overlapped on a copy stream; drained before the
         clock stops).
roofline the closest-hit traversal kernel: algorithmic bytes/ray (64 B x BVH2 inner nodes + 48 B x triangles tested, both
         counted by the oracle on the identical rays, + 64 B ray/hit record) x rays / its CUDA-event time, against the
         measured HBM copy peak in MEASURED_PEAKS.json.
cpu_baseline / --impl reference   the oracle (CPU restatement of the WGSL; kind "port") on the box's host cores; the
         reference's own wgpu path cannot run here (no Rust toolchain, no Vulkan ICD).
Multi-GPU: frames are sharded over ranks (rank r renders frames r+1, r+1+N, ...), scene replicated, per-rank SUM
         accumulation, one NCCL reduce of the image inside the timed region.  scaling = "weak".
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "sponza_1080p": dict(scene="sponza", width=1920, height=1080, max_bounces=4, config_index=2),
    "cornell_512": dict(scene="cornell", width=512, height=512, max_bounces=8, config_index=0),
    "dragon_1080p": dict(scene="dragon", width=1920, height=1080, max_bounces=4, config_index=1),
    "classroom_1080p": dict(scene="classroom", width=1920, height=1080, max_bounces=4, config_index=3),
    "san_miguel_4k": dict(scene="san_miguel", width=3840, height=2160, max_bounces=4, config_index=4),
    # non-reference quality mode: shadow rays with any-hit early-out + MIS; rays include the shadow rays
    "classroom_1080p_mis": dict(scene="classroom", width=1920, height=1080, max_bounces=4, config_index=3, light_sampling=1),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)          # nvidia-smi must be gone before anything else is timed
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # timed region shorter than one sample: fall back to everything sampled
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name):
    pkg = importlib.import_module("wgpu-pathtracer_b200")
    st = importlib.import_module("wgpu-pathtracer_b200.standins")
    w = WORKLOADS[name]
    scene, cam = st.BUILDERS[w["scene"]]()
    camera = st.make_camera(cam, w["width"], w["height"])
    return pkg, st, w, scene, camera


def cpu_baseline(pkg, w, scene, camera, seconds, accum_mode, nthreads=0, first_frame=1):
    """Times the oracle (CPU restatement of the reference's WGSL path) on whole frames of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as oracle
    cfg = pkg.make_config(w["width"], w["height"], max_bounces=w["max_bounces"], accum_mode=accum_mode, light_sampling=w.get("light_sampling", 0))
    orc = oracle.OracleScene(scene, cfg)
    cores = nthreads if nthreads > 0 else (os.cpu_count() or 1)
    color = np.zeros((w["height"], w["width"], 4), np.float32)
    rays = nodes = tris = frames = 0
    t0 = time.perf_counter()
    while True:
        r = orc.render_frame(camera.uniforms(frame=first_frame + frames), color, nthreads=cores, want_aux=False)
        rays += r["rays"]; nodes += r["nodes"]; tris += r["tris"]; frames += 1
        if time.perf_counter() - t0 >= seconds or frames >= 64:
            break
    dt = time.perf_counter() - t0
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{frames} full frame(s) of {w['scene']} {w['width']}x{w['height']} (frames {first_frame}..{first_frame + frames - 1}), {rays} rays, {dt:.1f} s",
            "spp_per_s": frames / dt, "nodes_per_ray": nodes / max(rays, 1), "tris_per_ray": tris / max(rays, 1),
            "frames": frames, "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port on all host threads (see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg, st, w, scene, camera = build_workload(args.workload)
    T = pkg.types
    # each step = one full frame on the CPU; bounded so that the whole run ends within a few minutes
    steps, warm = min(args.steps, 8), min(args.warmup, 1)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as oracle
    cfg = pkg.make_config(w["width"], w["height"], max_bounces=w["max_bounces"], light_sampling=w.get("light_sampling", 0))
    orc = oracle.OracleScene(scene, cfg)
    cores = os.cpu_count() or 1
    color = np.zeros((w["height"], w["width"], 4), np.float32)
    for f in range(warm):
        orc.render_frame(camera.uniforms(frame=1 + f), color, nthreads=cores, want_aux=False)
    rays = 0
    t0 = time.perf_counter()
    for f in range(steps):
        rays += orc.render_frame(camera.uniforms(frame=1 + warm + f), color, nthreads=cores, want_aux=False)["rays"]
    dt = time.perf_counter() - t0
    v = rays / dt / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "spp_per_s": steps / dt,
            "config": {"workload": f"{w['scene']} stand-in {w['width']}x{w['height']}, MAX_BOUNCES={w['max_bounces']}, 1 spp per step",
                       "baseline_config_index": w["config_index"]},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} full frame(s) per run on {cores} host threads; CPU restatement of the reference's WGSL path "
                                       "(reference wgpu/Vulkan binary cannot be built or run here)"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sponza_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bvh-build", default="sah", choices=["sah", "lbvh"], help="sah: binned SAH on the host (default); lbvh: linear BVH built by CUDA kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly ONE line, the JSON: whatever libraries print on fd 1 while the job runs (NCCL's version banner
    # at NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE) is sent to stderr, and fd 1 is restored for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj, allow_nan=False), flush=True)
        os.dup2(2, 1)                                      # and anything printed while tearing down goes to stderr again

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path-tracing core has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines go to stderr: stdout carries the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg, st, w, scene, camera = build_workload(args.workload)
    T = pkg.types
    W, H, K, Wm = w["width"], w["height"], args.steps, max(args.warmup, 3)
    accum = T.ACCUM_SUM if world > 1 else T.ACCUM_MEAN
    cfg = pkg.make_config(W, H, max_bounces=w["max_bounces"], accum_mode=accum, light_sampling=w.get("light_sampling", 0),
                          bvh_build=T.BVH_LBVH_DEVICE if args.bvh_build == "lbvh" else T.BVH_SAH_HOST)
    t_create = time.perf_counter()
    step = pkg.RTStep.create(scene, cfg, device=local)
    t_create = time.perf_counter() - t_create
    out = torch.zeros(H, W, 4, device=f"cuda:{local}", dtype=torch.float32)
    step.set_output(out.data_ptr())                       # the caller injects the accumulation target (raytracer.rs:86)
    ext = torch.cuda.ExternalStream(step.stream, device=local)
    pinned = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()   # the displayed image: tonemapped sRGB RGBA8 (BlitStep's output)

    def frame_of(i):                                      # frame sharding: the union over ranks is frames 1..N*K (+warm-up)
        return 1 + rank + i * world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # The image reduce goes through the library's own entry point, ptc_reduce = one ncclReduce on the step's
    # stream, with the communicator torch.distributed already holds; torch's dist.reduce is the fallback if that
    # communicator cannot be reached from Python.
    comm, reduce_how = None, None
    if dist is not None:
        try:
            warm = torch.zeros(1, device=f"cuda:{local}"); dist.all_reduce(warm); torch.cuda.synchronize()   # the communicator exists after the first collective
            comm = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda", local))._comm_ptr()
            reduce_how = "ptc_reduce (ncclReduce of the SUM-mode colour planes on the step's stream)"
        except Exception as e:                               # pragma: no cover
            comm, reduce_how = None, f"torch.distributed.reduce (ProcessGroupNCCL._comm_ptr unavailable: {type(e).__name__})"

    def reduce_image():
        if comm is not None:
            step.reduce(comm, 0)
        else:
            with torch.cuda.stream(ext):
                dist.reduce(out, dst=0)

    # ---- device-resident timing ---------------------------------------------------------------------------------
    uniforms = [camera.uniforms(frame=frame_of(i)) for i in range(Wm + K)]
    for i in range(Wm):
        step.update(uniforms[i]); step.render()
    step.poll(True)
    step.reset_accumulation(); step.poll(True)
    rays0 = step.stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(ext)
    for i in range(K):
        step.update(uniforms[Wm + i]); step.render()
    if dist is not None:
        reduce_image()
    e1.record(ext)
    e1.synchronize()
    barrier()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    rays1 = step.stats()
    rays = rays1["rays_traced"] - rays0["rays_traced"]
    launches = rays1["kernel_launches"] - rays0["kernel_launches"]
    stat = torch.tensor([ms, float(rays), float(launches)], device=f"cuda:{local}", dtype=torch.float64)
    if dist is not None:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, rays, launches = float(mx[0]), float(sm[1]), float(sm[2])
    value = rays / (ms * 1e-3) / 1e6

    # ---- per-stage device times: a separate, shorter pass with stage timing on.  Profiling keeps every launch on one
    # stream (no primary-ray prefetch), so each kernel's CUDA-event duration is exclusive — that is what the roofline uses.
    P = min(K, 32)
    step.set_profiling(True)
    r0 = step.stats()["rays_traced"]
    for i in range(P):
        step.update(uniforms[Wm + i]); step.render()
    step.poll(True)
    prof = step.profile()
    prof_rays = step.stats()["rays_traced"] - r0
    step.set_profiling(False)

    # ---- end to end through the reference-facing calls with host buffers ---------------------------------------
    # K steps, repeated E2E_REPS times; the reported value is the MEDIAN repetition (all are listed in e2e.runs):
    # a single wall-clock pass is exposed to host-side noise (other tenants' PCIe traffic, nvidia-smi holding driver locks).
    E2E_REPS = 3
    runs = []
    for rep in range(E2E_REPS):
        step.reset_accumulation(); step.poll(True)
        rays0 = step.stats()["rays_traced"]
        barrier()
        t_e0 = time.perf_counter()
        for i in range(K):
            u = camera.uniforms(frame=frame_of(Wm + K * (rep + 1) + i))   # host-side Uniforms built per step, 112 B H2D as kernel arguments
            step.update(u); step.render()
            step.blit()                                        # BlitStep: tonemap + sRGB encode of the accumulated image (1 kernel)
            step.read_async(T.PLANE_DISPLAY, pinned.data_ptr())  # D2H of the frame a user sees, overlapped on the copy stream
        if dist is not None:
            reduce_image()
        step.poll(True)
        step.read_wait()
        barrier()
        e2e_s = time.perf_counter() - t_e0                      # includes the drain of the last host copy
        e2e_rays = float(step.stats()["rays_traced"] - rays0)
        es = torch.tensor([e2e_s, e2e_rays], device=f"cuda:{local}", dtype=torch.float64)
        if dist is not None:
            mx = es.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = es.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            e2e_s, e2e_rays = float(mx[0]), float(sm[1])
        runs.append(e2e_rays / e2e_s / 1e6)
    e2e_value = float(np.median(runs))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline + algorithmic bytes per ray (oracle counters on the identical rays) ------------------------
    # N = 1: the oracle is timed on a bounded sample (cpu_baseline).  N > 1: one untimed oracle frame only supplies the
    # BVH2 node/triangle counters the roofline figure is defined on.
    base = None
    if not args.no_cpu_baseline:
        base = cpu_baseline(pkg, w, scene, camera, args.cpu_seconds if world == 1 else 0.0, T.ACCUM_MEAN, first_frame=frame_of(Wm))
    nodes_per_ray = base["nodes_per_ray"] if base else None
    tris_per_ray = base["tris_per_ray"] if base else None
    bytes_per_ray = 64.0 * nodes_per_ray + 48.0 * tris_per_ray + 64.0 if base else None
    peak, peak_src = load_peaks()
    trace_ms, trace_n = prof["trace"]["ms"], prof["trace"]["launches"]
    rank_rays = prof_rays                                   # this rank's rays in the profiled pass
    stage_total = sum(v["ms"] for v in prof.values())
    achieved = bytes_per_ray * rank_rays / (trace_ms * 1e-3) / 1e9 if (base and trace_ms > 0) else None
    bvh = step.bvh_info()
    traffic = l2_traffic = None                             # dram / L2 bytes per k_trace launch from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "trace_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(args.workload, {})
        traffic, l2_traffic = tj.get("dram_bytes_per_launch"), tj.get("l2_bytes_per_launch")
    # second denominator: the BVH of every configuration but San Miguel is L2-resident, so the traversal's memory
    # ceiling is the L2, measured live with a streaming read of a 32 MiB resident buffer
    import ctypes
    l2_peak = ctypes.c_double(0.0)
    if pkg.lib().ptc_debug_l2_bandwidth(local, 32 << 20, 400, ctypes.byref(l2_peak)) != 0:
        l2_peak = None
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "spp_per_s": K * world / (ms * 1e-3),
        "config": {"workload": f"{w['scene']} stand-in {W}x{H}, MAX_BOUNCES={w['max_bounces']}, 1 spp per step, {bvh['triangles']} triangles"
                               + (", PTC_LIGHTS_MIS (non-reference mode: one any-hit shadow ray per opaque hit, counted as a ray)" if w.get("light_sampling") else ""),
                   "bvh": {"build": "lbvh (CUDA kernels)" if bvh["built_on_device"] else "binned SAH (host threads)", "build_ms": bvh["build_us"] / 1e3,
                           "nodes": bvh["nodes"], "max_depth": bvh["max_depth"], "ptc_create_s": t_create},
                   "baseline_config_index": w["config_index"], "parallelism": f"frame-sharded x{world}, scene replicated; image reduce: {reduce_how}" if world > 1 else "single GPU",
                   "frame_parts": step.stats()["frame_parts"],
                   "l2": "inputs larger than L2: each frame streams ~%.0f MB of path-state queues; the %.1f MB BVH is L2-resident by design" % (
                       W * H * 144 / 1e6, (bvh["node_bytes"] + bvh["triangle_bytes"]) / 1e6)},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 112, "d2h_bytes_per_step": W * H * 4,
                "note": "per step: ptc_update(host Uniforms, 112 B) + ptc_render + ptc_blit (tonemap to the 8-bit display image, as the reference's "
                        "BlitStep) + ptc_read_async(DISPLAY -> pinned host, overlapped on a copy stream); wall clock incl. final copy drain"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_trace (closest-hit BVH traversal)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved is not None else None, "traffic": traffic, "peak_source": peak_src,
                     "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                     "rays_per_launch": rank_rays / max(trace_n, 1), "avg_launch_ms": trace_ms / max(trace_n, 1), "launches": int(trace_n),
                     "l2": {"peak_gbs": l2_peak.value if l2_peak is not None else None,
                            "peak_how": "measured live: ptc_debug_l2_bandwidth, 32 MiB resident buffer read 400 times with ld.global.cg.v4",
                            "frac_of_l2_peak": (achieved / l2_peak.value) if (achieved is not None and l2_peak is not None and l2_peak.value > 0) else None,
                            "traffic": l2_traffic,
                            "achieved_gbs": (l2_traffic / (trace_ms / max(trace_n, 1) * 1e-3) / 1e9) if (l2_traffic and trace_n) else None,
                            "note": "frac_of_l2_peak = the same algorithmic GB/s against the L2 ceiling; traffic = lts__t_sectors.sum x 32 B per whole-frame k_trace launch (ncu --set full, profiles/), achieved_gbs = that over the live launch duration"},
                     "kernel_share_of_step": trace_ms / stage_total if stage_total > 0 else None,
                     "stage_ms": {k: v["ms"] for k, v in prof.items()}, "stage_pass_steps": P,
                     "note": "stage times from a separate pass with stage timing on: the frame runs as ONE wavefront on one stream (no frame parts, no primary-ray prefetch), "
                             "so each launch covers a whole segment and its CUDA-event duration is exclusive; the timed region itself runs %d concurrent parts" % step.stats()["frame_parts"]},
        "cpu_baseline": ({k: base[k] for k in ("value", "unit", "cores", "kind", "sample")} if (base and world == 1) else None),
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
