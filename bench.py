#!/usr/bin/env python
"""bench.py — Mrays/s of the path-tracing hot path on the BASELINE.json workloads.

Headline workload: configs[2] of BASELINE.json, the scene the metric is quoted on — "sponza 1920x1080 ... Disney BRDF" —
as the procedural stand-in geometry of wgpu-pathtracer_b200/standins.py (the reference's .obj files are Git-LFS pointers
and /root/reference does not exist on the GPU box) with the reference's REAL sponza.mtl and textures (tests/golden/ref_res).
One step = one RTStep::render = one frame = 1 sample for every pixel = raygen + 5 x (closest-hit traversal +
shade/accumulate) kernels, MAX_BOUNCES = 4 as the reference.

value    Mrays/s (trace_ray invocations per second, whole job), inputs resident in HBM, device-timed with CUDA events on
         the step's stream, max over ranks; the K-step region is repeated REPS (>= 5) times, the MEDIAN repetition is
         reported and all are listed in "value_runs".
e2e      the same through the reference-facing call sequence with HOST buffers: every step uploads its 112-byte Uniforms
         (ptc_update), renders (ptc_render), tonemaps to the 8-bit display image (ptc_blit, the reference's BlitStep) and
         reads that image back into pinned host memory (ptc_read_async, overlapped on a copy stream; drained before the
         clock stops).  Median of E2E_REPS repetitions.
roofline the closest-hit traversal kernel.  Algorithmic bytes/ray = 64 B x BVH2 inner nodes + 48 B x triangles tested (both
         counted by the oracle on the identical rays) + 64 B ray/hit record; achieved = that x rays / the kernel's
         CUDA-event time.  "bound" is "l2" when nodes + triangles fit the L2 (every configuration but San Miguel), else
         "hbm"; the L2 peak is measured live (ptc_debug_l2_bandwidth) and printed with the SM clock it ran at, the HBM
         peak is MEASURED_PEAKS.json's.  "issue" is the roofline that actually binds this kernel: warp instructions per
         ray (from the committed ncu capture of this code, profiles/trace_counters.json — a hardware counter that cannot
         be read from inside the run, labelled as such) x live rays/s against 148 SMs x 4 schedulers x the live SM clock.
reduce_check (N > 1) rank 0 checks the image ptc_reduce received: sample count == ranks x steps everywhere, and a band of
         rows equals the oracle's mean over the same frame set; then one more frame per rank + a SECOND reduce must give
         ranks x (steps + 1) (the reduce does not double-count).  Outside the timed region.
other_workloads  BASELINE configs[3] (classroom, also with the shadow-ray/any-hit mode) and configs[4] (San Miguel
         3840x2160, with the NCCL reduce at N > 1), measured after the headline with fewer steps.
cpu_baseline / --impl reference   the oracle (CPU restatement of the WGSL; kind "port") on the box's host cores; the
         reference's own wgpu path cannot run here (no Rust toolchain, no Vulkan ICD).  The reference arm runs
         exactly --steps timed steps after --warmup untimed ones; a step is one frame, or a centred band of rows of one
         frame when full frames would not fit its time budget (PTC_REFERENCE_BUDGET_S, default 120 s).
Multi-GPU: frames are sharded over ranks (rank r renders frames r+1, r+1+N, ...), scene replicated, per-rank SUM
         accumulation, one NCCL reduce of the image inside the timed region.  scaling = "weak".
"""
import argparse
import ctypes
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
# measured after the headline workload, each with min(K, steps) steps per repetition: name -> (steps, repetitions)
OTHER_WORKLOADS = {"classroom_1080p": (32, 3), "classroom_1080p_mis": (32, 3), "san_miguel_4k": (16, 3)}
REPS, E2E_REPS = 5, 3
NUM_SMS, SCHEDULERS_PER_SM = 148, 4


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

    def window(self, t0, t1):
        """median SM clock over the samples taken in [t0, t1] (None when there is none)"""
        sm = []
        for ts, line in list(self.rows):
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(line.split(",")[0]))
                except Exception:
                    pass
        return float(np.median(sm)) if sm else None

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


def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as oracle
    return oracle


def cpu_baseline(pkg, w, scene, camera, seconds, nthreads=0, first_frame=1, rows=None, orc=None):
    """Times the oracle (CPU restatement of the reference's WGSL path) on frames (or a band of rows) of the same workload;
    also the source of the BVH2 node / triangle counters the algorithmic bytes per ray are defined on."""
    oracle = oracle_module()
    cfg = pkg.make_config(w["width"], w["height"], max_bounces=w["max_bounces"], light_sampling=w.get("light_sampling", 0))
    orc = orc or oracle.OracleScene(scene, cfg)
    cores = nthreads if nthreads > 0 else (os.cpu_count() or 1)
    color = np.zeros((w["height"], w["width"], 4), np.float32)
    rays = nodes = tris = frames = 0
    t0 = time.perf_counter()
    while True:
        r = orc.render_frame(camera.uniforms(frame=first_frame + frames), color, nthreads=cores, want_aux=False, rows=rows)
        rays += r["rays"]; nodes += r["nodes"]; tris += r["tris"]; frames += 1
        if time.perf_counter() - t0 >= seconds or frames >= 64:
            break
    dt = time.perf_counter() - t0
    what = "full frame(s)" if rows is None else f"band(s) of rows {rows[0]}..{rows[1] - 1}"
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{frames} {what} of {w['scene']} {w['width']}x{w['height']} (frames {first_frame}..{first_frame + frames - 1}), {rays} rays, {dt:.1f} s",
            "spp_per_s": frames / dt, "nodes_per_ray": nodes / max(rays, 1), "tris_per_ray": tris / max(rays, 1),
            "frames": frames, "seconds": dt, "_orc": orc}


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port on all host threads (see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg, st, w, scene, camera = build_workload(args.workload)
    # EXACTLY --steps timed steps after --warmup untimed ones, like the other arm.  A step is one frame on the CPU when the
    # whole run then fits the time budget (about two minutes), otherwise a centred band of rows of each frame, sized from
    # one untimed probe frame so that it does: a bounded sample of the same workload, said so in `sample`.
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget_s = float(os.environ.get("PTC_REFERENCE_BUDGET_S", "120"))
    oracle = oracle_module()
    cfg = pkg.make_config(w["width"], w["height"], max_bounces=w["max_bounces"], light_sampling=w.get("light_sampling", 0))
    orc = oracle.OracleScene(scene, cfg)
    cores = os.cpu_count() or 1
    H = w["height"]
    color = np.zeros((H, w["width"], 4), np.float32)
    t0 = time.perf_counter()
    orc.render_frame(camera.uniforms(frame=1), color, nthreads=cores, want_aux=False)          # probe (untimed; the first warm-up)
    t_frame = time.perf_counter() - t0
    rows = None
    if (steps + warm) * t_frame > budget_s:
        n = min(H, max(16, int(H * budget_s / ((steps + warm) * t_frame))))
        rows = ((H - n) // 2, (H - n) // 2 + n)
    for f in range(1, warm):
        orc.render_frame(camera.uniforms(frame=1 + f), color, nthreads=cores, want_aux=False, rows=rows)
    rays = 0
    t0 = time.perf_counter()
    for f in range(steps):
        rays += orc.render_frame(camera.uniforms(frame=1 + max(warm, 1) + f), color, nthreads=cores, want_aux=False, rows=rows)["rays"]
    dt = time.perf_counter() - t0
    per_step = "1 full frame" if rows is None else f"rows {rows[0]}..{rows[1] - 1} of 1 frame ({rows[1] - rows[0]} of {H} rows, sized from a {t_frame:.2f} s probe frame for a {budget_s:.0f} s run)"
    v = rays / dt / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "spp_per_s": steps / dt * (1.0 if rows is None else (rows[1] - rows[0]) / H),
            "config": {"workload": workload_string(w, scene.num_triangles), "baseline_config_index": w["config_index"]},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} steps of {per_step} on {cores} host threads; CPU restatement of the reference's WGSL path "
                                       "(reference wgpu/Vulkan binary cannot be built or run here)"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_string(w, triangles):
    return (f"{w['scene']} stand-in {w['width']}x{w['height']}, MAX_BOUNCES={w['max_bounces']}, 1 spp per step, {triangles} triangles"
            + (", PTC_LIGHTS_MIS (non-reference mode: one any-hit shadow ray per opaque hit, counted as a ray)" if w.get("light_sampling") else ""))


class Job:
    """process-wide state of one bench run: rank, torch.distributed, the NCCL communicator ptc_reduce is handed"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the path-tracing core has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = f"cuda:{self.local}"
        self.dist, self.comm, self.reduce_how = None, None, None
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines go to stderr: stdout carries the one JSON line
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
            # The image reduce goes through the library's own entry point, ptc_reduce = one ncclReduce on the step's
            # stream, with the communicator torch.distributed already holds; torch's dist.reduce is the fallback if that
            # communicator cannot be reached from Python.
            try:
                warm = torch.zeros(1, device=self.dev); dist.all_reduce(warm); torch.cuda.synchronize()   # the communicator exists after the first collective
                self.comm = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda", self.local))._comm_ptr()
                self.reduce_how = "ptc_reduce (ncclReduce of the SUM-mode colour planes into the REDUCED plane, on the step's stream)"
            except Exception as e:                               # pragma: no cover
                self.comm, self.reduce_how = None, f"torch.distributed.reduce (ProcessGroupNCCL._comm_ptr unavailable: {type(e).__name__})"

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_sum(self, values):
        """(max over ranks, sum over ranks) of a list of floats"""
        t = self.torch.tensor(values, device=self.dev, dtype=self.torch.float64)
        if self.dist is None:
            return t.tolist(), t.tolist()
        mx = t.clone(); self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
        sm = t.clone(); self.dist.all_reduce(sm, op=self.dist.ReduceOp.SUM)
        return mx.tolist(), sm.tolist()


def measure(job, name, K, Wm, reps, e2e_reps, bvh_build, sampler=None, want_cpu_seconds=0.0, counters_only=False):
    """Measures one workload on this rank's GPU (all ranks call it together).  Returns (result dict, extras) on every rank;
    only rank 0's is complete."""
    torch = job.torch
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    pkg, st, w, scene, camera = build_workload(name)
    T = pkg.types
    W, H, world, rank = w["width"], w["height"], job.world, job.rank
    accum = T.ACCUM_SUM if world > 1 else T.ACCUM_MEAN
    cfg = pkg.make_config(W, H, max_bounces=w["max_bounces"], accum_mode=accum, light_sampling=w.get("light_sampling", 0),
                          bvh_build={"sah": T.BVH_SAH_HOST, "lbvh": T.BVH_LBVH_DEVICE, "ploc": T.BVH_PLOC_DEVICE}[bvh_build])
    t_create = time.perf_counter()
    step = pkg.RTStep.create(scene, cfg, device=job.local)
    t_create = time.perf_counter() - t_create
    out = torch.zeros(H, W, 4, device=job.dev, dtype=torch.float32)
    step.set_output(out.data_ptr())                       # the caller injects the accumulation target (raytracer.rs:86)
    reduced = None
    if world > 1:
        reduced = torch.zeros(H, W, 4, device=job.dev, dtype=torch.float32)
        step.set_reduce_output(reduced.data_ptr())        # ... and the receive buffer of the reduce
    ext = torch.cuda.ExternalStream(step.stream, device=job.local)
    pinned = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()   # the displayed image: tonemapped sRGB RGBA8 (BlitStep's output)

    frame_at = lambda i: sh.frames_for_rank(rank, world, i + 1)[-1]   # i-th frame this rank renders (frame sharding, sharding.py)

    def reduce_image():
        if job.comm is not None:
            step.reduce(job.comm, 0)
        else:
            with torch.cuda.stream(ext):
                sh.reduce_into(reduced, out, job.dist, dst=0)

    # ---- device-resident timing: REPS repetitions of the K-step region --------------------------------------------
    for i in range(Wm):
        step.update(camera.uniforms(frame=frame_at(i))); step.render()
    step.poll(True)
    runs, run_ms, run_clock, launches = [], [], [], 0
    t_first = time.time()
    for rep in range(reps):
        uniforms = [camera.uniforms(frame=frame_at(Wm + rep * K + i)) for i in range(K)]
        step.reset_accumulation(); step.poll(True)
        s0 = step.stats()
        job.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.time()
        e0.record(ext)
        for u in uniforms:
            step.update(u); step.render()
        if world > 1:
            reduce_image()
        e1.record(ext)
        e1.synchronize()
        job.barrier()
        tw1 = time.time()
        s1 = step.stats()
        (ms, _, _), (_, rays, ln) = job.max_sum([e0.elapsed_time(e1), float(s1["rays_traced"] - s0["rays_traced"]), float(s1["kernel_launches"] - s0["kernel_launches"])])
        runs.append(rays / (ms * 1e-3) / 1e6); run_ms.append(ms); launches = int(ln)
        run_clock.append(sampler.window(tw0, tw1) if sampler is not None else None)
    t_last = time.time()
    order = np.argsort(runs)
    med = int(order[len(runs) // 2])
    value, ms = runs[med], run_ms[med]
    last_first_frame = 1 + world * (Wm + (reps - 1) * K)      # the frame set of the last repetition: last_first_frame .. +world*K-1

    # ---- multi-GPU: what did the reduce deliver?  (untimed) -----------------------------------------------------------
    check = None
    if world > 1:
        torch.cuda.synchronize()
        red_host = reduced.cpu().numpy() if rank == 0 else None
        # a second reduce after one more frame per rank: the receive plane must hold ranks x (K + 1) samples — the per-rank
        # accumulation was not overwritten by the first reduce, so nothing is counted twice
        step.update(camera.uniforms(frame=frame_at(Wm + reps * K))); step.render()
        reduce_image(); step.poll(True); torch.cuda.synchronize()
        job.barrier()
        if rank == 0:
            band = sh.band_rows(H, W, world * K)
            oracle = oracle_module()
            ocfg = pkg.make_config(W, H, max_bounces=w["max_bounces"], light_sampling=w.get("light_sampling", 0))
            orc = oracle.OracleScene(scene, ocfg)
            ref = np.zeros((H, W, 4), np.float32)
            t0 = time.perf_counter()
            for f in range(last_first_frame, last_first_frame + world * K):
                orc.render_frame(camera.uniforms(frame=f), ref, want_aux=False, rows=band)
            check = sh.reduce_check(red_host, world * K, band, ref[band[0]:band[1]])
            check["band_frames"] = [last_first_frame, last_first_frame + world * K - 1]
            check["oracle_seconds"] = time.perf_counter() - t0
            second = reduced.cpu().numpy()
            check["second_reduce_alpha_ok"] = bool((second[..., 3] == np.float32(world * (K + 1))).all())
            check["ok"] = bool(check["alpha_ok"] and check.get("band_ok", False) and check["second_reduce_alpha_ok"])
            orc.close()
        job.barrier()

    # ---- per-stage device times: a separate, shorter pass with stage timing on.  Profiling keeps every launch on one
    # stream (no primary-ray prefetch), so each kernel's CUDA-event duration is exclusive — that is what the roofline uses.
    P = min(K, 32)
    step.set_profiling(True)
    r0 = step.stats()["rays_traced"]
    tp0 = time.time()
    for i in range(P):
        step.update(camera.uniforms(frame=frame_at(Wm + i))); step.render()
    step.poll(True)
    tp1 = time.time()
    prof = step.profile()
    prof_rays = step.stats()["rays_traced"] - r0
    step.set_profiling(False)

    # ---- end to end through the reference-facing calls with host buffers ---------------------------------------
    # K steps, repeated; the reported value is the MEDIAN repetition (all are listed in e2e.runs): a single wall-clock
    # pass is exposed to host-side noise (other tenants' PCIe traffic, nvidia-smi holding driver locks).
    e2e_runs = []
    step.blit(); step.read_async(T.PLANE_DISPLAY, pinned.data_ptr()); step.read_wait()   # untimed: first use allocates the DISPLAY plane and the snapshot
    for rep in range(e2e_reps):
        step.reset_accumulation(); step.poll(True)
        rays0 = step.stats()["rays_traced"]
        job.barrier()
        t_e0 = time.perf_counter()
        for i in range(K):
            u = camera.uniforms(frame=frame_at(Wm + K * (rep + 1) + i))   # host-side Uniforms built per step, 112 B H2D as kernel arguments
            step.update(u); step.render()
            step.blit()                                        # BlitStep: tonemap + sRGB encode of the accumulated image
            step.read_async(T.PLANE_DISPLAY, pinned.data_ptr())  # D2H of the frame a user sees, overlapped on the copy stream
        if world > 1:
            reduce_image()
        step.poll(True)
        step.read_wait()
        job.barrier()
        e2e_s = time.perf_counter() - t_e0                      # includes the drain of the last host copy
        (e2e_s, _), (_, e2e_rays) = job.max_sum([e2e_s, float(step.stats()["rays_traced"] - rays0)])
        e2e_runs.append(e2e_rays / e2e_s / 1e6)
    e2e_value = float(np.median(e2e_runs)) if e2e_runs else None

    bvh = step.bvh_info()
    parts = step.stats()["frame_parts"]
    res = None
    if rank == 0:
        # ---- CPU baseline + algorithmic bytes per ray (oracle counters on the identical rays) ------------------------
        # headline at N = 1: the oracle is timed on a bounded sample (cpu_baseline).  Otherwise one untimed oracle band only
        # supplies the BVH2 node/triangle counters the roofline figure is defined on.
        base = None
        if not counters_only or want_cpu_seconds > 0:
            base = cpu_baseline(pkg, w, scene, camera, want_cpu_seconds, first_frame=frame_at(Wm))
        else:
            base = cpu_baseline(pkg, w, scene, camera, 0.0, first_frame=frame_at(Wm), rows=sh.band_rows(H, W, 4, max_rows=64))
        base.pop("_orc").close()
        nodes_per_ray, tris_per_ray = base["nodes_per_ray"], base["tris_per_ray"]
        bytes_per_ray = 64.0 * nodes_per_ray + 48.0 * tris_per_ray + 64.0
        hbm_peak, hbm_src = load_peaks()
        trace_ms, trace_n = prof["trace"]["ms"], prof["trace"]["launches"]
        stage_total = sum(v["ms"] for v in prof.values())
        achieved = bytes_per_ray * prof_rays / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
        # L2 ceiling, measured live with a streaming read of a 32 MiB resident buffer, with the clock it ran at
        l2_peak = ctypes.c_double(0.0)
        tl0 = time.time()
        l2_ok = pkg.lib().ptc_debug_l2_bandwidth(job.local, 32 << 20, 400, ctypes.byref(l2_peak)) == 0
        tl1 = time.time()
        l2_gbs = l2_peak.value if l2_ok and l2_peak.value > 0 else None
        l2_bytes = int(torch.cuda.get_device_properties(job.local).L2_cache_size)
        bvh_bytes = bvh["node_bytes"] + bvh["triangle_bytes"]
        bound = "l2" if bvh_bytes < l2_bytes else "hbm"
        peak = l2_gbs if (bound == "l2" and l2_gbs) else hbm_peak
        # instruction-issue roofline: hardware counters come from the committed ncu capture of this code
        sm_mhz_stage = sampler.window(tp0, tp1) if sampler is not None else None
        issue = None
        cp = os.path.join(ROOT, "profiles", "trace_counters.json")
        if os.path.exists(cp):
            cj = json.load(open(cp)).get(name)
            if cj and trace_ms > 0:
                clock = (sm_mhz_stage or cj.get("sm_mhz") or 1965.0) * 1e6
                inst_s = cj["warp_inst_per_ray"] * prof_rays / (trace_ms * 1e-3)
                issue = {"warp_inst_per_ray": cj["warp_inst_per_ray"], "simt_lanes_per_inst": cj.get("thread_inst_per_warp_inst"),
                         "achieved_gwarp_inst_s": inst_s / 1e9, "peak_gwarp_inst_s": NUM_SMS * SCHEDULERS_PER_SM * clock / 1e9,
                         "frac": inst_s / (NUM_SMS * SCHEDULERS_PER_SM * clock), "sm_mhz": clock / 1e6,
                         "counter_source": cj.get("source"), "counters_measured_in_this_run": False,
                         "note": "warp instructions per ray and lanes per instruction are ncu hardware counters of the k_trace launches of this "
                                 "code (committed capture, cold-cache and serialised); rays/s and the SM clock are this run's"}
        res = {
            "value": value, "unit": "Mrays/s", "ms_per_step": ms / K, "steps": K, "spp_per_s": K * world / (ms * 1e-3),
            "value_runs": runs, "value_how": f"median of {reps} repetitions of the {K}-step device-timed region",
            "run_sm_mhz": run_clock, "gpu_launches": launches,
            "config": {"workload": workload_string(w, bvh["triangles"]),
                       "bvh": {"build": f"{bvh_build} (CUDA kernels)" if bvh["built_on_device"] else "binned SAH (host threads)", "build_ms": bvh["build_us"] / 1e3,
                               "nodes": bvh["nodes"], "max_depth": bvh["max_depth"], "bytes": bvh_bytes, "ptc_create_s": t_create},
                       "baseline_config_index": w["config_index"],
                       "parallelism": f"frame-sharded x{world}, scene replicated; image reduce: {job.reduce_how}" if world > 1 else "single GPU",
                       "frame_parts": parts,
                       "l2": "inputs larger than L2: each frame streams ~%.0f MB of path-state queues through HBM; the %.1f MB of BVH nodes + triangles %s" % (
                           W * H * 144 / 1e6, bvh_bytes / 1e6,
                           "fit the %.0f MB L2" % (l2_bytes / 1e6) if bound == "l2" else "exceed the %.0f MB L2 (only the part a camera position touches stays resident)" % (l2_bytes / 1e6))},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 112, "d2h_bytes_per_step": W * H * 4, "runs": e2e_runs,
                    "note": "per step: ptc_update(host Uniforms, 112 B) + ptc_render + ptc_blit (tonemap to the 8-bit display image, as the reference's "
                            "BlitStep) + ptc_read_async(DISPLAY -> pinned host, overlapped on a copy stream); wall clock incl. final copy drain; median of the runs"},
            "roofline": {"bound": bound, "kernel": "k_trace (closest-hit BVH traversal)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if (achieved is not None and peak) else None, "traffic": None,
                         "peak_source": ("measured live: ptc_debug_l2_bandwidth, 32 MiB resident buffer read 400 times with ld.global.cg.v4"
                                         if (bound == "l2" and l2_gbs) else hbm_src),
                         "l2_gbs": l2_gbs, "l2_probe_sm_mhz": sampler.window(tl0, tl1) if sampler is not None else None, "l2_bytes": l2_bytes,
                         "hbm_gbs": hbm_peak, "hbm_source": hbm_src, "frac_of_hbm": achieved / hbm_peak if achieved is not None else None,
                         "frac_of_l2": achieved / l2_gbs if (achieved is not None and l2_gbs) else None,
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "counters_sample": base["sample"],
                         "rays_per_launch": prof_rays / max(trace_n, 1), "avg_launch_ms": trace_ms / max(trace_n, 1), "launches": int(trace_n),
                         "issue": issue,
                         "kernel_share_of_step": trace_ms / stage_total if stage_total > 0 else None,
                         "stage_ms": {k: v["ms"] for k, v in prof.items()}, "stage_pass_steps": P, "stage_pass_sm_mhz": sm_mhz_stage,
                         "note": "algorithmic bytes (oracle BVH2 counters) over the kernel's CUDA-event time; traffic (dram bytes per launch) is an ncu "
                                 "counter and is not measured in this run: see profiles/ for the ncu --set full capture.  Stage times come from a separate "
                                 "pass with stage timing on: the frame runs as ONE wavefront on one stream (no frame parts, no primary-ray prefetch), so each "
                                 "launch covers a whole segment and its CUDA-event duration is exclusive; the timed region itself runs %d concurrent parts" % parts},
            "reduce_check": check,
            "cpu_baseline": ({k: base[k] for k in ("value", "unit", "cores", "kind", "sample")} if want_cpu_seconds > 0 else None),
        }
    step.destroy()
    del out, reduced, pinned
    torch.cuda.empty_cache()
    return res, (t_first, t_last)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sponza_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads pass (A/B timing of the headline only)")
    ap.add_argument("--reps", type=int, default=REPS)
    ap.add_argument("--bvh-build", default="sah", choices=["sah", "lbvh", "ploc"],
                    help="sah: binned SAH on the host (default); lbvh: Morton-order linear BVH, ploc: bottom-up clustering — both built by CUDA kernels")
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

    job = Job()
    K, Wm = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(job.local)
    if job.rank == 0:
        sampler.start()
        time.sleep(0.25)
    cpu_s = 0.0 if (args.no_cpu_baseline or job.world > 1) else args.cpu_seconds
    head, (t0, t1) = measure(job, args.workload, K, Wm, max(args.reps, 1), E2E_REPS, args.bvh_build, sampler if job.rank == 0 else None,
                             want_cpu_seconds=cpu_s, counters_only=False)
    others = {}
    if not args.no_others and args.workload == "sponza_1080p":
        for name, (k2, reps2) in OTHER_WORKLOADS.items():
            try:
                r, _ = measure(job, name, min(K, k2), Wm, reps2, 1, args.bvh_build, sampler if job.rank == 0 else None, counters_only=True)
            except Exception as e:                           # a failing extra workload must not take the headline line with it
                r = {"error": f"{type(e).__name__}: {e}"}
                job.barrier()
            if job.rank == 0:
                if "config" in r:
                    r["workload"] = r["config"]["workload"]
                others[name] = r
    if job.rank != 0:
        if job.dist is not None:
            job.dist.destroy_process_group()
        return
    clocks = sampler.stop(t0, t1)
    line = {
        "metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": job.world, "steps": K, "warmup": Wm,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "spp_per_s": head["spp_per_s"], "value_runs": head["value_runs"], "value_how": head["value_how"],
        "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks,
        "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "reduce_check": head["reduce_check"],
        "other_workloads": others or None,
    }
    emit(line)
    if job.dist is not None:
        job.dist.destroy_process_group()


if __name__ == "__main__":
    main()
