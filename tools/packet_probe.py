"""Packet-traversal coherence per workload: node visits per packet, fraction of lanes entering a child, and the CUDA-event
time of the primary-ray traversal with and without packets (stage timing on, one part).  usage: python tools/packet_probe.py [workloads...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ptc = importlib.import_module("wgpu-pathtracer_b200"); st = importlib.import_module("wgpu-pathtracer_b200.standins")
sys.argv = sys.argv[:1] + (sys.argv[1:] or ["sponza", "classroom", "dragon", "san_miguel", "cornell"])
SIZES = {"san_miguel": (3840, 2160), "cornell": (512, 512)}
for name in sys.argv[1:]:
    scene, cam = st.BUILDERS[name]()
    W, H = SIZES.get(name, (1920, 1080))
    camera = st.make_camera(cam, W, H)
    res = {}
    for packets in ("1", "0"):
        os.environ["PTC_PRIMARY_PACKETS"] = packets
        step = ptc.RTStep.create(scene, ptc.make_config(W, H, max_bounces=0))     # primary segment only
        for f in (1, 2, 3):
            step.update(camera.uniforms(frame=f)); step.render()
        step.poll(True)
        s0 = step.stats()
        step.set_profiling(True)
        for f in range(4, 12):
            step.update(camera.uniforms(frame=f)); step.render()
        step.poll(True)
        prof = step.profile(); s1 = step.stats()
        res[packets] = (prof["trace"]["ms"] / 8, s1["packet_node_visits"] - s0["packet_node_visits"], s1["packet_lane_visits"] - s0["packet_lane_visits"])
        step.destroy()
    ms1, nv, lv = res["1"]; ms0 = res["0"][0]
    npk = 8 * ((W * H + 31) // 32)
    print(f"{name:11s} {W}x{H}: primary trace {ms1:.3f} ms with packets, {ms0:.3f} ms without ({ms0 / ms1:.2f}x); "
          f"node visits per packet {nv / npk:.1f}, lanes entering a child {lv / max(nv, 1):.1f} of 32")
