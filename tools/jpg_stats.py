"""Differences between the library's JPEG decode (nvJPEG, GPU) and OpenCV's libjpeg on the vendored Sponza textures: the numbers
behind the tolerance of tests/test_image_decode.py.  usage (GPU box): python tools/jpg_stats.py"""
import sys, os, importlib, numpy as np, cv2
sys.path.insert(0, "."); sys.path.insert(0, "tests")
ptc = importlib.import_module("wgpu-pathtracer_b200")
import test_image_decode as t
d = os.path.join(t.RES, "sponza")
for f in t.sponza_jpgs():
    data = open(os.path.join(d, f), "rb").read()
    rc, got = t.lib_decode(ptc, data)
    ref = cv2.imread(os.path.join(d, f), cv2.IMREAD_UNCHANGED)
    ref = np.repeat(ref[:, :, None], 3, axis=2) if ref.ndim == 2 else ref[:, :, ::-1]
    diff = np.abs(got[..., :3].astype(np.int32) - ref.astype(np.int32))
    sof, ncomp, samp = t.jpeg_sof(data)
    print(f"{f:22s} {'progressive' if sof == 0xC2 else 'baseline':11s} {got.shape[1]}x{got.shape[0]} components {ncomp} sampling {[hex(x) for x in samp]} max {diff.max()} mean {diff.mean():.3f} p99.9 {np.percentile(diff, 99.9):.1f}")
