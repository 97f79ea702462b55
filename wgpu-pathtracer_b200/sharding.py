"""Frame (sample) sharding across GPUs — the multi-GPU story of the path.

Every (pixel, frame) sample depends only on (x, y, frame) through the RNG seed (random.wgsl:24-28) and the scene is
read-only, so ranks render disjoint frame sets with no data-path collective; per-rank images are *sums* (PTC_ACCUM_SUM:
rgb += c, a += 1), combined with a sum-reduce INTO A SEPARATE RECEIVE IMAGE (ptc_reduce -> PTC_PLANE_REDUCED): the
per-rank accumulation is never overwritten, so the reduce can be repeated after more frames.  The union of the frame
sets equals the single-GPU run, so the result differs from it only by fp32 summation order.

Host-side logic only (no compute): bench.py and tests/test_multirank_cpu.py both go through these functions, the first
with libptcore steps + ptc_reduce on GPUs, the second with oracle frames + gloo on CPU.
"""
import numpy as np


def frames_for_rank(rank, world, steps, first_frame=1):
    """Frames rank `rank` of `world` renders in a job of `steps` frames per rank: first+rank, first+rank+world, ..."""
    return [first_frame + rank + i * world for i in range(steps)]


def all_frames(world, steps, first_frame=1):
    return sorted(f for r in range(world) for f in frames_for_rank(r, world, steps, first_frame))


def reduce_into(recv, image, dist=None, dst=0):
    """The semantics of ptc_reduce for torch tensors: recv (on `dst`) = sum over ranks of `image`; `image` — the rank's
    own accumulation — is left untouched.  (torch's dist.reduce works in place, hence the copy.)  NCCL on GPUs, gloo in
    the CPU tests."""
    recv.copy_(image)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(recv, dst=dst)
    return recv


def mean_from_sum(sum_image):
    """rgb / a — the running mean the reference keeps in its accumulation texture (pathtracer.wgsl:37)."""
    s = np.asarray(sum_image)
    out = s.copy()
    a = np.maximum(s[..., 3:4], 1.0)
    out[..., :3] = s[..., :3] / a
    return out


def band_rows(height, width, frames, budget_pixel_frames=3.0e6, max_rows=16):
    """Rows [y0, y1) around the image centre that a CPU checker can re-render for `frames` frames within a budget."""
    rows = int(max(2, min(max_rows, budget_pixel_frames // max(1, width * frames))))
    y0 = max(0, height // 2 - rows // 2)
    return y0, min(height, y0 + rows)


def reduce_check(reduced, expected_samples, band=None, band_reference_mean=None):
    """Checks a reduced SUM image: every pixel's sample count equals `expected_samples` (= ranks x frames per rank), and —
    when a reference is given — the mean radiance of rows band=[y0, y1) agrees with `band_reference_mean` (those rows of
    the single-process render of the SAME frame set).  Returns the dict bench.py prints as "reduce_check"."""
    red = np.asarray(reduced)
    out = {"alpha_ok": bool((red[..., 3] == np.float32(expected_samples)).all()), "expected_samples": int(expected_samples)}
    if band is not None and band_reference_mean is not None:
        y0, y1 = band
        got = mean_from_sum(red[y0:y1])[..., :3].astype(np.float64)
        ref = np.asarray(band_reference_mean)[..., :3].astype(np.float64)
        ok = np.isfinite(got).all(axis=-1) & np.isfinite(ref).all(axis=-1)
        num = float(np.sqrt(np.mean((got[ok] - ref[ok]) ** 2))) if ok.any() else float("nan")
        den = float(np.sqrt(np.mean(ref[ok] ** 2))) if ok.any() else float("nan")
        out.update({"band_rows": [int(y0), int(y1)], "band_rel_rmse": num / max(den, 1e-30),
                    "band_non_finite_mismatch": int((np.isfinite(got).all(axis=-1) != np.isfinite(ref).all(axis=-1)).sum())})
        out["band_ok"] = bool(out["band_rel_rmse"] <= 1e-5 and out["band_non_finite_mismatch"] == 0)
    return out
