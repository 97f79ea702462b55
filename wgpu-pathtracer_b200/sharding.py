"""Frame (sample) sharding across GPUs — the multi-GPU story of the path.

Every (pixel, frame) sample depends only on (x, y, frame) through the RNG seed (random.wgsl:24-28) and the scene is
read-only, so ranks render disjoint frame sets with no data-path collective; per-rank images are *sums* (PTC_ACCUM_SUM:
rgb += c, a += 1), combined once with a sum-reduce.  The union of the frame sets equals the single-GPU run, so the
result differs from it only by fp32 summation order.
"""
import numpy as np


def frames_for_rank(rank, world, steps, first_frame=1):
    """Frames rank `rank` of `world` renders in a job of `steps` frames per rank: first+rank, first+rank+world, ..."""
    return [first_frame + rank + i * world for i in range(steps)]


def all_frames(world, steps, first_frame=1):
    return sorted(f for r in range(world) for f in frames_for_rank(r, world, steps, first_frame))


def reduce_sum(image, dist=None, dst=0):
    """Sum-reduce a (H, W, 4) torch tensor of per-rank sums onto `dst` (NCCL on GPUs, gloo in the CPU tests)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(image, dst=dst)
    return image


def mean_from_sum(sum_image):
    """rgb / a — the running mean the reference keeps in its accumulation texture (pathtracer.wgsl:37)."""
    s = np.asarray(sum_image)
    out = s.copy()
    a = np.maximum(s[..., 3:4], 1.0)
    out[..., :3] = s[..., :3] / a
    return out
