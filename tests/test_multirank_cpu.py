"""world_size-2 test of the frame-sharding + sum-reduce logic on CPU (gloo).  Each rank renders its frames with the
oracle in SUM mode (the CUDA step cannot run here); the reduced image must equal the single-process run."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, steps, W, H, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ptc = importlib.import_module("wgpu-pathtracer_b200")
    st = importlib.import_module("wgpu-pathtracer_b200.standins")
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    import binding as oracle
    scene, cam = st.cornell()
    camera = st.make_camera(cam, W, H)
    orc = oracle.OracleScene(scene, ptc.make_config(W, H, accum_mode=ptc.types.ACCUM_SUM))
    img = np.zeros((H, W, 4), np.float32)
    for f in sh.frames_for_rank(rank, world, steps):
        orc.render_frame(camera.uniforms(frame=f), img, nthreads=1, want_aux=False)
    t = torch.from_numpy(img)
    sh.reduce_sum(t, dist, dst=0)
    dist.barrier()
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.destroy_process_group()


def test_frame_sharding_partition():
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    for world in (1, 2, 4, 8):
        fr = [sh.frames_for_rank(r, world, 5) for r in range(world)]
        flat = sorted(f for x in fr for f in x)
        assert flat == list(range(1, 5 * world + 1)) == sh.all_frames(world, 5)      # disjoint, union = 1..N*K
        assert all(len(set(a) & set(b)) == 0 for i, a in enumerate(fr) for b in fr[i + 1:])


def test_two_rank_gloo_reduce_equals_single_run(tmp_path, ptc, standins, oracle):
    import torch.multiprocessing as mp
    world, steps, W, H = 2, 3, 40, 32
    out = str(tmp_path / "sum.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, steps, W, H, out), nprocs=world, join=True)
    summed = np.load(out)
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    assert (summed[..., 3] == world * steps).all()
    scene, cam = standins.cornell()
    camera = standins.make_camera(cam, W, H)
    orc = oracle.OracleScene(scene, ptc.make_config(W, H))
    ref = np.zeros((H, W, 4), np.float32)
    for f in sh.all_frames(world, steps):
        orc.render_frame(camera.uniforms(frame=f), ref, want_aux=False)
    np.testing.assert_allclose(sh.mean_from_sum(summed)[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6)
