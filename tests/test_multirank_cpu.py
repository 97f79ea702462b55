"""world_size-2 test of the frame-sharding + sum-reduce logic on CPU (gloo).  Each rank renders its frames with the
oracle in SUM mode (the CUDA step cannot run here); the reduced image must equal the single-process run."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, steps, extra, W, H, out_path):
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
    mine, recv = torch.from_numpy(img), torch.zeros(H, W, 4)
    sh.reduce_into(recv, mine, dist, dst=0)                 # the semantics of ptc_reduce: COLOR -> REDUCED, COLOR untouched
    first = recv.numpy().copy()
    # a second report interval: more frames on every rank, then ANOTHER reduce of the whole accumulation — an in-place
    # reduce would count the first interval of the other ranks twice here
    for f in sh.frames_for_rank(rank, world, extra, first_frame=1 + world * steps):
        orc.render_frame(camera.uniforms(frame=f), img, nthreads=1, want_aux=False)
    sh.reduce_into(recv, mine, dist, dst=0)
    dist.barrier()
    if rank == 0:
        np.savez(out_path, first=first, second=recv.numpy())
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
    world, steps, extra, W, H = 2, 3, 2, 40, 32
    out = str(tmp_path / "sum.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, steps, extra, W, H, out), nprocs=world, join=True)
    z = np.load(out)
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    scene, cam = standins.cornell()
    camera = standins.make_camera(cam, W, H)
    orc = oracle.OracleScene(scene, ptc.make_config(W, H))
    ref = np.zeros((H, W, 4), np.float32)
    band = sh.band_rows(H, W, world * (steps + extra))
    for n_frames, key in ((world * steps, "first"), (world * (steps + extra), "second")):
        first_new = 1 if key == "first" else 1 + world * steps
        for f in range(first_new, n_frames + 1):             # the single-process render of the same frame set, continued
            orc.render_frame(camera.uniforms(frame=f), ref, want_aux=False)
        chk = sh.reduce_check(z[key], n_frames, band, ref[band[0]:band[1]])      # the function bench.py reports with
        assert chk["alpha_ok"] and chk["band_ok"], chk
        np.testing.assert_allclose(sh.mean_from_sum(z[key])[..., :3], ref[..., :3], rtol=1e-5, atol=1e-6)
    assert not sh.reduce_check(z["second"], world * steps)["alpha_ok"]


def test_band_rows_budget():
    sh = importlib.import_module("wgpu-pathtracer_b200.sharding")
    assert sh.band_rows(1080, 1920, 20) == (532, 548)          # 16 rows around the centre
    y0, y1 = sh.band_rows(2160, 3840, 2048)
    assert y1 - y0 == 2 and 0 <= y0 < y1 <= 2160               # never fewer than two rows
    assert sh.band_rows(8, 16, 1) == (0, 8)
