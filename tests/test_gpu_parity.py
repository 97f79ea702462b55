"""GPU parity tests: the CUDA wavefront (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): primary-hit (instance, geometry) ids bit-exact; radiance relative RMSE <= 1e-3
at matched frames (tolerance written in `assert_radiance`).  Because kernels and oracle share the fp32 primitives
of include/ptmath.h the images are in practice bit-identical, which the tests report and also assert where the
scene has no known edge case.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_RMSE_TOL = 1e-3


def assert_radiance(got, ref, tol=REL_RMSE_TOL):
    g, r = got[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
    assert np.isfinite(r).all() and np.isfinite(g).all()
    rel = np.sqrt(np.mean((g - r) ** 2)) / max(np.sqrt(np.mean(r ** 2)), 1e-30)
    assert rel <= tol, f"relative RMSE {rel:.3e} > {tol}"
    return rel


def render_both(ptc, standins, oracle, scene, cam, W, H, frames, device_cfg=None, **cfg_kw):
    cfg = ptc.make_config(W, H, **cfg_kw)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg, device=0)
    orc = oracle.OracleScene(scene, cfg)
    ref = None
    for f in range(1, frames + 1):
        u = camera.uniforms(frame=f)
        step.update(u)
        step.render()
        ref = orc.render_frame(u, None if ref is None else ref["color"])
    step.poll(True)
    return step, ref


def check_all_planes(ptc, step, ref, exact=True):
    T = ptc.types
    got = {"color": step.read(T.PLANE_COLOR), "depth": step.read(T.PLANE_DEPTH), "normal": step.read(T.PLANE_NORMAL),
           "albedo": step.read(T.PLANE_ALBEDO), "hit_id": step.read(T.PLANE_HIT_ID)}
    assert np.array_equal(got["hit_id"], ref["hit_id"]), "primary hit ids differ"
    assert_radiance(got["color"], ref["color"])
    assert np.array_equal(got["color"][..., 3], ref["color"][..., 3])
    if exact:
        for k in ("color", "depth", "normal", "albedo"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), f"{k} plane is not bit-identical"
    return got


def test_device_present(ptc):
    assert ptc.lib().ptc_device_count() >= 1


def test_device_primitives_match_host(ptc, oracle):
    """include/ptmath.h evaluated on the device (nvcc -fmad=false) equals the host build (g++ -ffp-contract=off) bit for bit."""
    import ctypes as C
    rng = np.random.RandomState(0)
    n = 1 << 18
    sets = [(rng.uniform(-8, 8, n), rng.uniform(-8, 8, n), range(13)), (rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), range(13)),
            (rng.randint(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32), rng.uniform(0.1, 3, n), [8])]
    for a, b, fns in sets:
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        for fn in fns:
            out = np.zeros(n, np.float32)
            assert ptc.lib().ptc_debug_math(0, fn, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n) == 0
            ref = oracle.math_fn(fn, a, b)
            assert np.array_equal(out, ref, equal_nan=True), f"primitive {fn} differs between device and host"


def test_unorm8_texture_conversion(ptc):
    """8-bit texels must reach the kernels as rn(v / 255) (texture.rs:241 to_rgba32f).  The library reads them through the
    texture unit's normalised-float mode only if that conversion is exact for all 256 values (self-test at ptc_create);
    report which path this box uses and check the claim."""
    import ctypes as C
    out = np.zeros(512, np.float32)
    rc = ptc.lib().ptc_debug_unorm8(0, out.ctypes.data_as(C.c_void_p))
    assert rc in (0, 1)
    want = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    exact = np.array_equal(out[:256], want) and np.array_equal(out[256:], want[::-1])
    assert bool(rc) == exact
    print("unorm8 via texture unit exact:", exact)


@pytest.mark.parametrize("which", ["soup", "sponza"])
def test_texture_and_sky_lookup_bit_exact(ptc, standins, oracle, which):
    """Every texture of the scene (8-bit and float, 1x1 constants included) sampled on the device at random uv —
    wrapping, negative and texel-edge coordinates — equals the oracle's sample_texture_rgba bit for bit; same for sky_color."""
    scene, _ = standins.random_soup(600, seed=2) if which == "soup" else standins.sponza()
    cfg = ptc.make_config(8, 8)
    step = ptc.RTStep.create(scene, cfg)
    orc = oracle.OracleScene(scene, cfg)
    rng = np.random.RandomState(4)
    uv = (rng.rand(20000, 2) * 6 - 3).astype(np.float32)
    uv[:64] = (rng.randint(-8, 9, (64, 2)) / 4.0).astype(np.float32)          # exact texel-grid fractions
    ntex = len(scene.textures)
    for idx in range(ntex):
        got, ref = step.sample_texture(idx, uv), orc.sample_texture(idx, uv)
        assert np.array_equal(got, ref), f"texture {idx} of {ntex}: {int((got != ref).any(axis=1).sum())} lookups differ"
    # a different texture in every lane of a warp (what the shading kernel does)
    got = step.sample_texture(3, uv, per_lane=True)
    ref = np.zeros_like(got)
    for idx in range(ntex):
        sel = (3 + np.arange(len(uv))) % ntex == idx
        ref[sel] = orc.sample_texture(idx, uv[sel])
    assert np.array_equal(got, ref), f"per-lane texture handles: {int((got != ref).any(axis=1).sum())} lookups differ"
    d = rng.randn(50000, 3).astype(np.float32)
    d[:6] = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    assert np.array_equal(step.sky(d), orc.sky(d))
    step.destroy()


def test_non_finite_rays_miss(ptc, standins, oracle):
    """A ray with a NaN/inf component hits nothing (and must not walk the whole tree: fmin/fmax drop NaNs)."""
    from conftest import random_rays
    scene, _ = standins.random_soup(20000, seed=5, textured=False)
    cfg = ptc.make_config(8, 8)
    step = ptc.RTStep.create(scene, cfg)
    orc = oracle.OracleScene(scene, cfg)
    rays = random_rays(4096, 11)
    bad = np.arange(0, 4096, 7)
    rays[bad[0::3], 4] = np.nan
    rays[bad[1::3], 1] = np.inf
    rays[bad[2::3], 6] = -np.inf
    hits, ids = step.trace_rays(rays)
    rh, rids, _ = orc.trace_rays(rays)
    assert (hits[bad, 3] == 0xFFFFFFFF).all()
    assert np.array_equal(hits, rh) and np.array_equal(ids, rids)
    assert (step.trace_rays_any(rays)[bad] == 0).all()


def test_trace_rays_soup_bit_exact(ptc, standins, oracle):
    from conftest import random_rays
    scene, _ = standins.random_soup(20000, seed=5, textured=False)
    cfg = ptc.make_config(8, 8)
    step = ptc.RTStep.create(scene, cfg)
    orc = oracle.OracleScene(scene, cfg)
    rays = random_rays(200000, 9)
    hits, ids = step.trace_rays(rays)
    rh, rids, _ = orc.trace_rays(rays)
    assert (hits[:, 3] != 0xFFFFFFFF).mean() > 0.2
    assert np.array_equal(hits, rh)
    assert np.array_equal(ids, rids)
    occ = step.trace_rays_any(rays)
    assert np.array_equal(occ, (rh[:, 3] != 0xFFFFFFFF).astype(np.uint8))


def test_cornell_config_c1(ptc, standins, oracle):
    """configs[0]: Cornell 512x512, max depth 8 (a 4-frame cut of the 64 spp; full 64 in the bench)."""
    scene, cam = standins.cornell()
    step, ref = render_both(ptc, standins, oracle, scene, cam, 512, 512, 4, max_bounces=8)
    check_all_planes(ptc, step, ref)
    step.destroy()


def test_ray_count_matches_oracle(ptc, standins, oracle):
    scene, cam = standins.cornell()
    W = H = 128
    cfg = ptc.make_config(W, H, max_bounces=4)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg)
    orc = oracle.OracleScene(scene, cfg)
    rays = 0
    for f in (1, 2, 3):
        u = camera.uniforms(frame=f)
        step.update(u); step.render()
        rays += orc.render_frame(u)["rays"]
    step.poll(True)
    s = step.stats()
    assert s["rays_traced"] == rays
    # per frame and part: raygen, (trace + shade) x (MAX_BOUNCES + 1), and the dense sky pass of the escaped paths
    assert s["frames_rendered"] == 3 and s["kernel_launches"] == 3 * s["frame_parts"] * (1 + 2 * 5 + 1)


@pytest.mark.parametrize("nee", [0.0, 0.5])
def test_soup_all_material_branches(ptc, standins, oracle, nee):
    scene, cam = standins.random_soup(6000, seed=3, textured=True)
    step, ref = render_both(ptc, standins, oracle, scene, cam, 320, 200, 3, max_bounces=6, send_to_light_probability=nee)
    check_all_planes(ptc, step, ref)
    step.destroy()


def test_furnace_exact(ptc, standins, oracle):
    ke = (0.75, 0.5, 0.25)
    scene, cam = standins.furnace(ke=ke)
    step, ref = render_both(ptc, standins, oracle, scene, cam, 64, 48, 2)
    got = check_all_planes(ptc, step, ref)
    assert np.array_equal(got["color"][..., :3], np.broadcast_to(np.array(ke, np.float32), (48, 64, 3)))
    assert (got["depth"] == 0).all()      # light hit at bounce 0: first_t stays 0 (pathtracer/pathtracer.wgsl:204-210)


def test_empty_scene_is_sky(ptc, standins, oracle):
    scene, cam = standins.empty()
    step, ref = render_both(ptc, standins, oracle, scene, cam, 96, 64, 1)
    got = check_all_planes(ptc, step, ref)
    assert (got["hit_id"] == 0xFFFFFFFF).all()


def test_accumulation_modes_and_grab(ptc, standins, oracle):
    T = ptc.types
    scene, cam = standins.cornell()
    W = H = 64
    camera = standins.make_camera(cam, W, H)
    mean = ptc.RTStep.create(scene, ptc.make_config(W, H, accum_mode=T.ACCUM_MEAN))
    summ = ptc.RTStep.create(scene, ptc.make_config(W, H, accum_mode=T.ACCUM_SUM))
    for f in range(1, 6):
        u = camera.uniforms(frame=f)
        for s in (mean, summ):
            s.update(u); s.render()
    a, b = mean.read(T.PLANE_COLOR), summ.read(T.PLANE_COLOR)
    assert (a[..., 3] == 5).all() and (b[..., 3] == 5).all()          # .a == number of frames (pathtracer.wgsl:37)
    np.testing.assert_allclose(b[..., :3] / 5.0, a[..., :3], rtol=2e-6, atol=1e-7)
    u = camera.uniforms(frame=6, is_grabbed=1)                          # grabbed: overwrite, alpha back to 1 (:39)
    mean.update(u); mean.render()
    c = mean.read(T.PLANE_COLOR)
    assert (c[..., 3] == 1).all()
    orc = oracle.OracleScene(scene, ptc.make_config(W, H))
    np.testing.assert_array_equal(c, orc.render_frame(u)["color"])
    mean.reset_accumulation()
    assert (mean.read(T.PLANE_COLOR) == 0).all()


def test_resize_and_injected_output(ptc, standins, oracle):
    import torch
    T = ptc.types
    scene, cam = standins.cornell()
    step = ptc.RTStep.create(scene, ptc.make_config(64, 64))
    step.resize(80, 48)
    camera = standins.make_camera(cam, 80, 48)
    out = torch.zeros(48, 80, 4, device="cuda:0", dtype=torch.float32)
    step.set_output(out.data_ptr())                                     # raytracer.rs:86: caller injects the target
    u = camera.uniforms(frame=1)
    step.update(u); step.render(); step.poll(True)
    orc = oracle.OracleScene(scene, ptc.make_config(80, 48))
    ref = orc.render_frame(u)
    assert np.array_equal(out.cpu().numpy(), ref["color"])
    assert np.array_equal(step.read(T.PLANE_HIT_ID), ref["hit_id"])
    bad = ptc.Camera((64, 64), 60.0, (0, 0, 0)).uniforms(frame=1)
    step.update(bad)
    with pytest.raises(ptc.PtcError):
        step.render()                                                   # uniform size != step size


def test_read_async_matches_read(ptc, standins):
    import torch
    T = ptc.types
    scene, cam = standins.cornell()
    W = H = 128
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    camera = standins.make_camera(cam, W, H)
    pinned = torch.empty(H, W, 4, dtype=torch.float32).pin_memory()
    snaps = []
    for f in (1, 2, 3):
        step.update(camera.uniforms(frame=f)); step.render()
        step.read_async(T.PLANE_COLOR, pinned.data_ptr())
        step.read_wait()
        snaps.append(pinned.numpy().copy())
    assert np.array_equal(snaps[-1], step.read(T.PLANE_COLOR))
    assert (snaps[0][..., 3] == 1).all() and (snaps[1][..., 3] == 2).all()


def test_errors_do_not_abort(ptc, standins):
    T = ptc.types
    scene, _ = standins.cornell()
    with pytest.raises(ptc.PtcError):
        ptc.RTStep.create(scene, ptc.make_config(0, 0))
    with pytest.raises(ptc.PtcError):
        ptc.RTStep.create(scene, ptc.make_config(16, 16, samples_per_pixel=0))
    with pytest.raises(ptc.PtcError):
        ptc.RTStep.create(scene, ptc.make_config(16, 16), device=99)
    step = ptc.RTStep.create(scene, ptc.make_config(16, 16))
    with pytest.raises(ptc.PtcError):
        step.render()                                                   # render before update


def test_sponza_1080p_properties(ptc, standins, oracle):
    """BASELINE configs[2] at full resolution: size-independent properties + an oracle check on a band of rows."""
    T = ptc.types
    scene, cam = standins.sponza()
    W, H, frames = 1920, 1080, 3
    cfg = ptc.make_config(W, H, max_bounces=4)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg)
    us = [camera.uniforms(frame=f) for f in range(1, frames + 1)]
    for u in us:
        step.update(u); step.render()
    step.poll(True)
    col = step.read(T.PLANE_COLOR)
    assert (col[..., 3] == frames).all() and np.isfinite(col).all() and (col[..., :3] >= 0).all()
    s = step.stats()
    assert W * H * frames <= s["rays_traced"] <= W * H * frames * 5
    # oracle on rows [500, 532) of the same frames
    orc = oracle.OracleScene(scene, cfg)
    ref = np.zeros((H, W, 4), np.float32)
    for u in us:
        r = orc.render_frame(u, ref, rows=(500, 532))
    band = slice(500, 532)
    assert np.array_equal(step.read(T.PLANE_HIT_ID)[band], r["hit_id"][band])
    assert_radiance(col[band], ref[band])
    assert np.array_equal(col[band], ref[band])
    # frame-sharded SUM accumulation (what 2 GPUs would do) equals the single-step run up to fp32 summation order
    a = ptc.RTStep.create(scene, ptc.make_config(W, H, accum_mode=T.ACCUM_SUM))
    b = ptc.RTStep.create(scene, ptc.make_config(W, H, accum_mode=T.ACCUM_SUM))
    for i, u in enumerate(us):
        s_ = a if i % 2 == 0 else b
        s_.update(u); s_.render()
    tot = a.read(T.PLANE_COLOR) + b.read(T.PLANE_COLOR)
    assert (tot[..., 3] == frames).all()
    np.testing.assert_allclose(tot[..., :3] / frames, col[..., :3], rtol=1e-5, atol=1e-6)


def band_parity(ptc, standins, oracle, scene, cam, W, H, frames, rows, **cfg_kw):
    """Full-resolution render on the GPU; the oracle renders only rows [rows[0], rows[1]) of the same frames."""
    T = ptc.types
    cfg = ptc.make_config(W, H, **cfg_kw)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg)
    us = [camera.uniforms(frame=f) for f in range(1, frames + 1)]
    for u in us:
        step.update(u); step.render()
    step.poll(True)
    col = step.read(T.PLANE_COLOR)
    assert (col[..., 3] == frames).all()
    orc = oracle.OracleScene(scene, cfg)
    ref = np.zeros((H, W, 4), np.float32)
    for u in us:
        r = orc.render_frame(u, ref, rows=rows)
    band = slice(*rows)
    assert np.array_equal(step.read(T.PLANE_HIT_ID)[band], r["hit_id"][band]), "primary hit ids differ"
    finite = np.isfinite(ref[band]).all(axis=-1) & np.isfinite(col[band]).all(axis=-1)
    assert np.array_equal(np.isfinite(ref[band]), np.isfinite(col[band]))
    assert_radiance(col[band][finite], ref[band][finite])
    assert np.array_equal(col[band], ref[band], equal_nan=True), "colour band is not bit-identical"
    for plane, key in ((T.PLANE_DEPTH, "depth"), (T.PLANE_NORMAL, "normal"), (T.PLANE_ALBEDO, "albedo")):
        assert np.array_equal(step.read(plane)[band], r[key][band], equal_nan=True), key
    stats = step.stats()
    step.destroy()
    return stats


def test_dragon_1080p_c2(ptc, standins, oracle):
    """configs[1]: one 871 k-triangle glass mesh, 1920x1080 (2 of the 256 spp), band of 24 rows against the oracle."""
    scene, cam = standins.dragon()
    s = band_parity(ptc, standins, oracle, scene, cam, 1920, 1080, 2, (528, 552))
    assert s["rays_traced"] >= 1920 * 1080 * 2


@pytest.mark.parametrize("nee", [0.0, 0.5])
def test_classroom_1080p_c4(ptc, standins, oracle, nee):
    """configs[3]: ~0.6 M triangles, 79 constant materials, emitters; reference mode and the light-sampling branch."""
    scene, cam = standins.classroom()
    band_parity(ptc, standins, oracle, scene, cam, 1920, 1080, 2, (600, 616), send_to_light_probability=nee)


def test_san_miguel_4k_c5(ptc, standins, oracle):
    """configs[4]: 10.7 M triangles, 287 materials, 323 textures, 3840x2160; band of 8 rows against the oracle."""
    scene, cam = standins.san_miguel()
    assert scene.num_triangles > 10_000_000
    band_parity(ptc, standins, oracle, scene, cam, 3840, 2160, 1, (1200, 1208))
    # scene ingest at scale: the same 10.7 M triangles through the device LBVH build — a different tree, the
    # same image, bit for bit, over the whole 4K frame
    T = ptc.types
    W, H = 3840, 2160
    u = standins.make_camera(cam, W, H).uniforms(frame=1)
    planes = []
    for build in (T.BVH_SAH_HOST, T.BVH_LBVH_DEVICE):
        step = ptc.RTStep.create(scene, ptc.make_config(W, H, bvh_build=build))
        info = step.bvh_info()
        print("san_miguel bvh_build", build, info)
        assert info["built_on_device"] == (build == T.BVH_LBVH_DEVICE) and info["max_depth"] < 64
        step.update(u); step.render(); step.poll(True)
        planes.append((step.read(T.PLANE_COLOR), step.read(T.PLANE_HIT_ID), step.read(T.PLANE_DEPTH)))
        step.destroy()
    for a, b in zip(*planes):
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("which", ["lit_room", "soup", "classroom"])
def test_light_sampling_mis_bit_exact(ptc, standins, oracle, which):
    """PTC_LIGHTS_MIS through the wavefront (k_shade<MIS> queues shadow rays, k_trace<ANY> resolves them with the
    any-hit early-out and adds into the next queue's accumulated colour) equals the oracle's restatement bit for bit;
    the device ray counter includes the shadow rays."""
    if which == "lit_room":
        scene, cam = standins.lit_room()
    elif which == "soup":
        scene, cam = standins.random_soup(6000, seed=3, textured=True)
    else:
        scene, cam = standins.classroom(detail=0.3)
    W, H = 320, 200
    cfg = ptc.make_config(W, H, max_bounces=5, light_sampling=ptc.types.LIGHTS_MIS)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg)
    orc = oracle.OracleScene(scene, cfg)
    ref, rays = None, 0
    for f in (1, 2, 3):
        u = camera.uniforms(frame=f)
        step.update(u); step.render()
        ref = orc.render_frame(u, None if ref is None else ref["color"])
        rays += ref["rays"]
    step.poll(True)
    check_all_planes(ptc, step, ref)
    s = step.stats()
    assert s["rays_traced"] == rays
    assert s["kernel_launches"] == 3 * s["frame_parts"] * (1 + 2 * 6 + 5 + 1)        # one any-hit launch per bounce that can take a light sample; the sky pass
    step.destroy()


def test_light_sampling_mis_rejects_reference_light_branch(ptc, standins):
    scene, _ = standins.lit_room()
    cfg = ptc.make_config(16, 16, light_sampling=ptc.types.LIGHTS_MIS, send_to_light_probability=0.5)
    with pytest.raises(Exception):
        ptc.RTStep.create(scene, cfg)


@pytest.mark.parametrize("build", [1, 2], ids=["lbvh", "ploc"])
@pytest.mark.parametrize("which", ["soup", "sponza", "cornell", "coincident"])
def test_linear_bvh_device_build(ptc, standins, oracle, hostemu, which, build):
    """PTC_BVH_LBVH_DEVICE / PTC_BVH_PLOC_DEVICE: the tree built by CUDA kernels (Morton keys, CUB radix sort,
    then Karras' radix tree + atomic bottom-up fit, or rounds of nearest-partner search / mutual merge / compaction; collapse,
    quantised emit) equals the host restatement of the same build WORD FOR WORD (FNV hash of node and triangle arrays), and
    renders the same image as the oracle, which uses its own SAH tree."""
    from conftest import random_rays
    T = ptc.types
    if which == "soup":
        scene, cam = standins.random_soup(20000, seed=5, textured=True)
    elif which == "sponza":
        scene, cam = standins.sponza(tex_size=64)
    elif which == "cornell":
        scene, cam = standins.cornell()
    else:
        scene, cam = standins.cornell()
        tri = np.array([[1.0, 0.5, -0.3], [1.0, 1.2, 0.0], [1.0, 0.5, 0.3]], np.float32)
        for _ in range(6):
            scene.add_mesh(np.tile(tri, (100, 1)), np.tile([-1, 0, 0], (300, 1)), None, np.arange(300), 1)
    step, ref = render_both(ptc, standins, oracle, scene, cam, 320, 200, 2, max_bounces=5, bvh_build=build)
    info = step.bvh_info()
    assert info["built_on_device"] and info["max_depth"] < 64
    emu = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=build))
    assert emu.bvh_info()["nodes"] == info["nodes"] and emu.bvh_info()["max_depth"] == info["max_depth"]
    assert step.bvh_hash() == emu.bvh_hash(), "device-built BVH differs from the host restatement"
    check_all_planes(ptc, step, ref)
    rays = random_rays(50000, 13)
    hits, ids = step.trace_rays(rays)
    rh, rids, _ = oracle.OracleScene(scene, ptc.make_config(8, 8)).trace_rays(rays)
    assert np.array_equal(hits, rh) and np.array_equal(ids, rids)
    step.destroy()


def test_ptc_reduce_binds_process_nccl(ptc, standins):
    """ptc_reduce with the communicator torch.distributed holds: on one rank the sum over ranks is the image
    itself — this checks the run-time binding to the process's NCCL and the call on the step's stream; the 2-GPU sum is
    exercised by `bench.py --gpus 2` and, without GPUs, by tests/test_multirank_cpu.py."""
    import os
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("process group already initialised")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        t = torch.zeros(1, device="cuda:0"); dist.all_reduce(t); torch.cuda.synchronize()
        comm = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda", 0))._comm_ptr()
        T = ptc.types
        scene, cam = standins.cornell()
        step = ptc.RTStep.create(scene, ptc.make_config(96, 64, accum_mode=T.ACCUM_SUM))
        camera = standins.make_camera(cam, 96, 64)
        for f in (1, 2):
            step.update(camera.uniforms(frame=f)); step.render()
        step.poll(True)
        before = step.read(T.PLANE_COLOR)
        step.reduce(comm, 0); step.poll(True)
        # the sum arrives in the REDUCED plane; the rank's own accumulation is untouched, so a reduce can be repeated
        assert np.array_equal(step.read(T.PLANE_REDUCED), before) and (before[..., 3] == 2).all()
        assert np.array_equal(step.read(T.PLANE_COLOR), before)
        step.update(camera.uniforms(frame=3)); step.render()
        step.reduce(comm, -1); step.poll(True)                            # root < 0: ncclAllReduce; after one more frame
        after = step.read(T.PLANE_COLOR)
        assert (after[..., 3] == 3).all() and np.array_equal(step.read(T.PLANE_REDUCED), after)
        recv = torch.zeros(64, 96, 4, device="cuda:0")                    # caller-injected receive buffer
        step.set_reduce_output(recv.data_ptr()); step.reduce(comm, 0); step.poll(True)
        assert np.array_equal(recv.cpu().numpy(), after)
        step.set_reduce_output(0)
        mean_step = ptc.RTStep.create(scene, ptc.make_config(96, 64))     # PTC_ACCUM_MEAN: running means cannot be summed
        with pytest.raises(ptc.PtcError, match="PTC_ACCUM_SUM"):
            mean_step.reduce(comm, 0)
        mean_step.destroy()
        with pytest.raises(ptc.PtcError):
            step.reduce(0, 0)
        step.destroy()
    finally:
        dist.destroy_process_group()


def test_frame_parts_do_not_change_the_image(ptc, standins, oracle, monkeypatch):
    """A frame is dealt warp by warp to several concurrent wavefronts (frame parts, pt_step.cu): whatever the number of
    parts — and with a frame size that leaves partial tiles and a partial last warp — every plane equals the oracle's."""
    T = ptc.types
    scene, cam = standins.random_soup(6000, seed=3, textured=True)
    W, H, frames = 333, 217, 3                                            # 72 261 pixels: not a multiple of 8, 4 or 32
    cfg = ptc.make_config(W, H, max_bounces=5)
    camera = standins.make_camera(cam, W, H)
    orc = oracle.OracleScene(scene, cfg)
    ref = None
    for f in range(1, frames + 1):
        ref = orc.render_frame(camera.uniforms(frame=f), None if ref is None else ref["color"])
    seen = set()
    for parts in ("1", "2", "3", "5", "8"):
        monkeypatch.setenv("PTC_PARTS", parts)
        step = ptc.RTStep.create(scene, cfg)
        for f in range(1, frames + 1):
            step.update(camera.uniforms(frame=f)); step.render()
            if f == 2:
                step.blit(); step.read(T.PLANE_DISPLAY)                   # work on the ordering stream between frames
        step.poll(True)
        seen.add(step.stats()["frame_parts"])
        check_all_planes(ptc, step, ref)
        step.destroy()
    assert seen == {1, 2, 3, 5, 8}


def test_modes_combined_and_resize(ptc, standins, oracle):
    """Light sampling + device-built LBVH + SUM accumulation together, then a resize to a frame size with a different number
    of frame parts: still the oracle's image."""
    T = ptc.types
    scene, cam = standins.lit_room()
    kw = dict(max_bounces=4, light_sampling=T.LIGHTS_MIS, bvh_build=T.BVH_LBVH_DEVICE, accum_mode=T.ACCUM_SUM)
    step = ptc.RTStep.create(scene, ptc.make_config(400, 300, **kw))
    for (W, H) in ((400, 300), (97, 61), (640, 360)):
        step.resize(W, H)
        camera = standins.make_camera(cam, W, H)
        orc = oracle.OracleScene(scene, ptc.make_config(W, H, **kw))
        ref = None
        for f in (1, 2):
            u = camera.uniforms(frame=f)
            step.update(u); step.render()
            ref = orc.render_frame(u, None if ref is None else ref["color"])
        step.poll(True)
        check_all_planes(ptc, step, ref)
        assert (step.read(T.PLANE_COLOR)[..., 3] == 2).all()
    step.destroy()


def test_checkpoint_resume_bit_exact(ptc, standins):
    """ptc_write: a progressive render stopped after 2 frames, its accumulation image saved, loaded into a NEW step and
    continued with frames 3-4, equals the uninterrupted 4-frame render bit for bit (running-mean and sum modes)."""
    T = ptc.types
    scene, cam = standins.cornell()
    W, H = 160, 120
    camera = standins.make_camera(cam, W, H)
    for mode in (T.ACCUM_MEAN, T.ACCUM_SUM):
        cfg = ptc.make_config(W, H, accum_mode=mode)
        whole = ptc.RTStep.create(scene, cfg)
        for f in (1, 2, 3, 4):
            whole.update(camera.uniforms(frame=f)); whole.render()
        first = ptc.RTStep.create(scene, cfg)
        for f in (1, 2):
            first.update(camera.uniforms(frame=f)); first.render()
        saved = first.read(T.PLANE_COLOR)
        first.destroy()
        second = ptc.RTStep.create(scene, cfg)
        second.write(saved)
        for f in (3, 4):
            second.update(camera.uniforms(frame=f)); second.render()
        a, b = whole.read(T.PLANE_COLOR), second.read(T.PLANE_COLOR)
        assert np.array_equal(a, b) and (a[..., 3] == 4).all()
        with pytest.raises(ptc.PtcError):
            second.write(saved, plane=T.PLANE_DEPTH)
        whole.destroy(); second.destroy()


def test_cpp_harness_matches_python_mirror(ptc, standins, tmp_path):
    """The compiled host (harness/render_main.cpp over include/ptcore.hpp: Scene::load_obj from a file, Camera::new, the
    Raytracer::render loop, blit) produces the same accumulation image, bit for bit, as the ctypes mirror the tests use."""
    import os
    import subprocess
    T = ptc.types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "wgpu-pathtracer_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "ptrender"], stdout=subprocess.DEVNULL)
    # the Cornell stand-in as OBJ + MTL files, and its environment map as raw float32 RGBA
    env = standins.make_envmap(128, 64)
    room = [
        ("floor", "white", standins.grid((0, 0, -1), (0, 0, 2), (2, 0, 0), 1, 1)),
        ("back", "white", standins.grid((2, 0, -1), (0, 2, 0), (0, 0, 2), 1, 1)),
        ("left", "red", standins.grid((0, 0, -1), (2, 0, 0), (0, 2, 0), 1, 1)),
        ("box", "blue", standins.transform(standins.box((-0.3, 0, -0.3), (0.3, 1.2, 0.3)), rot_y=-0.35, translate=(1.35, 0, -0.35))),
    ]
    obj_text = standins._obj_text(room).replace("mtllib standin.mtl", "mtllib room.mtl")
    (tmp_path / "room.obj").write_text(obj_text)
    (tmp_path / "room.mtl").write_text(standins.CORNELL_MTL)
    env.texels.astype(np.float32).tofile(tmp_path / "sky.f32")
    W, H, frames = 200, 120, 3
    pos, look = (-2.6, 1.0, 0.0), (1.0, 1.0, 0.0)
    r = subprocess.run([os.path.join(csrc, "ptrender"), "--obj", str(tmp_path / "room.obj"), "--env-raw", str(tmp_path / "sky.f32"), "128", "64",
                        "--size", f"{W}x{H}", "--frames", str(frames), "--pos", *map(str, pos), "--look", *map(str, look), "--fov", "60",
                        "--out", str(tmp_path / "cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "cpp.rgba32f", np.float32).reshape(H, W, 4)
    scene = ptc.Scene(env)
    scene.load_obj(obj_text=obj_text, mtl_text=standins.CORNELL_MTL, obj_dir=str(tmp_path))
    camera = standins.make_camera({"position": pos, "fov": 60.0, "target": look}, W, H)
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    for f in range(1, frames + 1):
        step.update(camera.uniforms(frame=f)); step.render()
    ref = step.read(T.PLANE_COLOR)
    assert (ref[..., 3] == frames).all() and np.array_equal(got, ref)
    step.blit()
    ppm = (tmp_path / "cpp.ppm").read_bytes()
    assert np.array_equal(np.frombuffer(ppm[-W * H * 3:], np.uint8).reshape(H, W, 3), step.read(T.PLANE_DISPLAY)[..., :3])
    step.destroy()


@pytest.mark.parametrize("window", [1, 9, 64])
def test_ploc_window_is_honoured_on_device_and_host(ptc, standins, oracle, hostemu, window):
    """ptc_config.ploc_window reaches both builds: device tree == host restatement word for word at every window, trees differ
    between windows, images do not"""
    scene, cam = standins.random_soup(20000, seed=5, textured=True)
    T = ptc.types
    step, ref = render_both(ptc, standins, oracle, scene, cam, 160, 100, 1, max_bounces=4, bvh_build=T.BVH_PLOC_DEVICE, ploc_window=window)
    emu = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=T.BVH_PLOC_DEVICE, ploc_window=window))
    assert step.bvh_hash() == emu.bvh_hash()
    default = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=T.BVH_PLOC_DEVICE))
    assert default.bvh_hash() != emu.bvh_hash()              # the default window is 4
    check_all_planes(ptc, step, ref)
    step.destroy()


def test_non_finite_throughput_is_carried_like_the_oracle(ptc, standins, oracle):
    """The accumulated-colour queue entry is only written when it is not all zero bits (k_shade).  Without light sampling that
    happens when the throughput overflows: an albedo of 1e30 makes `mask` inf after two bounces and `acc += 0 * mask` NaN.  The
    kernels must carry exactly that: every plane equal to the oracle's, NaNs and infinities included."""
    T = ptc.types
    env = standins.make_envmap(64, 32)
    sc = ptc.Scene(env)
    sc.load_obj(obj_text=standins._obj_text([("room", "huge", standins.box((-1, -1, -1), (1, 1, 1))),
                                             ("slab", "plain", standins.grid((-0.5, -0.2, -0.5), (1, 0, 0), (0, 0, 1), 2, 2))]),
                mtl_text="newmtl huge\nKd 1 1 1\nPr 0.8\nmap_Kd huge.f32tex\n\nnewmtl plain\nKd 0.7 0.6 0.5\nPr 0.6\n", obj_dir="standin",
                decode=lambda path: ptc.Texture(np.full((2, 2, 4), 1e30, np.float32), path=path))
    cam = {"position": (0.0, 0.3, 0.0), "fov": 70.0, "target": (1.0, 0.0, 0.2)}
    step, ref = render_both(ptc, standins, oracle, sc, cam, 160, 96, 2, max_bounces=5)
    got = {"color": step.read(T.PLANE_COLOR), "albedo": step.read(T.PLANE_ALBEDO), "depth": step.read(T.PLANE_DEPTH), "normal": step.read(T.PLANE_NORMAL)}
    step.destroy()
    for k in got:
        assert np.array_equal(got[k], ref[k], equal_nan=True), f"{k} plane differs"
    rgb = ref["color"][..., :3]
    assert np.isnan(rgb).any() and np.isfinite(rgb).any()          # the fixture does produce non-finite radiance next to ordinary pixels


@pytest.mark.parametrize("build", [0, 1, 2])
def test_degenerate_and_tied_geometry_gpu(ptc, standins, oracle, build):
    """Coincident duplicates (equal t: smaller triangle id wins), zero-area and sliver triangles, coplanar overlaps and a
    far-away triangle that stretches the quantisation grid: kernels == oracle on every plane, for both tree builders."""
    from test_hostemu_parity import _degenerate_scene
    scene, cam = _degenerate_scene(ptc, standins)
    step, ref = render_both(ptc, standins, oracle, scene, cam, 320, 240, 2, max_bounces=5, bvh_build=build)
    check_all_planes(ptc, step, ref)
    step.destroy()


def test_many_frames_in_flight(ptc, standins, oracle):
    """150 frames enqueued back to back without a poll (more than the ring of per-frame device counters holds), two frame
    parts: the running mean after the last one is the oracle's, bit for bit."""
    T = ptc.types
    scene, cam = standins.cornell()
    W, H, frames = 160, 120, 150
    cfg = ptc.make_config(W, H, max_bounces=3)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(scene, cfg)
    us = [camera.uniforms(frame=f) for f in range(1, frames + 1)]
    for u in us:
        step.update(u); step.render()
    orc = oracle.OracleScene(scene, cfg)
    ref = None
    for u in us:
        ref = orc.render_frame(u, None if ref is None else ref["color"], want_aux=False)
    step.poll(True)
    got = step.read(T.PLANE_COLOR)
    assert step.stats()["frame_parts"] == 2 and (got[..., 3] == frames).all()
    assert np.array_equal(got, ref["color"])
    step.destroy()


def test_samples_per_pixel_loop(ptc, standins, oracle):
    """SAMPLES_PER_PIXEL > 1 (pathtracer/pathtracer.wgsl:258-266): the samples of a pixel re-trace the same primary ray with
    the RNG stream running on, their mean enters the accumulation once per frame, aux planes come from the last sample."""
    scene, cam = standins.random_soup(3000, seed=3)
    for spp in (2, 3):
        step, ref = render_both(ptc, standins, oracle, scene, cam, 96, 72, 2, samples_per_pixel=spp, max_bounces=3)
        check_all_planes(ptc, step, ref)
        assert (ref["color"][..., 3] == 2).all()
        step.destroy()


def test_blit_by_parts_equals_whole_frame_blit(ptc, standins):
    """ptc_blit right after ptc_render tonemaps per frame part on the parts' streams; the DISPLAY plane must equal the one
    the single whole-frame kernel produces (explicit src pointer = ordering-stream path), also while read_async snapshots
    are in flight and on an odd frame size (right / bottom strips outside the 8x4 tiles)."""
    import torch
    T = ptc.types
    scene, cam = standins.cornell()
    W, H = 203, 131
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    camera = standins.make_camera(cam, W, H)
    pinned = torch.empty(H, W, 4, dtype=torch.uint8).pin_memory()
    shots = []
    for f in range(1, 6):
        step.update(camera.uniforms(frame=f)); step.render(); step.blit()
        step.read_async(T.PLANE_DISPLAY, pinned.data_ptr()); step.read_wait()
        shots.append(pinned.numpy().copy())
    step.poll(True)
    by_parts = step.read(T.PLANE_DISPLAY)
    assert np.array_equal(by_parts, shots[-1])
    step.blit(step.device_ptr(T.PLANE_COLOR)); step.poll(True)
    assert np.array_equal(step.read(T.PLANE_DISPLAY), by_parts)
    assert step.stats()["frame_parts"] >= 2 and any(not np.array_equal(shots[0], x) for x in shots[1:])
    step.destroy()


def test_copy_plane_with_row_pitch(ptc, standins):
    """ptc_copy_plane: device-to-device copy into a caller buffer whose rows are rounded up to 256 bytes (pt.rs:470-472)."""
    import torch
    T = ptc.types
    scene, cam = standins.cornell()
    W, H = 100, 40                                   # 1600-byte rows -> 1792-byte pitch
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    camera = standins.make_camera(cam, W, H)
    step.update(camera.uniforms(frame=1)); step.render()
    pitch = (W * 16 + 255) // 256 * 256
    for plane in (T.PLANE_NORMAL, T.PLANE_ALBEDO, T.PLANE_COLOR):
        buf = torch.full((H, pitch // 4), -7.0, device="cuda:0")
        step.copy_plane(plane, buf.data_ptr(), pitch); step.poll(True)
        got = buf.cpu().numpy()
        assert np.array_equal(got[:, :W * 4].reshape(H, W, 4), step.read(plane)) and (got[:, W * 4:] == -7.0).all()
    with pytest.raises(ptc.PtcError):
        step.copy_plane(T.PLANE_COLOR, buf.data_ptr(), 64)
    step.destroy()


def test_resize_failure_keeps_the_step_usable(ptc, standins, oracle):
    """A bad size is refused before anything is freed; the step renders on at its old size."""
    T = ptc.types
    scene, cam = standins.cornell()
    step = ptc.RTStep.create(scene, ptc.make_config(64, 48))
    camera = standins.make_camera(cam, 64, 48)
    step.update(camera.uniforms(frame=1)); step.render(); step.poll(True)
    before = step.read(T.PLANE_COLOR)
    for w, h in ((0, 10), (70000, 70000)):
        with pytest.raises(ptc.PtcError):
            step.resize(w, h)
    assert np.array_equal(step.read(T.PLANE_COLOR), before)
    step.update(camera.uniforms(frame=2)); step.render(); step.poll(True)
    assert (step.read(T.PLANE_COLOR)[..., 3] == 2).all()
    step.destroy()


@pytest.mark.parametrize("build", [0, 1, 2], ids=["sah", "lbvh", "ploc"])
def test_inactive_triangles_gpu(ptc, standins, oracle, build):
    """A mesh whose triangles have NaN / infinite / beyond-1e37 corners is inactive (never hit, as in the acceleration structures
    the reference builds; csrc/scene_prep.cpp): the kernels render the Cornell box with such a mesh appended exactly as the
    oracle renders the box without it, under the host SAH build and both device builds; mesh numbering is unchanged."""
    from test_hostemu_parity import _inactive_mesh
    clean, cam = standins.cornell()
    dirty, _ = standins.cornell()
    bad = _inactive_mesh()
    dirty.add_mesh(bad, np.tile([0, 1, 0], (len(bad), 1)), None, np.arange(len(bad)), 0)
    W, H = 160, 120
    cfg = ptc.make_config(W, H, max_bounces=4, bvh_build=build)
    camera = standins.make_camera(cam, W, H)
    step = ptc.RTStep.create(dirty, cfg, device=0)
    orc = oracle.OracleScene(clean, cfg)
    ref = None
    for f in (1, 2):
        u = camera.uniforms(frame=f)
        step.update(u); step.render()
        ref = orc.render_frame(u, None if ref is None else ref["color"])
    step.poll(True)
    check_all_planes(ptc, step, ref)
    assert step.bvh_info()["triangles"] == len(dirty.indices) // 3
    step.destroy()
