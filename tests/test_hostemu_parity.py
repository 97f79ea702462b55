"""CPU-only parity: the product's device functions (csrc/pt_device.cuh compiled for the host by tests/hostemu) and
its BVH builder against the oracle.  Bit-exact on every plane — what the -m gpu tests then confirm for the kernels."""
import numpy as np
import pytest


def both(ptc, standins, oracle, hostemu, scene, cam, W, H, frames, **kw):
    cfg = ptc.make_config(W, H, **kw)
    camera = standins.make_camera(cam, W, H)
    orc, emu = oracle.OracleScene(scene, cfg), hostemu.scene(scene, cfg)
    ro = re = None
    for f in range(1, frames + 1):
        u = camera.uniforms(frame=f)
        ro = orc.render_frame(u, None if ro is None else ro["color"])
        re = emu.render_frame(u, None if re is None else re["color"])
    return ro, re


def assert_identical(ro, re):
    for k in ("hit_id", "color", "depth", "normal", "albedo"):
        assert np.array_equal(ro[k], re[k], equal_nan=True), k
    assert ro["rays"] == re["rays"]


def test_cornell_depth8(ptc, standins, oracle, hostemu):
    scene, cam = standins.cornell()
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 160, 160, 3, max_bounces=8)
    assert_identical(ro, re)
    assert ro["rays"] > 160 * 160 * 3


def test_cornell_without_normals_reference_bug(ptc, standins, oracle, hostemu):
    """OBJ without vn: scene.rs:202-215 corrupts normals -> NaN radiance in both, at the same pixels."""
    scene, cam = standins.cornell(with_normals=False)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 64, 64, 1)
    assert_identical(ro, re)
    fixed, _ = standins.cornell(with_normals=False, fix_normals=True)
    ro2, re2 = both(ptc, standins, oracle, hostemu, fixed, cam, 64, 64, 1)
    assert_identical(ro2, re2)
    assert np.isfinite(ro2["color"]).all()
    assert np.array_equal(ro["hit_id"], ro2["hit_id"])                 # geometry identical, only normals differ


@pytest.mark.parametrize("nee", [0.0, 0.5, 1.0])
def test_soup_textured_all_branches(ptc, standins, oracle, hostemu, nee):
    scene, cam = standins.random_soup(5000, seed=7, textured=True)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 200, 120, 2, max_bounces=6, send_to_light_probability=nee)
    assert_identical(ro, re)


def test_sum_mode(ptc, standins, oracle, hostemu):
    scene, cam = standins.cornell()
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 48, 48, 3, accum_mode=ptc.types.ACCUM_SUM)
    assert_identical(ro, re)
    assert (ro["color"][..., 3] == 3).all()


def test_sponza_standin_band(ptc, standins, oracle, hostemu):
    scene, cam = standins.sponza(tex_size=128)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 384, 216, 2)
    assert_identical(ro, re)
    assert (ro["hit_id"][..., 0] != 0xFFFFFFFF).mean() > 0.9


@pytest.mark.parametrize("name,kw,cfg", [
    ("dragon", {"n": 120}, {}),
    ("classroom", {"detail": 0.04}, {}),
    ("classroom", {"detail": 0.04}, {"send_to_light_probability": 0.5}),      # emitters present: exercises the light list
    ("san_miguel", {"n_plants": 12, "leaves_per_plant": 400, "tex_size": 32}, {}),
])
def test_reduced_baseline_standins(ptc, standins, oracle, hostemu, name, kw, cfg):
    """Small cuts of the C2 / C4 / C5 stand-ins (same generators, fewer triangles): every plane bit-identical."""
    scene, cam = standins.BUILDERS[name](**kw)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 160, 90, 2, **cfg)
    assert_identical(ro, re)
    assert (ro["hit_id"][..., 0] != 0xFFFFFFFF).mean() > 0.5


def test_trace_rays_product_bvh_vs_brute(ptc, standins, oracle, hostemu):
    from conftest import random_rays
    scene, _ = standins.random_soup(30000, seed=11, textured=False, tri_size=0.12)
    cfg = ptc.make_config(4, 4)
    orc, emu = oracle.OracleScene(scene, cfg), hostemu.scene(scene, cfg)
    rays = random_rays(60000, 21)
    he, ie = emu.trace_rays(rays)
    ho, io, _ = orc.trace_rays(rays[:3000], brute=True)                 # brute force on a subset, BVH oracle on all
    assert np.array_equal(he[:3000], ho) and np.array_equal(ie[:3000], io)
    hv, iv, _ = orc.trace_rays(rays)
    assert np.array_equal(he, hv) and np.array_equal(ie, iv)
    assert np.array_equal(emu.trace_rays_any(rays), (hv[:, 3] != 0xFFFFFFFF).astype(np.uint8))
    info = emu.bvh_info()
    assert info["triangles"] == 30000 and info["max_depth"] < 64


def test_axis_parallel_and_degenerate_rays(ptc, standins, oracle, hostemu):
    """Direction components that are exactly 0, -0 or tiny: the fma slab test must stay conservative."""
    scene, _ = standins.sponza(tex_size=16)
    cfg = ptc.make_config(4, 4)
    orc, emu = oracle.OracleScene(scene, cfg), hostemu.scene(scene, cfg)
    rng = np.random.RandomState(3)
    n = 6000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform([-14, 0.2, -5.5], [14, 13, 5.5], size=(n, 3))
    d = rng.randn(n, 3)
    kind = rng.randint(0, 6, n)
    for k in range(n):
        ax = rng.randint(0, 3)
        if kind[k] == 0: d[k, ax] = 0.0
        elif kind[k] == 1: d[k, ax] = -0.0
        elif kind[k] == 2: d[k, ax] = 1e-30
        elif kind[k] == 3: d[k, ax] = 1e-7 * rng.randn()
        elif kind[k] == 4: d[k] = 0; d[k, ax] = rng.choice([-1.0, 1.0])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    he, ie = emu.trace_rays(rays)
    hv, iv, _ = orc.trace_rays(rays, brute=True)
    assert np.array_equal(he, hv) and np.array_equal(ie, iv)


def test_empty_and_single_triangle_scenes(ptc, standins, oracle, hostemu):
    scene, cam = standins.empty()
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 32, 24, 1)
    assert_identical(ro, re)
    sc = ptc.Scene(standins.make_envmap(32, 16))
    sc.add_mesh(np.array([[2, -1, -1], [2, 1, -1], [2, 0, 1.5]], np.float32), np.tile([-1, 0, 0], (3, 1)), None, [0, 1, 2], 0)
    ro, re = both(ptc, standins, oracle, hostemu, sc, {"position": (0, 0, 0), "fov": 60.0, "target": (1, 0, 0)}, 32, 24, 2)
    assert_identical(ro, re)
    assert (ro["hit_id"][..., 0] == 0).any()


@pytest.mark.parametrize("which", ["lit_room", "soup", "classroom"])
def test_light_sampling_mis_mode(ptc, standins, oracle, hostemu, which):
    """PTC_LIGHTS_MIS (shadow rays + balance heuristic, a non-parity quality mode defined in include/ptmath.h): the
    product's device code and the oracle's independent restatement agree bit for bit, shadow rays included in the count."""
    if which == "lit_room":
        scene, cam = standins.lit_room()
    elif which == "soup":
        scene, cam = standins.random_soup(5000, seed=7, textured=True)          # glass, emissive and textured materials
    else:
        scene, cam = standins.classroom(detail=0.04)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 120, 80, 3, max_bounces=5, light_sampling=ptc.types.LIGHTS_MIS)
    assert_identical(ro, re)
    r0, _ = both(ptc, standins, oracle, hostemu, scene, cam, 120, 80, 3, max_bounces=5)
    assert ro["rays"] > r0["rays"]                                              # shadow rays were traced
    assert np.array_equal(ro["hit_id"], r0["hit_id"])                           # primary visibility is untouched


@pytest.mark.parametrize("build", [1, 2], ids=["lbvh", "ploc"])
@pytest.mark.parametrize("which", ["soup", "sponza", "cornell", "coincident"])
def test_linear_bvh_build_same_images(ptc, standins, oracle, hostemu, which, build):
    """PTC_BVH_LBVH_DEVICE, host restatement of the device build: a different tree, the same hits —
    every plane bit-identical to the oracle, which uses its own SAH tree."""
    if which == "soup":
        scene, cam = standins.random_soup(5000, seed=7, textured=True)
    elif which == "sponza":
        scene, cam = standins.sponza(tex_size=32)
    elif which == "cornell":
        scene, cam = standins.cornell()
    else:
        # 600 coincident triangles (identical Morton keys: the tree falls back to index order) in front of a wall
        scene, cam = standins.cornell()
        tri = np.array([[1.0, 0.5, -0.3], [1.0, 1.2, 0.0], [1.0, 0.5, 0.3]], np.float32)
        for _ in range(6):
            scene.add_mesh(np.tile(tri, (100, 1)), np.tile([-1, 0, 0], (300, 1)), None, np.arange(300), 1)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 160, 100, 2, max_bounces=5, bvh_build=build)
    assert_identical(ro, re)
    info = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=build)).bvh_info()
    assert info["max_depth"] < 64 and info["triangles"] == len(scene.indices) // 3


@pytest.mark.parametrize("window", [1, 9, 64])
def test_ploc_window_changes_the_tree_not_the_image(ptc, standins, oracle, hostemu, window):
    """ptc_config.ploc_window (PLOC's merge-partner search window): every window builds a different tree and the same image"""
    scene, cam = standins.random_soup(5000, seed=7, textured=True)
    T = ptc.types
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 128, 80, 1, max_bounces=4, bvh_build=T.BVH_PLOC_DEVICE, ploc_window=window)
    assert_identical(ro, re)
    a = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=T.BVH_PLOC_DEVICE, ploc_window=window)).bvh_hash()
    b = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=T.BVH_PLOC_DEVICE)).bvh_hash()
    c = hostemu.scene(scene, ptc.make_config(8, 8, bvh_build=T.BVH_PLOC_DEVICE, ploc_window=4)).bvh_hash()
    assert a != b and b == c                          # 0 means the default window, which is 4


def _degenerate_scene(ptc, standins):
    """Cornell plus geometry that stresses ties and degeneracy: exact duplicates (equal t: the smaller triangle id must win),
    zero-area triangles, a sliver, coplanar overlapping quads, and a far-away huge triangle that stretches the scene box
    (coarser quantisation grid for everything else)."""
    scene, cam = standins.cornell()
    tri = np.array([[1.2, 0.4, -0.4], [1.2, 1.3, 0.0], [1.2, 0.4, 0.4]], np.float32)
    nrm = np.tile([-1, 0, 0], (3, 1))
    for _ in range(3):
        scene.add_mesh(tri, nrm, None, [0, 1, 2], 1)                                       # three coincident copies
    scene.add_mesh(np.array([[1, 1, 0], [1, 1, 0], [1, 1, 0]], np.float32), nrm, None, [0, 1, 2], 2)          # point
    scene.add_mesh(np.array([[0.5, 0.2, -0.5], [0.5, 0.7, 0.0], [0.5, 1.2, 0.5]], np.float32), nrm, None, [0, 1, 2], 2)   # collinear
    scene.add_mesh(np.array([[0.8, 0.1, -0.9], [0.8, 1.9, -0.9], [0.8, 1.0, -0.9 + 1e-6]], np.float32), nrm, None, [0, 1, 2], 3)  # sliver
    quad = np.array([[1.6, 0.2, -0.6], [1.6, 1.0, -0.6], [1.6, 1.0, 0.6], [1.6, 0.2, 0.6]], np.float32)
    scene.add_mesh(quad, np.tile([-1, 0, 0], (4, 1)), None, [0, 1, 2, 0, 2, 3], 2)
    scene.add_mesh(quad + np.array([0, 0.3, 0.2], np.float32), np.tile([-1, 0, 0], (4, 1)), None, [0, 1, 2, 0, 2, 3], 4)   # coplanar overlap
    scene.add_mesh(np.array([[4000, -3000, -3000], [4000, 3000, -3000], [4000, 0, 5000]], np.float32), nrm, None, [0, 1, 2], 1)
    return scene, cam


@pytest.mark.parametrize("build", [0, 1, 2])
def test_degenerate_and_tied_geometry(ptc, standins, oracle, hostemu, build):
    scene, cam = _degenerate_scene(ptc, standins)
    ro, re = both(ptc, standins, oracle, hostemu, scene, cam, 200, 150, 2, max_bounces=5, bvh_build=build)
    assert_identical(ro, re)
    # the duplicates resolve to the first copy (smaller global triangle id) at every pixel that sees them
    first_dup = len(scene.meshes) - 9
    seen = set(np.unique(ro["hit_id"][..., 0]).tolist())
    assert first_dup in seen and first_dup + 1 not in seen and first_dup + 2 not in seen


@pytest.mark.parametrize("which", ["sponza_tiles", "soup_random", "degenerate"])
def test_packet_traversal_same_hits(ptc, standins, oracle, hostemu, which):
    """Host model of k_trace_packet (32 rays walking the tree together with one uniform node and stack): whatever the
    packet contains — a coherent 8x4 tile of primary rays, 32 unrelated rays, NaN rays, ties — every ray's hit equals
    the per-ray traversal's and the oracle's."""
    import ctypes as C
    from conftest import random_rays
    L = hostemu.L
    L.emu_packet_trace.restype = C.c_int
    L.emu_packet_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
    L.emu_collect_rays.restype = C.c_uint64
    L.emu_collect_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
    if which == "sponza_tiles":
        scene, cam = standins.sponza(tex_size=16)
        W, H = 240, 136
        es = hostemu.scene(scene, ptc.make_config(W, H))
        un = standins.make_camera(cam, W, H).uniforms(frame=1)
        buf = np.zeros((W * H, 8), np.float32)
        n = int(L.emu_collect_rays(es._h, C.byref(un), 0, buf.ctypes.data_as(C.c_void_p), W * H))
        ys, xs = np.mgrid[0:H, 0:W]
        key = ((ys // 4) * (W // 8) + xs // 8) * 32 + (ys % 4) * 8 + xs % 8
        rays = np.ascontiguousarray(buf[:n][np.argsort(key.ravel(), kind="stable")])
    elif which == "soup_random":
        scene, _ = standins.random_soup(20000, seed=5, textured=False)
        es = hostemu.scene(scene, ptc.make_config(8, 8))
        rays = random_rays(20011, 17)                                    # not a multiple of 32: a partial last packet
        rays[5::97, 4] = np.nan; rays[11::89, 1] = np.inf
    else:
        scene, _ = _degenerate_scene(ptc, standins)
        es = hostemu.scene(scene, ptc.make_config(8, 8))
        rays = random_rays(6000, 23, extent=1.5)
        rays[:, 0:3] = rays[:, 0:3] * 0.3 + np.array([0.6, 1.0, 0.0], np.float32)
    hits = np.zeros((len(rays), 4), np.uint32)
    stats = (C.c_uint64 * 3)()
    L.emu_packet_trace(es._h, rays.ctypes.data_as(C.c_void_p), len(rays), hits.ctypes.data_as(C.c_void_p), stats)
    assert stats[2] == 0
    ref, _, _ = oracle.OracleScene(scene, ptc.make_config(8, 8)).trace_rays(rays)
    assert np.array_equal(hits, ref)
    if which == "sponza_tiles":
        assert stats[0] / (len(rays) / 32) < 90                          # even at this low resolution a tile visits ~2x what one ray does (1.2x at 1080p)


def _inactive_mesh():
    """five triangles in front of the Cornell camera, each with one corner no tree builder can place"""
    return np.array([[np.nan, 0.5, 0.0], [1.0, 0.5, 0.5], [1.0, 1.0, 0.0],            # NaN corner
                     [1.0, np.inf, 0.0], [1.0, 0.5, 0.5], [1.5, 1.0, 0.0],            # +inf
                     [1.0, 0.2, 0.0], [-np.inf, 0.5, 0.5], [1.5, 1.0, 0.3],           # -inf
                     [3e38, 0.2, 0.0], [1.0, 0.5, 0.5], [1.5, 1.0, 0.3],              # finite, but edges overflow fp32
                     [1.0, 0.2, 0.0], [1.0, -2e37, 0.5], [1.5, 1.0, 0.3]], np.float32)


@pytest.mark.parametrize("build", ["sah", "lbvh", "ploc"])
def test_inactive_triangles_do_not_change_the_image(ptc, standins, hostemu, build):
    """Triangles with a NaN / infinite / beyond-1e37 corner are inactive, as in the acceleration structures the reference
    builds (a NaN position makes a Vulkan triangle inactive; pt.rs:391-433 passes the vertex buffer on unchanged): never
    hit, and harmless to every builder (found by tools/fuzz_scene.cpp: the SAH binning converted a NaN to int).  The scene
    with such a mesh appended renders bit-identically to the scene without it; numbering of the other meshes is unchanged."""
    T = ptc.types
    kind = {"sah": T.BVH_SAH_HOST, "lbvh": T.BVH_LBVH_DEVICE, "ploc": T.BVH_PLOC_DEVICE}[build]
    clean, cam = standins.cornell()
    dirty, _ = standins.cornell()
    bad = _inactive_mesh()
    dirty.add_mesh(bad, np.tile([0, 1, 0], (len(bad), 1)), None, np.arange(len(bad)), 0)
    W, H = 96, 96
    cfg = ptc.make_config(W, H, max_bounces=4, bvh_build=kind)
    camera = standins.make_camera(cam, W, H)
    a, b = hostemu.scene(clean, cfg), hostemu.scene(dirty, cfg)
    assert b.bvh_info()["triangles"] == a.bvh_info()["triangles"] + 5                      # the five triangles are in the tree's arrays, just never hit
    u = camera.uniforms(frame=1)
    ra, rb = a.render_frame(u), b.render_frame(u)
    for k in ("hit_id", "color", "depth", "normal", "albedo"):
        assert np.array_equal(ra[k], rb[k], equal_nan=True), k
    assert ra["rays"] == rb["rays"]
    # rays aimed straight at where the finite corners of the inactive triangles are: same answers too
    from conftest import random_rays
    rays = random_rays(4000, seed=11, extent=2.5)
    ha, ia = a.trace_rays(rays)
    hb, ib = b.trace_rays(rays)
    assert np.array_equal(ha, hb) and np.array_equal(ia, ib)
    assert int((ib[:, 0] == len(dirty.meshes) - 1).sum()) == 0
