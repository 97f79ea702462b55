"""Analytic fixtures for the oracle (SURVEY §8c "analytic fixtures the oracle should add"): results that follow
from the reference's WGSL by inspection, independent of any implementation detail."""
import numpy as np
import pytest


def cam_uniforms(standins, cam, W, H, frame=1, **kw):
    return standins.make_camera(cam, W, H).uniforms(frame=frame, **kw)


def test_furnace_every_path_returns_ke(ptc, standins, oracle):
    """Closed box of emitters: ray_color adds emissive*mask at the first hit and breaks (pathtracer/pathtracer.wgsl:200-206)."""
    ke = (0.75, 0.5, 0.25)
    scene, cam = standins.furnace(ke=ke)
    W, H = 48, 32
    cfg = ptc.make_config(W, H)
    orc = oracle.OracleScene(scene, cfg)
    assert orc.num_lights == 12                                          # every triangle is a light (pt.rs:217)
    r = orc.render_frame(cam_uniforms(standins, cam, W, H))
    assert np.array_equal(r["color"][..., :3], np.broadcast_to(np.array(ke, np.float32), (H, W, 3)))
    assert (r["color"][..., 3] == 1).all()
    assert (r["depth"] == 0).all() and (r["albedo"][..., :3] == 0).all()  # light hit: break before the i == 0 aux block
    assert r["rays"] == W * H


def test_empty_scene_equals_sky_lookup(ptc, standins, oracle):
    """No geometry: colour = sky_color(primary direction) (sky.wgsl:1-12); checked against an independent float64
    bilinear lookup of the equirect map."""
    scene, cam = standins.empty()
    W, H = 64, 40
    cfg = ptc.make_config(W, H)
    orc = oracle.OracleScene(scene, cfg)
    assert orc.num_lights == 1                                           # the dummy light triangle (pt.rs:232-239)
    u = cam_uniforms(standins, cam, W, H)
    r = orc.render_frame(u)
    assert (r["hit_id"] == 0xFFFFFFFF).all() and r["rays"] == W * H
    env = scene.textures[9].texels.astype(np.float64)
    eh, ew = env.shape[:2]
    c = u.camera
    for (x, y) in [(0, 0), (W - 1, H - 1), (W // 2, H // 2), (7, 31)]:
        _, _, fl = oracle.rand_stream(x, y, W, 1, 2)
        tgt = (np.array(list(c.first_pixel_pos), np.float64) + np.array(list(c.pixel_space_x), np.float64) * (x + 0.5 * float(fl[0]))
               + np.array(list(c.pixel_space_y), np.float64) * (y + 0.5 * float(fl[1])))
        d = tgt - np.array(list(c.position), np.float64)
        d /= np.linalg.norm(d)
        v = np.arccos(d[1]) / np.pi
        uu = (np.arctan2(d[2], d[0]) + np.pi) / (2 * np.pi)
        px, py = uu * ew - 0.5, v * eh - 0.5
        x0, y0 = int(np.floor(px)), int(np.floor(py))
        fx, fy = px - x0, py - y0
        t = lambda xx, yy: env[yy % eh, xx % ew, :3]
        want = (t(x0, y0) * (1 - fx) + t(x0 + 1, y0) * fx) * (1 - fy) + (t(x0, y0 + 1) * (1 - fx) + t(x0 + 1, y0 + 1) * fx) * fy
        np.testing.assert_allclose(r["color"][y, x, :3], want, rtol=0, atol=2e-4)


def test_single_triangle_hit_id_and_barycentrics(ptc, standins, oracle):
    """One triangle facing the camera: ids (mesh 1, triangle 0) inside its silhouette, t and barycentrics analytic."""
    sc = ptc.Scene(standins.make_envmap(32, 16))
    m = sc.add_material(albedo_texture_idx=0)
    P = np.array([[2, -1, -1], [2, 1, -1], [2, 0, 1.5]], np.float32)
    sc.add_mesh(np.zeros((3, 3), np.float32) + 50, np.tile([1, 0, 0], (3, 1)), None, [0, 1, 2], 0)     # far-away filler = mesh 0
    sc.add_mesh(P, np.tile([-1, 0, 0], (3, 1)), np.array([[0, 0], [1, 0], [0, 1]], np.float32), [0, 1, 2], m)
    cfg = ptc.make_config(8, 8)
    orc = oracle.OracleScene(sc, cfg)
    rays = np.zeros((3, 8), np.float32)
    rays[:, 4:7] = [[1, 0, 0], [1, 0.1, 0.2], [1, 0, 2.0]]
    rays[:, 4:7] /= np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True)
    for brute in (True, False):
        hits, ids, _ = orc.trace_rays(rays, brute=brute)
        assert ids.tolist() == [[1, 0], [1, 0], [0xFFFFFFFF, 0xFFFFFFFF]]
        t = hits[:2, 0].view(np.float32)
        np.testing.assert_allclose(t, 2.0 / rays[:2, 4], rtol=1e-6)
        for k in range(2):
            p = rays[k, 4:7].astype(np.float64) * float(t[k])
            A = np.stack([P[1] - P[0], P[2] - P[0]], axis=1).astype(np.float64)[1:]
            b12 = np.linalg.solve(A, (p - P[0])[1:])
            np.testing.assert_allclose(hits[k, 1:3].view(np.float32), b12, atol=1e-6)


def test_tmin_tmax_window(ptc, standins, oracle):
    """RayDesc(0, 0xFF, 0.01, 10000, ..) (rt_helpers.wgsl:49): hits at t <= 0.01 or >= 10000 do not exist."""
    sc = ptc.Scene(standins.make_envmap(32, 16))
    quad = standins.grid((0, -1e5, -1e5), (0, 2e5, 0), (0, 0, 2e5), 1, 1)
    sc.add_mesh(*quad, 0)
    orc = oracle.OracleScene(sc, ptc.make_config(4, 4))
    rays = np.zeros((4, 8), np.float32)
    rays[:, 4] = 1.0
    rays[:, 0] = [-0.005, -0.02, -9999.0, -10001.0]
    hits, ids, _ = orc.trace_rays(rays, brute=True)
    assert (hits[:, 3] != 0xFFFFFFFF).tolist() == [False, True, True, False]


def test_accumulation_running_mean_and_grab(ptc, standins, oracle):
    """pathtracer.wgsl:32-42: alpha counts frames, rgb is the running mean; is_grabbed != 0 overwrites."""
    scene, cam = standins.cornell()
    W = H = 24
    orc = oracle.OracleScene(scene, ptc.make_config(W, H))
    camera = standins.make_camera(cam, W, H)
    acc = np.zeros((H, W, 4), np.float32)
    singles = []
    for f in range(1, 5):
        singles.append(orc.render_frame(camera.uniforms(frame=f))["color"][..., :3].astype(np.float64))
        orc.render_frame(camera.uniforms(frame=f), acc)
        assert (acc[..., 3] == f).all()
    np.testing.assert_allclose(acc[..., :3], np.mean(singles, axis=0), rtol=1e-5, atol=1e-6)
    orc.render_frame(camera.uniforms(frame=9, is_grabbed=1), acc)
    assert (acc[..., 3] == 1).all()
    np.testing.assert_array_equal(acc, orc.render_frame(camera.uniforms(frame=9))["color"])
    # negative / NaN garbage in the target is clamped by max(vec4(0), prev)
    dirty = np.full((H, W, 4), -3.0, np.float32)
    orc.render_frame(camera.uniforms(frame=1), dirty)
    np.testing.assert_array_equal(dirty, orc.render_frame(camera.uniforms(frame=1))["color"])


def test_frames_are_distinct_streams(ptc, standins, oracle):
    scene, cam = standins.cornell()
    orc = oracle.OracleScene(scene, ptc.make_config(32, 32))
    camera = standins.make_camera(cam, 32, 32)
    a = orc.render_frame(camera.uniforms(frame=1))["color"]
    b = orc.render_frame(camera.uniforms(frame=2))["color"]
    a2 = orc.render_frame(camera.uniforms(frame=1))["color"]
    assert np.array_equal(a, a2) and not np.array_equal(a, b)


@pytest.mark.parametrize("seed,n", [(1, 500), (2, 4000)])
def test_oracle_bvh_equals_brute_force(ptc, standins, oracle, seed, n):
    from conftest import random_rays
    scene, _ = standins.random_soup(n, seed=seed, textured=False)
    orc = oracle.OracleScene(scene, ptc.make_config(4, 4))
    rays = random_rays(20000, seed + 10)
    hb, ib, cb = orc.trace_rays(rays, brute=True)
    hv, iv, cv = orc.trace_rays(rays, brute=False)
    assert np.array_equal(hb, hv) and np.array_equal(ib, iv)
    assert cv["tris"] < cb["tris"] / 10
    occ = orc.trace_rays_any(rays)
    assert np.array_equal(occ, (hb[:, 3] != 0xFFFFFFFF).astype(np.uint8))


def test_tie_break_smaller_triangle_id(ptc, standins, oracle):
    """Two coincident triangles: the documented tie rule picks the smaller global id in every traversal order."""
    sc = ptc.Scene(standins.make_envmap(32, 16))
    P = np.array([[2, -1, -1], [2, 1, -1], [2, 0, 1.5]], np.float32)
    for _ in range(3):
        sc.add_mesh(P, np.tile([-1, 0, 0], (3, 1)), None, [0, 1, 2], 0)
    orc = oracle.OracleScene(sc, ptc.make_config(4, 4))
    rays = np.zeros((1, 8), np.float32); rays[0, 4] = 1.0
    for brute in (True, False):
        _, ids, _ = orc.trace_rays(rays, brute=brute)
        assert ids.tolist() == [[0, 0]]


def test_nee_branch_consumes_rng_and_changes_image(ptc, standins, oracle):
    scene, cam = standins.random_soup(800, seed=4, textured=False)
    W, H = 48, 32
    camera = standins.make_camera(cam, W, H)
    a = oracle.OracleScene(scene, ptc.make_config(W, H)).render_frame(camera.uniforms(frame=1))
    b = oracle.OracleScene(scene, ptc.make_config(W, H, send_to_light_probability=0.75)).render_frame(camera.uniforms(frame=1))
    assert np.array_equal(a["hit_id"], b["hit_id"])                    # primary rays identical
    assert not np.array_equal(a["color"], b["color"])
    assert np.isfinite(b["color"]).all()


def test_light_sampling_mis_same_expectation(ptc, standins, oracle):
    """PTC_LIGHTS_MIS must converge to the image of the reference estimator (BSDF sampling alone): a closed room under
    a black sky, lit only by emissive triangles.  Means over 16x12-pixel blocks after 2400 / 800 frames agree within the
    Monte-Carlo noise (measured: global mean 0.2 %, block RMSE 1 % at 3x these frame counts)."""
    scene, cam = standins.lit_room()
    W, H = 32, 24
    camera = standins.make_camera(cam, W, H)
    means = []
    for mode, frames in ((ptc.types.LIGHTS_REFERENCE, 2400), (ptc.types.LIGHTS_MIS, 800)):
        orc = oracle.OracleScene(scene, ptc.make_config(W, H, max_bounces=4, light_sampling=mode))
        img = None
        for f in range(1, frames + 1):
            img = orc.render_frame(camera.uniforms(frame=f), None if img is None else img["color"])
        assert np.isfinite(img["color"]).all()
        means.append(img["color"][..., :3].astype(np.float64))
    a, b = means
    assert a.mean() > 0.02                                                       # the lamp does light the room
    assert abs(b.mean() - a.mean()) / a.mean() < 0.03
    A = a.reshape(2, 12, 2, 16, 3).mean(axis=(1, 3, 4)); B = b.reshape(2, 12, 2, 16, 3).mean(axis=(1, 3, 4))
    assert np.abs(B - A).max() / a.mean() < 0.08


def test_light_sampling_mis_is_less_noisy(ptc, standins, oracle):
    """What the mode is for: at equal frame count the light-sampled image is closer to the converged one."""
    scene, cam = standins.lit_room()
    W, H = 32, 24
    camera = standins.make_camera(cam, W, H)

    def render(mode, frames, first=1):
        orc = oracle.OracleScene(scene, ptc.make_config(W, H, max_bounces=4, light_sampling=mode))
        img = None
        for f in range(first, first + frames):
            img = orc.render_frame(camera.uniforms(frame=f), None if img is None else img["color"])
        return img["color"][..., :3].astype(np.float64)

    truth = render(ptc.types.LIGHTS_MIS, 3000, first=1000)
    err_ref = np.sqrt(np.mean((render(ptc.types.LIGHTS_REFERENCE, 64) - truth) ** 2))
    err_mis = np.sqrt(np.mean((render(ptc.types.LIGHTS_MIS, 64) - truth) ** 2))
    assert err_mis < 0.7 * err_ref, (err_mis, err_ref)


def test_light_sampling_mis_furnace_exact(ptc, standins, oracle):
    """Every camera ray ends on an emitter at its first hit, where the weight is 1: the image is Ke exactly, as in the reference mode."""
    ke = (0.75, 0.5, 0.25)
    scene, cam = standins.furnace(ke=ke)
    orc = oracle.OracleScene(scene, ptc.make_config(40, 30, light_sampling=ptc.types.LIGHTS_MIS))
    img = orc.render_frame(standins.make_camera(cam, 40, 30).uniforms(frame=1))
    assert np.array_equal(img["color"][..., :3], np.broadcast_to(np.array(ke, np.float32), (30, 40, 3)))
