"""Known-answer tests that pin the oracle's restatement of random.wgsl / camera.rs / the WGSL builtins.

The reference ships no tests or golden vectors: the vectors below were derived by hand from
src/shaders/common/random.wgsl:7-36 and src/camera.rs:54-79 in exact u32 / fp32 arithmetic and are
committed in tests/golden/kat.json together with the script that checks them independently in numpy.
"""
import json
import os

import numpy as np

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def test_jenkins(oracle):
    L = oracle.lib()
    for x, y in GOLD["jenkins"]:
        assert L.orc_jenkins_hash(int(x, 16)) == int(y, 16)


def test_rand_streams(oracle):
    for case in GOLD["rand_streams"]:
        state, ints, floats = oracle.rand_stream(case["x"], case["y"], case["width"], case["frame"], 4)
        assert state == int(case["state"], 16)
        assert [int(v) for v in ints] == [int(v, 16) for v in case["ints"]]
        if "floats" in case:
            np.testing.assert_array_equal(floats[: len(case["floats"])], np.array(case["floats"], dtype=np.float64).astype(np.float32))


def test_rand_float_range(oracle):
    L = oracle.lib()
    # f32(0xffffffff) == 2^32, so rand_float reaches exactly 1.0 for u >= 0xffffff80 (random.wgsl:34-36)
    assert L.orc_rand_float_of(0xFFFFFF80) == 1.0
    assert L.orc_rand_float_of(0xFFFFFFFF) == 1.0
    assert np.float32(L.orc_rand_float_of(0xFFFFFF7F)) == np.float32(0.99999994)
    assert L.orc_rand_float_of(0) == 0.0
    assert L.orc_rand_float_of(0x80000000) == 0.5


def test_rng_matches_numpy_restatement(oracle):
    """Independent numpy restatement of random.wgsl (uint32 wrap-around) for a grid of pixels and frames."""
    def jenkins(x):
        x = np.uint32(x)
        with np.errstate(over="ignore"):
            x = x + (x << np.uint32(10)); x ^= x >> np.uint32(6); x = x + (x << np.uint32(3)); x ^= x >> np.uint32(11); x = x + (x << np.uint32(15))
        return np.uint32(x)

    def pcg(state):
        with np.errstate(over="ignore"):
            state = np.uint32(state * np.uint32(747796405) + np.uint32(2891336453))
            word = np.uint32(((state >> np.uint32((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737))
        return state, np.uint32((word >> np.uint32(22)) ^ word)

    rng = np.random.RandomState(0)
    for _ in range(200):
        x, y, w, f = int(rng.randint(0, 4096)), int(rng.randint(0, 2160)), int(rng.choice([512, 800, 1920, 3840])), int(rng.randint(0, 100000))
        with np.errstate(over="ignore"):
            seed = np.uint32(np.uint32(x) + np.uint32(y) * np.uint32(w)) ^ jenkins(f)
        st = jenkins(seed)
        state, ints, floats = oracle.rand_stream(x, y, w, f, 3)
        assert state == int(st)
        for k in range(3):
            st, v = pcg(st)
            assert int(ints[k]) == int(v)
            assert floats[k] == np.float32(np.float32(v) / np.float32(4294967296.0))


def test_camera_kat(oracle):
    for case in GOLD["camera"]:
        u = oracle.camera_uniforms(case["width"], case["height"], 60.0, (0, 0, 0))
        c = u.camera
        np.testing.assert_allclose(list(c.first_pixel_pos), case["first_pixel"], rtol=2e-7, atol=0)
        np.testing.assert_allclose(c.pixel_space_x[2], case["psx_z"], rtol=2e-7)
        np.testing.assert_allclose(c.pixel_space_y[1], -case["psx_z"] * (1.0 if case["width"] == case["height"] else c.pixel_space_y[1] / -c.pixel_space_x[2]), rtol=1e-6)
        assert c.pixel_space_x[0] == 0 and c.pixel_space_x[1] == 0 and c.pixel_space_y[0] == 0 and c.pixel_space_y[2] == 0
        assert list(c.lookat) == [1.0, 0.0, 0.0]          # steady state: looking down +X
        assert c.width == case["width"] and c.height == case["height"]


def test_camera_matches_product_host_code(oracle, ptc):
    """The product's camera (csrc/pt_host.cpp) and the oracle's restatement of camera.rs agree bit for bit."""
    for (w, h, fov, pos, tx, ty) in [(512, 512, 60.0, (0, 0, 0), 0.0, 0.0), (1920, 1080, 60.0, (-3.5, 1.0, 0.25), 0.3, -0.2),
                                     (3840, 2160, 45.0, (10, 2, -7), 2.5, 0.6), (800, 600, 90.0, (0, 5, 0), -1.0, 1.2)]:
        cam = ptc.Camera((w, h), fov, pos)
        cam.set(theta_x=tx, theta_y=ty)
        for _ in range(3):
            cam.update()
        a = cam.uniforms(frame=0)
        b = oracle.camera_uniforms(w, h, fov, pos, tx, ty, extra_updates=3)
        assert bytes(a) == bytes(b)


def test_transcendentals_accuracy(oracle):
    """pt_sin/cos/atan/acos/atan2 (include/ptmath.h) against float64 libm: a few ulp over the shader's ranges."""
    rng = np.random.RandomState(1)

    def ulps(got, want):
        want32 = want.astype(np.float32)
        spacing = np.spacing(np.maximum(np.abs(want32), np.float32(1e-30)))
        return np.abs(got.astype(np.float64) - want) / spacing

    x = np.concatenate([rng.uniform(0, 2 * np.pi, 20000), rng.uniform(-50, 50, 5000), [0.0, np.pi / 2, np.pi, 2 * np.pi]]).astype(np.float32)
    for fn, ref in (("sin", np.sin), ("cos", np.cos)):
        got = oracle.math_fn(fn, x)
        err = np.abs(got.astype(np.float64) - ref(x.astype(np.float64)))
        assert err.max() < 2.5e-7, (fn, err.max())       # absolute: relative ulps are meaningless near the zeros
    x = np.concatenate([rng.uniform(-20, 20, 20000), 10.0 ** rng.uniform(-6, 6, 5000), [0.0, 1.0, -1.0, np.inf, -np.inf]]).astype(np.float32)
    assert ulps(oracle.math_fn("atan", x), np.arctan(x.astype(np.float64))).max() <= 3
    x = np.concatenate([rng.uniform(-1, 1, 20000), [1.0, -1.0, 0.0, 0.5, -0.5]]).astype(np.float32)
    got, want = oracle.math_fn("acos", x), np.arccos(x.astype(np.float64))
    assert (np.abs(got - want) < 4e-7).all()
    assert oracle.math_fn("acos", np.array([1.0000001, -1.0000001], np.float32)).tolist() == [0.0, np.float32(np.pi)]   # clamped, not NaN
    y, xx = rng.uniform(-1, 1, 20000).astype(np.float32), rng.uniform(-1, 1, 20000).astype(np.float32)
    got, want = oracle.math_fn("atan2", y, xx), np.arctan2(y.astype(np.float64), xx.astype(np.float64))
    assert (np.abs(got - want) < 6e-7).all()


def test_camera_against_an_independent_restatement(ptc):
    """Camera::new + Camera::update (reference src/camera.rs:31-79) written again in numpy float32, sharing nothing with
    oracle/ or csrc/: note the order — the basis is built from the CURRENT lookat, and only then is lookat moved to where
    theta_x / theta_y point, so a new pose takes effect one update later.  Random poses, several updates; numpy's float32
    tan / sin / cos may differ from libm's in the last place, hence 1e-6 relative instead of bytes."""
    F = np.float32
    rng = np.random.RandomState(11)

    def norm(a):
        return (a * (F(1.0) / F(np.sqrt(np.dot(a, a))))).astype(F)                  # glam: v * length_recip

    def update(c):
        vh = F(2.0) * F(np.tan(F(c["fov"] * F(np.pi / 180.0)) / F(2.0)))
        vw = F(vh * (F(c["w"]) / F(c["h"])))
        vz = norm((c["pos"] - c["lookat"]).astype(F))
        vx = norm(np.cross(np.array([0, 1, 0], F), vz).astype(F))
        vy = np.cross(vz, vx).astype(F)
        xv, yv = (vx * vw).astype(F), (-vy * vh).astype(F)
        c["psx"], c["psy"] = (xv / F(c["w"])).astype(F), (yv / F(c["h"])).astype(F)
        top_left = (c["pos"] - vz - xv / F(2.0) - yv / F(2.0)).astype(F)
        c["first"] = (top_left + F(0.5) * (c["psx"] + c["psy"])).astype(F)
        tx, ty = F(c["tx"]), F(c["ty"])
        c["lookat"] = np.array([c["pos"][0] + F(np.cos(tx)) * F(np.cos(ty)), c["pos"][1] + F(np.sin(ty)), c["pos"][2] + F(np.sin(tx)) * F(np.cos(ty))], F)

    for _ in range(40):
        w, h = int(rng.randint(16, 4000)), int(rng.randint(16, 2200))
        fov = float(rng.uniform(20, 110))
        pos = rng.uniform(-20, 20, 3).astype(F)
        tx, ty = float(rng.uniform(-3.1, 3.1)), float(rng.uniform(-1.4, 1.4))
        mine = {"w": w, "h": h, "fov": F(fov), "pos": pos, "lookat": (pos + np.array([0, 0, 1], F)).astype(F), "tx": 0.0, "ty": 0.0}
        update(mine)                                                                # Camera::new ends with update()
        cam = ptc.Camera((w, h), fov, tuple(float(x) for x in pos))
        mine["tx"], mine["ty"] = tx, ty
        cam.set(theta_x=tx, theta_y=ty)
        for _ in range(2):
            update(mine)
            cam.update()
        got = cam.uniforms(frame=5).camera
        for name, want in (("pixel_space_x", mine["psx"]), ("pixel_space_y", mine["psy"]), ("first_pixel_pos", mine["first"]), ("lookat", mine["lookat"])):
            have = np.array(list(getattr(got, name)), F)
            scale = max(1e-3, float(np.abs(want).max()))
            assert np.abs(have - want).max() <= 2e-6 * scale + 1e-7, (name, have, want)
        assert got.width == w and got.height == h and got.frame == 5
