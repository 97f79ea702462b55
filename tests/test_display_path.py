"""The two passes after the path tracer in the reference's frame loop (raytracer.rs:215-241): rayproject.wgsl (forward
reprojection of the last real frame) and blit.wgsl (Khronos PBR-neutral tonemap to the sRGB surface).  SURVEY 8f row 3.

Bars: reprojection target indices bit-exact (integer work); display image within 1 LSB of 8 bits (the sRGB curve needs
pow(), whose last ulp differs between the oracle's double-precision evaluation and the fp32 product path; the
reference's own result depends on the driver's pow and the hardware's sRGB encoder)."""
import numpy as np
import pytest

DISPLAY_TOL_LSB = 1


def _cams(standins, ptc, W, H):
    a = standins.make_camera({"position": (-2.6, 1.0, 0.0), "fov": 60.0, "target": (1.0, 1.0, 0.0)}, W, H)
    b = standins.make_camera({"position": (-2.45, 1.08, 0.12), "fov": 60.0, "target": (1.0, 0.95, 0.1)}, W, H)
    return a, b


def _test_images():
    rng = np.random.RandomState(5)
    img = np.zeros((64, 96, 4), np.float32)
    img[..., :3] = rng.rand(64, 96, 3).astype(np.float32) ** 3 * 4.0          # dark to well above the compression knee
    img[..., 3] = rng.randint(1, 9, (64, 96))
    img[0, :8, :3] = 0.0                                                     # black: NaN through the blended arms -> 0
    img[1, :8, :3] = [[0.04, 0.04, 0.04]] * 8
    img[2, :8, :3] = [[0.76, 0.2, 0.1]] * 8                                  # peak exactly at the knee
    img[3, :8, :3] = np.nan
    img[4, :8, :3] = 1e9
    img[5, :8, :3] = -1.0
    return img


def test_blit_host_matches_oracle(oracle, hostemu):
    img = _test_images()
    ref, _ = oracle.blit(img)
    got = hostemu.blit(img)
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= DISPLAY_TOL_LSB
    assert (got[0, :8, :3] == 0).all() and (got[3, :8, :3] == 0).all() and (got[4, :8, :3] == 255).all()
    assert (got[..., 3] == 255).all()                                        # alpha = sample count >= 1 saturates


def test_tonemap_known_answers(oracle):
    """to_khronos_pbr_neutral: identity minus the toe below the knee; monotone, < 1 above it."""
    x = np.array([[0.5, 0.4, 0.3, 1.0], [0.08, 0.08, 0.08, 1.0], [10.0, 10.0, 10.0, 1.0], [0.3, 0.02, 0.5, 1.0]], np.float32)
    _, tm = oracle.blit(x)
    np.testing.assert_allclose(tm[0], np.array([0.5, 0.4, 0.3]) - 0.04, rtol=1e-6)            # min > 2*F90: offset F90
    np.testing.assert_allclose(tm[1], np.array([0.08] * 3) - 0.04, rtol=1e-5)                 # min == 2*F90: x - x^2/(4 F90) = 0.04
    assert (tm[2] < 1.0).all() and (tm[2] > 0.9).all()
    f = 0.02 - 0.02 * 0.02 / 0.16
    np.testing.assert_allclose(tm[3], np.array([0.3, 0.02, 0.5]) - f, rtol=1e-5)


def test_rayproject_host_matches_oracle(ptc, standins, oracle, hostemu):
    W, H = 96, 64
    a, b = _cams(standins, ptc, W, H)
    rng = np.random.RandomState(2)
    depth = rng.uniform(0.5, 6.0, (H, W)).astype(np.float32)
    depth[:2] = 0.0                                                           # sky / light pixels keep first_t = 0
    depth[2, :4] = np.nan
    ua, ub = a.uniforms(frame=3), b.uniforms(frame=4)
    ref = oracle.rayproject(ub, ua, depth)
    got = hostemu.rayproject(ub, ua, depth)
    assert np.array_equal(got, ref)
    assert (ref != 0xFFFFFFFF).mean() > 0.5
    same = oracle.rayproject(ua, ua, depth)                                   # unmoved camera: never off by more than the
    ys, xs = np.divmod(np.arange(W * H), W)                                   # reference's own +0.5 / round placement
    ok = same != 0xFFFFFFFF
    ok &= np.isfinite(depth.reshape(-1))
    assert (np.abs((same[ok] % W).astype(int) - xs[ok]) <= 1).all() and (np.abs((same[ok] // W).astype(int) - ys[ok]) <= 1).all()


@pytest.mark.gpu
def test_display_path_gpu(ptc, standins, oracle):
    """Render, blit, move the camera, reproject — through the C ABI, against the oracle."""
    T = ptc.types
    scene, _ = standins.cornell()
    W, H = 320, 200
    a, b = _cams(standins, ptc, W, H)
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    with pytest.raises(ptc.PtcError):
        step.read(T.PLANE_DISPLAY)                                            # not produced yet
    for f in (1, 2, 3):
        ua = a.uniforms(frame=f)
        step.update(ua); step.render()
    step.blit()
    color, depth = step.read(T.PLANE_COLOR), step.read(T.PLANE_DEPTH)
    disp = step.read(T.PLANE_DISPLAY)
    ref, _ = oracle.blit(color)
    assert np.abs(disp.astype(np.int32) - ref.astype(np.int32)).max() <= DISPLAY_TOL_LSB
    assert disp[..., :3].std() > 10                                           # an actual image
    ub = b.uniforms(frame=4)
    step.rayproject(ua, ub)
    rep = step.read(T.PLANE_REPROJECTED)
    dst = oracle.rayproject(ub, ua, depth)
    want = np.zeros((H * W, 4), np.float32)
    src = np.arange(H * W)
    ok = dst != 0xFFFFFFFF
    order = np.argsort(src[ok], kind="stable")                                # ascending source index: the last write wins
    want[dst[ok][order]] = color.reshape(-1, 4)[src[ok][order]]
    assert np.array_equal(rep.reshape(-1, 4), want)
    assert ok.mean() > 0.5
    # a second reprojection only overwrites the targets it lands on (the reference never clears blit_storage_texture)
    step.rayproject(ua, ua)
    rep2 = step.read(T.PLANE_REPROJECTED).reshape(-1, 4)
    dst2 = oracle.rayproject(ua, ua, depth)
    hit2 = np.zeros(H * W, bool); hit2[dst2[dst2 != 0xFFFFFFFF]] = True
    assert np.array_equal(rep2[~hit2], want[~hit2])
    # blit of an arbitrary device image (the reprojected one) and async read of the 8-bit plane
    step.blit(step.device_ptr(T.PLANE_REPROJECTED))
    ref2, _ = oracle.blit(rep2.reshape(H, W, 4))
    assert np.abs(step.read(T.PLANE_DISPLAY).astype(np.int32) - ref2.astype(np.int32)).max() <= DISPLAY_TOL_LSB
    step.destroy()


def test_blit_against_an_independent_restatement(oracle, hostemu):
    """blit.wgsl:36-41 + color.wgsl:2-27 written again in numpy float64 from the WGSL (no code shared with oracle/ or csrc/),
    followed by the sRGB transfer function an *Srgb surface applies on store (IEC 61966-2-1) and round-to-nearest to 8 bits.
    The oracle and the product's host build agree with it within 1 LSB on random HDR pixels, greys around both knees
    (x = 2 F90 and p = 0.8 - F90) and the extremes."""
    F90 = 0.04
    rng = np.random.RandomState(17)
    rgb = np.concatenate([rng.rand(4000, 3) ** 3 * 6.0, np.repeat(np.linspace(0.0, 1.5, 400)[:, None], 3, 1),
                          np.array([[0.08, 0.08, 0.08], [0.76, 0.76, 0.76], [0.8, 0.1, 0.05], [50.0, 40.0, 30.0], [1e-4, 2e-4, 3e-4]])])
    img = np.ones((len(rgb), 1, 4), np.float32)
    img[:, 0, :3] = rgb
    c = img[:, 0, :3].astype(np.float64)
    k_s, k_d = 0.8 - F90, 0.15
    x = c.min(axis=1, keepdims=True)
    f = np.where(x <= 2.0 * F90, x - x * x / (4.0 * F90), F90)
    p = (c - f).max(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        p_n = 1.0 - (1.0 - k_s) ** 2 / (p + 1.0 - 2.0 * k_s)
        g = 1.0 / (k_d * (p - p_n) + 1.0)
        out = np.where(p <= k_s, c - f, (c - f) * (p_n / p) * g + p_n * (1.0 - g))
    out = np.clip(out, 0.0, 1.0)
    srgb = np.where(out <= 0.0031308, out * 12.92, 1.055 * np.power(out, 1.0 / 2.4) - 0.055)
    want = np.floor(srgb * 255.0 + 0.5).astype(np.int32)
    ref, _ = oracle.blit(img)
    got = hostemu.blit(img)
    assert np.abs(ref[:, 0, :3].astype(np.int32) - want).max() <= DISPLAY_TOL_LSB
    assert np.abs(got[:, 0, :3].astype(np.int32) - want).max() <= DISPLAY_TOL_LSB
    assert (want.max(axis=1) - want.min(axis=1)).max() > 100 and want.min() == 0 and want.max() >= 250       # the sample spans the range


def test_rayproject_against_an_independent_restatement(ptc, standins, oracle):
    """rayproject.wgsl:45-74 written again in numpy float64 from the WGSL.  Destination pixels are integers produced by
    round(): the fp32 oracle may land on the neighbouring pixel where the float64 position sits within 1e-3 of a half;
    everywhere else (> 99 % of the pixels) the two agree exactly, and never by more than one pixel."""
    W, H = 96, 64
    a, b = _cams(standins, ptc, W, H)
    rng = np.random.RandomState(4)
    depth = rng.uniform(0.5, 6.0, (H, W)).astype(np.float32)
    prev, cur = a.uniforms(frame=3), b.uniforms(frame=4)
    ref = oracle.rayproject(cur, prev, depth).reshape(H, W)

    def v3(x):
        return np.array([x[0], x[1], x[2]], np.float64)
    pc, cc = prev.camera, cur.camera
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    origin = v3(pc.first_pixel_pos) + xs[..., None] * v3(pc.pixel_space_x) + ys[..., None] * v3(pc.pixel_space_y)
    d = origin - v3(pc.position)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    to_old = origin + d * depth[..., None].astype(np.float64) - v3(cc.position)
    on_plane = to_old / (to_old @ (v3(cc.lookat) - v3(cc.position)))[..., None]
    psx, psy = v3(cc.pixel_space_x), v3(cc.pixel_space_y)
    rx = (on_plane @ (psx / np.linalg.norm(psx))) / np.linalg.norm(psx) + W * 0.5 + 0.5
    ry = (on_plane @ (psy / np.linalg.norm(psy))) / np.linalg.norm(psy) + H * 0.5 + 0.5
    qx, qy = np.floor(rx + 0.5), np.floor(ry + 0.5)                           # WGSL round(): halves are the uncertain cases below
    qx, qy = np.maximum(qx, 0.0), np.maximum(qy, 0.0)                         # u32() of a negative float clamps to 0 (WGSL value conversion):
    inside = (qx < W) & (qy < H)                                              # what leaves to the left / top lands in column / row 0
    hit = ref != 0xFFFFFFFF
    near_half = (np.abs(rx - np.floor(rx) - 0.5) < 1e-3) | (np.abs(ry - np.floor(ry) - 0.5) < 1e-3)
    near_edge = (np.abs(rx + 0.5 - W) < 1e-3) | (np.abs(ry + 0.5 - H) < 1e-3)
    sure = ~near_half & ~near_edge
    assert sure.mean() > 0.99
    assert np.array_equal(hit[sure], inside[sure])
    both = sure & hit
    assert both.mean() > 0.5
    assert np.array_equal((ref[both] % W).astype(np.float64), qx[both]) and np.array_equal((ref[both] // W).astype(np.float64), qy[both])
    rest = hit & inside & ~sure
    assert (np.abs((ref[rest] % W) - qx[rest]) <= 1).all() and (np.abs((ref[rest] // W) - qy[rest]) <= 1).all()
