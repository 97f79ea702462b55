"""Procedural stand-in scenes for the BASELINE.json configurations.

Every ``.obj`` under the reference's ``res/`` is a Git-LFS pointer and the GPU box has no copy of
the reference at all, so the benchmark / parity inputs are generated here, deterministically, with the triangle
counts, material parameters and texture counts of the named scenes:

* ``cornell``  — C1: room + two boxes, materials white/red/green/blue/light as ``res/cornell_box.mtl`` (no Ke:
  all light comes from the environment through the open front).  Goes through the OBJ/MTL loader.
* ``dragon``   — C2: one 871 200-triangle displaced sphere with the glass material of ``res/dragon.mtl``
  (Pr 0.01, Tf 1 1 1, Ni 1.5) over a ground quad.
* ``sponza``   — C3: ~66 k-triangle two-storey atrium, 20 materials with the parameters every material of
  ``res/sponza/sponza.mtl`` carries (Ni 1.5, Tf 1 1 1 -> transmission 1), 19 albedo + 18 bump
  textures (procedural, same role: grey height maps consumed as tangent-space normals).
* ``classroom``— C4: ~0.6 M triangles of desks/chairs, 79 constant-colour materials, three emitters.
* ``san_miguel`` — C5: ~10 M triangles (courtyard + 2 000 plants of 5 000 leaf triangles), 287 materials,
  textured, divergent shading.

The environment map (texture 9; the reference uses ``res/knob/envmap.exr``, 4096x2048 rgba32f, values
0.045..1.0) is a procedural sky of the same size and range.
"""
import os

import numpy as np

from . import _types as T
from .scene import Scene, Texture

F = np.float32


# --------------------------------------------------------------------------------------------------------
# textures
# --------------------------------------------------------------------------------------------------------
def make_envmap(width=4096, height=2048):
    """Sky gradient + sun disc + ground, float32 RGBA, range ~[0.045, 1.0] like res/knob/envmap.exr."""
    v = (np.arange(height, dtype=F) + F(0.5)) / F(height)
    u = (np.arange(width, dtype=F) + F(0.5)) / F(width)
    theta = v * F(np.pi)                      # v = acos(d.y)/pi (sky.wgsl:7)
    phi = u * F(2 * np.pi) - F(np.pi)         # u = (atan2(d.z, d.x) + pi) / 2pi
    dy = np.cos(theta)[:, None]
    st = np.sin(theta)[:, None]
    dx = st * np.cos(phi)[None, :]
    dz = st * np.sin(phi)[None, :]
    up = np.clip(dy, 0, 1)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.45 + 0.25 * (1 - up), 0.75 + 0.2 * (1 - up)], axis=-1).astype(F)
    ground = np.array([0.16, 0.14, 0.12], dtype=F)
    img = np.where((dy > 0)[..., None], sky * np.ones_like(dx)[..., None], ground[None, None, :] * np.ones_like(dx)[..., None])
    sun = np.array([0.45, 0.75, 0.48], dtype=F)
    sun /= np.linalg.norm(sun)
    c = dx * sun[0] + dy * sun[1] + dz * sun[2]
    img = img + (np.clip((c - 0.985) / 0.015, 0, 1) ** 2)[..., None].astype(F)
    img = np.clip(img, 0.045, 1.0).astype(F)
    out = np.ones((height, width, 4), dtype=F)
    out[..., :3] = img
    return Texture(out)


def _value_noise(w, h, cells, seed):
    rng = np.random.RandomState(seed)
    g = rng.rand(cells + 1, cells + 1).astype(F)
    g[-1, :] = g[0, :]
    g[:, -1] = g[:, 0]
    ys = np.linspace(0, cells, h, endpoint=False, dtype=F)
    xs = np.linspace(0, cells, w, endpoint=False, dtype=F)
    y0 = ys.astype(np.int32)
    x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_albedo_texture(w, h, seed, base):
    n = 0.55 * _value_noise(w, h, 8, seed) + 0.3 * _value_noise(w, h, 32, seed + 1) + 0.15 * _value_noise(w, h, 128, seed + 2)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    for k in range(3):
        img[..., k] = np.clip(255.0 * base[k] * (0.55 + 0.6 * n), 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return Texture(img)


def make_bump_texture(w, h, seed):
    """Greyscale height map (what Sponza's *-bump.jpg files are); the reference feeds it to the normal slot."""
    n = 0.6 * _value_noise(w, h, 16, seed) + 0.4 * _value_noise(w, h, 64, seed + 7)
    g = np.clip(255.0 * n, 0, 255).astype(np.uint8)
    img = np.stack([g, g, g, np.full_like(g, 255)], axis=-1)
    return Texture(img)


# --------------------------------------------------------------------------------------------------------
# geometry helpers (all float32, deterministic)
# --------------------------------------------------------------------------------------------------------
def grid(origin, du, dv, nu, nv, uv_scale=(1.0, 1.0), displace=None):
    """(nu x nv) quads spanning origin + s*du + t*dv; returns positions, normals, uvs, indices (2 tris / quad)."""
    s = np.linspace(0, 1, nu + 1, dtype=F)
    t = np.linspace(0, 1, nv + 1, dtype=F)
    S, Tt = np.meshgrid(s, t, indexing="xy")
    o, du, dv = np.asarray(origin, F), np.asarray(du, F), np.asarray(dv, F)
    P = o[None, None, :] + S[..., None] * du[None, None, :] + Tt[..., None] * dv[None, None, :]
    n = np.cross(du, dv)
    n = (n / np.linalg.norm(n)).astype(F)
    if displace is not None:
        P = P + displace(S, Tt)[..., None] * n[None, None, :]
    P = P.astype(F)
    # smooth normals from the grid
    dPs = np.gradient(P, axis=1)
    dPt = np.gradient(P, axis=0)
    N = np.cross(dPs, dPt)
    N = (N / np.maximum(np.linalg.norm(N, axis=-1, keepdims=True), 1e-20)).astype(F)
    UV = np.stack([S * F(uv_scale[0]), Tt * F(uv_scale[1])], axis=-1).astype(F)
    idx = _grid_indices(nu, nv)
    return P.reshape(-1, 3), N.reshape(-1, 3), UV.reshape(-1, 2), idx


def _grid_indices(nu, nv, wrap_u=False):
    cols = nu if wrap_u else nu + 1
    i = np.arange(nu, dtype=np.uint32)[None, :]
    j = np.arange(nv, dtype=np.uint32)[:, None]
    a = j * cols + i
    b = j * cols + (i + 1) % cols if wrap_u else a + 1
    c = a + cols
    d = b + cols
    tri = np.stack([a, b, d, a, d, c], axis=-1)
    return tri.reshape(-1).astype(np.uint32)


def box(lo, hi):
    lo, hi = np.asarray(lo, F), np.asarray(hi, F)
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    faces = [  # origin, du, dv with outward normal = du x dv
        ((x0, y0, z1), (x1 - x0, 0, 0), (0, y1 - y0, 0)),   # +z
        ((x1, y0, z0), (x0 - x1, 0, 0), (0, y1 - y0, 0)),   # -z
        ((x1, y0, z1), (0, 0, z0 - z1), (0, y1 - y0, 0)),   # +x
        ((x0, y0, z0), (0, 0, z1 - z0), (0, y1 - y0, 0)),   # -x
        ((x0, y1, z1), (x1 - x0, 0, 0), (0, 0, z0 - z1)),   # +y
        ((x0, y0, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0)),   # -y
    ]
    return merge([grid(o, du, dv, 1, 1) for o, du, dv in faces])


def merge(parts):
    P, N, UV, I = [], [], [], []
    base = 0
    for p, n, uv, idx in parts:
        P.append(p); N.append(n); UV.append(uv); I.append(idx + np.uint32(base))
        base += len(p)
    return np.concatenate(P), np.concatenate(N), np.concatenate(UV), np.concatenate(I)


def transform(part, rot_y=0.0, scale=1.0, translate=(0, 0, 0)):
    p, n, uv, idx = part
    c, s = F(np.cos(rot_y)), F(np.sin(rot_y))
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=F)
    return (p * F(scale)) @ R.T + np.asarray(translate, F), n @ R.T, uv, idx


def revolve(radius_fn, y0, y1, nseg, nring, center=(0, 0), uv_scale=(1.0, 1.0)):
    """Surface of revolution around the y axis (columns)."""
    t = np.linspace(0, 1, nring + 1, dtype=F)
    a = np.linspace(0, 2 * np.pi, nseg, endpoint=False, dtype=F)
    r = radius_fn(t).astype(F)
    y = (y0 + (y1 - y0) * t).astype(F)
    P = np.stack([center[0] + r[:, None] * np.cos(a)[None, :], np.broadcast_to(y[:, None], (nring + 1, nseg)),
                  center[1] + r[:, None] * np.sin(a)[None, :]], axis=-1).astype(F)
    N = np.stack([np.broadcast_to(np.cos(a)[None, :], (nring + 1, nseg)), np.zeros((nring + 1, nseg), F),
                  np.broadcast_to(np.sin(a)[None, :], (nring + 1, nseg))], axis=-1).astype(F)
    UV = np.stack([np.broadcast_to((a / F(2 * np.pi) * F(uv_scale[0]))[None, :], (nring + 1, nseg)),
                   np.broadcast_to((t * F(uv_scale[1]))[:, None], (nring + 1, nseg))], axis=-1).astype(F)
    idx = _grid_indices(nseg, nring, wrap_u=True)
    # winding: make the outward normal agree with the geometric one
    idx = idx.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
    return P.reshape(-1, 3), N.reshape(-1, 3), UV.reshape(-1, 2), idx.astype(np.uint32)


def sphere_displaced(nu, nv, radius, center, disp):
    """UV sphere with radial displacement disp(theta, phi); nu x nv x 2 triangles."""
    th = np.linspace(0, np.pi, nv + 1, dtype=np.float64)
    ph = np.linspace(0, 2 * np.pi, nu + 1, dtype=np.float64)
    PH, TH = np.meshgrid(ph, th, indexing="xy")
    r = radius * (1.0 + disp(TH, PH))
    P = np.stack([r * np.sin(TH) * np.cos(PH), r * np.cos(TH), r * np.sin(TH) * np.sin(PH)], axis=-1)
    dPu = np.gradient(P, axis=1)
    dPv = np.gradient(P, axis=0)
    N = np.cross(dPu, dPv)
    nn = np.linalg.norm(N, axis=-1, keepdims=True)
    radial = P / np.maximum(np.linalg.norm(P, axis=-1, keepdims=True), 1e-20)
    N = np.where(nn > 1e-12, N / np.maximum(nn, 1e-20), radial)
    N = np.where((np.sum(N * radial, axis=-1, keepdims=True) < 0), -N, N)
    P = P + np.asarray(center, np.float64)
    UV = np.stack([PH / (2 * np.pi), TH / np.pi], axis=-1)
    idx = _grid_indices(nu, nv).reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
    return P.reshape(-1, 3).astype(F), N.reshape(-1, 3).astype(F), UV.reshape(-1, 2).astype(F), idx.astype(np.uint32)


# --------------------------------------------------------------------------------------------------------
# MTL text -> materials through the product's loader
# --------------------------------------------------------------------------------------------------------
def _materials_from_mtl(scene, mtl_text, decode=None):
    """Run an MTL through the C++ loader (scene.rs:134-167 semantics: texture push order, constant de-dup).
    The OBJ carries no faces; tobj still emits its final (empty) model, which becomes an empty mesh 0."""
    names = [ln.split(None, 1)[1].strip() for ln in mtl_text.splitlines() if ln.strip().startswith("newmtl")]
    kw = {} if decode is None else {"decode": decode}
    scene.load_obj(obj_text="mtllib standin.mtl\n", mtl_text=mtl_text, obj_dir="standin", **kw)
    return {n: i + 1 for i, n in enumerate(names)}


CORNELL_MTL = """newmtl white
Ka 0 0 0
Kd 1 1 1
Ks 0 0 0

newmtl red
Ka 0 0 0
Kd 1 0 0
Ks 0 0 0

newmtl green
Ka 0 0 0
Kd 0 1 0
Ks 1 1 1

newmtl blue
Ka 0 0 0
Kd 0 0 1
Ks 0 0 0

newmtl light
Ka 6 6 6
Kd 1 1 1
Ks 0 0 0
"""


def _obj_text(groups, with_normals=True):
    """groups: list of (name, material, (P, N, UV, idx)) -> OBJ text (1-based global indices)."""
    out = ["mtllib standin.mtl"]
    base = 0
    for name, mat, (P, N, UV, idx) in groups:
        out.append(f"o {name}")
        out += ["v %.9g %.9g %.9g" % tuple(p) for p in P]
        out += ["vt %.9g %.9g" % (uv[0], uv[1]) for uv in UV]
        if with_normals:
            out += ["vn %.9g %.9g %.9g" % tuple(n) for n in N]
        out.append(f"usemtl {mat}")
        for a, b, c in idx.reshape(-1, 3).astype(np.int64) + base + 1:
            if with_normals:
                out.append(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}")
            else:
                out.append(f"f {a}/{a} {b}/{b} {c}/{c}")
        base += len(P)
    return "\n".join(out) + "\n"


def cornell(env=None, with_normals=True, fix_normals=False):
    """C1: Cornell box, 36 triangles, through the OBJ/MTL loader.  Open towards -x; camera looks along +x."""
    sc = Scene(env or make_envmap(512, 256))
    room = [
        ("floor", "white", grid((0, 0, -1), (0, 0, 2), (2, 0, 0), 1, 1)),
        ("ceiling", "white", grid((0, 2, -1), (2, 0, 0), (0, 0, 2), 1, 1)),
        ("back", "white", grid((2, 0, -1), (0, 2, 0), (0, 0, 2), 1, 1)),
        ("left", "red", grid((0, 0, -1), (2, 0, 0), (0, 2, 0), 1, 1)),
        ("right", "green", grid((0, 0, 1), (0, 2, 0), (2, 0, 0), 1, 1)),
        ("short_box", "white", transform(box((-0.3, 0, -0.3), (0.3, 0.6, 0.3)), rot_y=0.3, translate=(0.9, 0, 0.4))),
        ("tall_box", "blue", transform(box((-0.3, 0, -0.3), (0.3, 1.2, 0.3)), rot_y=-0.35, translate=(1.35, 0, -0.35))),
        ("light", "light", grid((0.8, 1.995, -0.25), (0.5, 0, 0), (0, 0, 0.5), 1, 1)),
    ]
    sc.load_obj(obj_text=_obj_text(room, with_normals), mtl_text=CORNELL_MTL, obj_dir="standin", fix_normals=fix_normals)
    cam = {"position": (-2.6, 1.0, 0.0), "fov": 60.0, "target": (1.0, 1.0, 0.0)}
    return sc, cam


DRAGON_MTL = """newmtl transmit
Kd 0.800000 0.800000 0.800000
Ks 0.500000 0.500000 0.500000
Ke 0.000000 0.000000 0.000000
Ni 1.500000
d 0.000000
illum 2
Pr 0.010000
Pm 0.000000
Ps 0.000000
Pc 0.000000
Pcr 0.030000
aniso 0.000000
anisor 0.000000
Tf 1.000000 1.000000 1.000000

newmtl ground
Kd 0.6 0.6 0.6
Ks 0 0 0
"""


def dragon(env=None, n=660):
    """C2: a single (n x n x 2)-triangle mesh (871 200 at n=660) of glass over a ground quad."""
    sc = Scene(env or make_envmap())
    mats = _materials_from_mtl(sc, DRAGON_MTL)

    def disp(th, ph):
        return (0.15 * np.sin(5 * th) * np.sin(4 * ph) + 0.07 * np.sin(13 * th + 1.0) * np.sin(11 * ph)
                + 0.03 * np.sin(29 * th) * np.sin(31 * ph + 2.0) + 0.012 * np.sin(61 * th + 0.5) * np.sin(67 * ph))

    P, N, UV, idx = sphere_displaced(n, n, 1.0, (0.0, 1.25, 0.0), disp)
    sc.add_mesh(P, N, UV, idx, mats["transmit"])
    g = grid((-6, 0, -6), (0, 0, 12), (12, 0, 0), 8, 8, uv_scale=(8, 8))
    sc.add_mesh(*g, mats["ground"])
    cam = {"position": (-3.2, 1.6, 0.4), "fov": 60.0, "target": (0.0, 1.2, 0.0)}
    return sc, cam


def _sponza_mtl(n_mats=20, n_albedo=19, n_bump=18):
    base = [(0.745098, 0.709804, 0.674510), (0.713726, 0.705882, 0.658824), (0.784314, 0.784314, 0.784314),
            (0.737255, 0.709804, 0.670588), (0.647059, 0.647059, 0.603922), (0.588235, 0.556863, 0.501961)]
    out = []
    for i in range(n_mats):
        kd = base[i % len(base)]
        out.append(f"newmtl sp_{i:02d}")
        out += ["\tNs 50.000000", "\tNi 1.500000", "\td 1.000000", "\tTr 0.000000", "\tTf 1.000000 1.000000 1.000000 ",
                "\tillum 2", "\tKa 0.000000 0.000000 0.000000", "\tKd %.6f %.6f %.6f" % kd, "\tKs 0.000000 0.000000 0.000000",
                "\tKe 0.000000 0.000000 0.000000"]
        if i < n_albedo:
            out += [f"\tmap_Ka albedo_{i:02d}.jpg", f"\tmap_Kd albedo_{i:02d}.jpg"]
        if i < n_bump:
            out += [f"\tmap_bump bump_{i:02d}.jpg", f"\tbump bump_{i:02d}.jpg"]
        out.append("")
    return "\n".join(out) + "\n"


# The reference's REAL Sponza material library and textures (byte copies of /root/reference/res/sponza, made by
# tests/golden/fetch_ref_res.py).  The .obj beside them is a Git-LFS pointer, so the geometry below stays procedural.
REF_RES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_res")


def sponza(env=None, tex_size=512, real_materials=None):
    """C3: ~78 k-triangle atrium with the reference's real sponza.mtl (20 materials, `Tf 1 1 1` on every one: SURVEY F10)
    and its real 19 albedo + 18 bump JPGs on the objects round-robin; real_materials=False (or a checkout without
    tests/golden/ref_res) uses a synthetic MTL of the same shape with procedural textures instead."""
    sc = Scene(env or make_envmap())
    base = [(0.745, 0.710, 0.675), (0.714, 0.706, 0.659), (0.784, 0.784, 0.784), (0.737, 0.710, 0.671), (0.647, 0.647, 0.604), (0.588, 0.557, 0.502)]
    real_dir = os.path.join(REF_RES, "sponza")
    if real_materials is None:
        real_materials = os.path.exists(os.path.join(real_dir, "sponza.mtl"))

    def decode(path):
        name = path.rsplit("/", 1)[-1]
        k = int(name.split("_")[1].split(".")[0])
        if name.startswith("albedo"):
            return make_albedo_texture(tex_size, tex_size, 100 + k, base[k % len(base)])
        return make_bump_texture(tex_size, tex_size, 500 + k)

    if real_materials:
        from .scene import decode_image
        mtl_text = open(os.path.join(real_dir, "sponza.mtl"), encoding="utf-8", errors="replace").read()
        mats = _materials_from_mtl(sc, mtl_text, decode=lambda p: decode_image(os.path.join(real_dir, os.path.basename(p))))
        ids = list(mats.values())                     # the 20 materials in file order
    else:
        mats = _materials_from_mtl(sc, _sponza_mtl(), decode=decode)
        ids = [mats[f"sp_{i:02d}"] for i in range(20)]
    L, Wd, Hh = 30.0, 12.0, 14.0          # atrium extents: x in [-15,15], z in [-6,6], y in [0,14]
    objs = []
    objs.append(grid((-L / 2, 0, -Wd / 2), (0, 0, Wd), (L, 0, 0), 96, 40, uv_scale=(4, 4)))                      # floor 7680
    objs.append(grid((-L / 2, Hh, -Wd / 2), (L, 0, 0), (0, 0, Wd), 48, 20, uv_scale=(4, 4)))                     # ceiling 1920
    objs.append(grid((-L / 2, 0, -Wd / 2), (L, 0, 0), (0, Hh, 0), 64, 32, uv_scale=(4, 4)))                      # wall -z 4096
    objs.append(grid((-L / 2, 0, Wd / 2), (0, Hh, 0), (L, 0, 0), 32, 64, uv_scale=(4, 4)))                       # wall +z 4096
    objs.append(grid((-L / 2, 0, -Wd / 2), (0, Hh, 0), (0, 0, Wd), 16, 24, uv_scale=(4, 4)))                     # wall -x 768
    objs.append(grid((L / 2, 0, -Wd / 2), (0, 0, Wd), (0, Hh, 0), 24, 16, uv_scale=(4, 4)))                      # wall +x 768
    # gallery floors (second storey) on both sides
    for zs in (-1, 1):
        z0 = zs * Wd / 2
        z1 = zs * (Wd / 2 - 2.5)
        a, b = (z0, z1) if zs < 0 else (z1, z0)
        objs.append(grid((-L / 2, 6.0, a), (0, 0, b - a), (L, 0, 0), 8, 64, uv_scale=(4, 1)))                    # 2 x 1024
        objs.append(grid((-L / 2, 5.7, a), (L, 0, 0), (0, 0, b - a), 64, 8, uv_scale=(4, 1)))                    # underside 2 x 1024
    # columns: 2 rows x 12 x 2 storeys
    def col_r(t):
        return 0.32 + 0.06 * np.cos(np.clip(t, 0, 1) * np.pi * 2) ** 8 + 0.02 * np.sin(t * 40.0)
    xs = np.linspace(-L / 2 + 1.8, L / 2 - 1.8, 12)
    for zs in (-1, 1):
        zc = zs * (Wd / 2 - 2.5)
        for x in xs:
            objs.append(revolve(col_r, 0.0, 5.7, 16, 10, center=(x, zc), uv_scale=(2, 4)))                       # 320 each
            objs.append(revolve(col_r, 6.0, 11.0, 16, 8, center=(x, zc), uv_scale=(2, 4)))                       # 256 each
    # arches between neighbouring columns: swept half-tubes
    def arch(x0, x1, y, z, nseg=20, nring=10, tube=0.18):
        t = np.linspace(0, np.pi, nseg + 1)
        a = np.linspace(0, 2 * np.pi, nring, endpoint=False)
        cx, r = 0.5 * (x0 + x1), 0.5 * (x1 - x0)
        C = np.stack([cx - r * np.cos(t), y + r * np.sin(t) * 0.8, np.full_like(t, z)], axis=-1)
        Tn = np.stack([np.sin(t), np.cos(t) * 0.8, np.zeros_like(t)], axis=-1)
        Tn /= np.linalg.norm(Tn, axis=-1, keepdims=True)
        B = np.array([0.0, 0.0, 1.0])[None, :].repeat(len(t), 0)
        Nn = np.cross(B, Tn)
        P = C[:, None, :] + tube * (np.cos(a)[None, :, None] * Nn[:, None, :] + np.sin(a)[None, :, None] * B[:, None, :])
        Nr = np.cos(a)[None, :, None] * Nn[:, None, :] + np.sin(a)[None, :, None] * B[:, None, :]
        UV = np.stack([np.broadcast_to((a / (2 * np.pi))[None, :], (nseg + 1, nring)), np.broadcast_to((t / np.pi * 2)[:, None], (nseg + 1, nring))], axis=-1)
        idx = _grid_indices(nring, nseg, wrap_u=True)
        return P.reshape(-1, 3).astype(F), Nr.reshape(-1, 3).astype(F), UV.reshape(-1, 2).astype(F), idx
    for zs in (-1, 1):
        zc = zs * (Wd / 2 - 2.5)
        for i in range(len(xs) - 1):
            objs.append(arch(xs[i], xs[i + 1], 4.4, zc))                                                          # 400 each x 22
            objs.append(arch(xs[i], xs[i + 1], 9.8, zc, nseg=16, nring=8))                                        # 256 each x 22
    # hanging curtains: wavy grids
    for k in range(6):
        x = -L / 2 + 4.0 + k * 4.4
        objs.append(grid((x, 11.5, -1.2), (0, -5.0, 0), (0, 0, 2.4), 40, 28, uv_scale=(2, 1),
                         displace=lambda S, Tt, k=k: 0.18 * np.sin(Tt * 18.0 + k) * (0.3 + S)))                   # 2240 each
    # two relief panels
    for zs in (-1, 1):
        objs.append(grid((-3.0, 1.0, zs * (Wd / 2 - 0.02)), (6.0, 0, 0) if zs > 0 else (0, 4.0, 0), (0, 4.0, 0) if zs > 0 else (6.0, 0, 0), 56, 56,
                         uv_scale=(1, 1), displace=lambda S, Tt: -0.12 * np.abs(np.sin(S * 9.0) * np.cos(Tt * 7.0)) - 0.02))  # 6272 each
    for i, o in enumerate(objs):
        sc.add_mesh(*o, ids[i % len(ids)])
    cam = {"position": (-12.5, 3.2, 0.6), "fov": 60.0, "target": (6.0, 4.5, -0.4)}
    return sc, cam


def furnace(env=None, ke=(0.75, 0.5, 0.25)):
    """Analytic fixture: a closed box whose walls all emit Ke.  Every path ends at its first hit with
    radiance exactly Ke (pathtracer/pathtracer.wgsl:200-206)."""
    sc = Scene(env or make_envmap(64, 32))
    mtl = "newmtl glow\nKd 0.5 0.5 0.5\nKe %g %g %g\n" % ke
    parts = box((-1, -1, -1), (1, 1, 1))
    sc.load_obj(obj_text=_obj_text([("room", "glow", parts)]), mtl_text=mtl, obj_dir="standin")
    return sc, {"position": (0.0, 0.0, 0.0), "fov": 60.0, "target": (1.0, 0.0, 0.0)}


LIT_ROOM_MTL = """newmtl wall
Kd 0.75 0.75 0.75
Pr 0.9

newmtl red
Kd 0.8 0.15 0.1
Pr 0.7

newmtl green
Kd 0.15 0.7 0.2
Pr 0.5

newmtl lamp
Kd 0.8 0.8 0.8
Ke 24 22 18

newmtl glow
Kd 0.5 0.5 0.5
Ke 2 3 5
"""


def lit_room(env=None):
    """Light-sampling fixture: a CLOSED room under a black sky, lit only by a small ceiling lamp and a dim glowing
    block; a box casts a shadow.  Whatever reaches the image came from an emissive triangle, so the two estimators
    (BSDF sampling alone / PTC_LIGHTS_MIS) can be compared on their converged means."""
    black = np.zeros((16, 32, 4), dtype=F); black[..., 3] = 1.0
    sc = Scene(env or Texture(black))
    parts = [
        ("floor", "wall", grid((0, 0, -1), (0, 0, 2), (2, 0, 0), 2, 2)),
        ("ceiling", "wall", grid((0, 2, -1), (2, 0, 0), (0, 0, 2), 2, 2)),
        ("back", "wall", grid((2, 0, -1), (0, 2, 0), (0, 0, 2), 2, 2)),
        ("front", "wall", grid((0, 0, -1), (0, 0, 2), (0, 2, 0), 2, 2)),
        ("left", "red", grid((0, 0, -1), (2, 0, 0), (0, 2, 0), 2, 2)),
        ("right", "green", grid((0, 0, 1), (0, 2, 0), (2, 0, 0), 2, 2)),
        ("blocker", "wall", transform(box((-0.3, 0, -0.3), (0.3, 0.9, 0.3)), rot_y=0.4, translate=(1.2, 0, 0.2))),
        ("lamp", "lamp", grid((0.9, 1.99, -0.2), (0.4, 0, 0), (0, 0, 0.4), 1, 1)),
        ("glow", "glow", transform(box((-0.1, 0, -0.1), (0.1, 0.2, 0.1)), translate=(0.6, 0.0, -0.6))),
    ]
    sc.load_obj(obj_text=_obj_text(parts), mtl_text=LIT_ROOM_MTL, obj_dir="standin")
    return sc, {"position": (0.15, 1.0, 0.0), "fov": 70.0, "target": (1.5, 0.8, 0.0)}


def empty(env=None):
    """No geometry: every pixel equals the sky lookup of its primary direction."""
    return Scene(env or make_envmap(256, 128)), {"position": (0.0, 0.0, 0.0), "fov": 60.0, "target": (1.0, 0.0, 0.0)}


def random_soup(n_tris, seed=1, env=None, textured=True, extent=4.0, tri_size=0.35):
    """Random triangle soup with a handful of materials (diffuse / glass / emissive / textured) — the
    parity workhorse: exercises every shading branch with arbitrary geometry."""
    rng = np.random.RandomState(seed)
    sc = Scene(env or make_envmap(256, 128))
    mtl = ("newmtl diffuse\nKd 0.8 0.7 0.6\n\nnewmtl glass\nKd 0.9 0.9 0.9\nNi 1.5\nPr 0.05\nTf 1 1 1\n\n"
           "newmtl metal\nKd 0.9 0.6 0.3\nPm 0.5\nPr 0.3\n\nnewmtl lamp\nKd 1 1 1\nKe 4 3 2\n\n"
           "newmtl rough_glass\nKd 0.7 0.9 0.8\nNi 1.33\nPr 0.6\nTf 0.5 0.5 0.5\n")
    if textured:
        mtl += "\nnewmtl tex\nKd 1 1 1\nmap_Kd albedo_00.png\nmap_bump bump_00.png\nmap_Pr rough_00.png\nNi 1.2\nTf 0.3 0.3 0.3\n"

    def decode(path):
        name = path.rsplit("/", 1)[-1]
        if name.startswith("albedo"):
            return make_albedo_texture(64, 32, 11, (0.9, 0.8, 0.7))
        if name.startswith("rough"):
            t = make_bump_texture(48, 48, 5)
            return Texture((t.texels.astype(np.float32) / np.float32(255.0)).astype(np.float32))   # a float texture
        return make_bump_texture(32, 64, 3)

    mats = _materials_from_mtl(sc, mtl, decode=decode)
    names = list(mats)
    per = max(1, n_tris // len(names))
    for k, name in enumerate(names):
        n = per if k < len(names) - 1 else n_tris - per * (len(names) - 1)
        if n <= 0:
            continue
        c = (rng.rand(n, 1, 3) * 2 - 1) * extent
        P = (c + (rng.rand(n, 3, 3) * 2 - 1) * tri_size).astype(F).reshape(-1, 3)
        N = rng.randn(n * 3, 3).astype(F)
        N /= np.linalg.norm(N, axis=-1, keepdims=True)
        UV = (rng.rand(n * 3, 2) * 3 - 1).astype(F)
        idx = np.arange(n * 3, dtype=np.uint32)
        sc.add_mesh(P, N, UV, idx, mats[name])
    return sc, {"position": (-extent * 1.6, 0.3, 0.2), "fov": 60.0, "target": (0.0, 0.0, 0.0)}


# --------------------------------------------------------------------------------------------------------
# C4 classroom / C5 san_miguel
# --------------------------------------------------------------------------------------------------------
def _classroom_mtl(n_mats=79):
    """79 constant-colour materials as res/classroom/classroom.mtl: no textures, no Tf, three emitters (Ke 5 / 7 / 15)."""
    rng = np.random.RandomState(0xC1A55)
    out = []
    for i in range(n_mats):
        kd = 0.15 + 0.75 * rng.rand(3)
        out.append(f"newmtl cl_{i:02d}")
        out += ["Ns %.6f" % (20 + 400 * rng.rand()), "Ka 1.000000 1.000000 1.000000", "Kd %.6f %.6f %.6f" % tuple(kd),
                "Ks %.6f %.6f %.6f" % ((0.5 * rng.rand(),) * 3), "Ni 1.450000", "d 1.000000", "illum 2"]
        if i in (10, 11, 12):
            out.append("Ke %d %d %d" % (((5, 7, 15)[i - 10],) * 3))
        else:
            out.append("Ke 0.000000 0.000000 0.000000")
        if i % 9 == 4:
            out += ["Pr %.6f" % (0.05 + 0.5 * rng.rand()), "Pm %.6f" % (0.2 + 0.6 * rng.rand())]
        out.append("")
    return "\n".join(out) + "\n"


def _ref_mtl(rel):
    p = os.path.join(REF_RES, rel)
    return open(p, encoding="utf-8", errors="replace").read() if os.path.exists(p) else None


def _emitter_names(mtl_text):
    """names of the materials whose `Ke` is not all zero (scene.rs:307-309: those are the lights)"""
    out, cur = [], None
    for ln in mtl_text.splitlines():
        w = ln.split()
        if w[:1] == ["newmtl"]:
            cur = ln.split(None, 1)[1].strip()
        elif w[:1] == ["Ke"] and any(float(x) != 0.0 for x in w[1:4]):
            out.append(cur)
    return out


def classroom(env=None, detail=1.32, real_materials=None):
    """C4: a classroom of ~0.6 M triangles (detail=1): shell, 30 desks + chairs, shelves of books, blinds, a globe, ceiling
    light panels (the emitters), the 79 constant-colour materials of the reference's real res/classroom/classroom.mtl
    (no Tf at all: no transmission; three emitters, Ke 5 / 7 / 15, which the light panels get);
    real_materials=False (or no tests/golden/ref_res) uses a synthetic MTL of the same shape."""
    sc = Scene(env or make_envmap())
    real = _ref_mtl("classroom/classroom.mtl") if real_materials in (None, True) else None
    if real is not None:
        mats = _materials_from_mtl(sc, real)
        emit = _emitter_names(real)
        assert len(mats) == 79 and len(emit) == 3
        rest = [i for n, i in mats.items() if n not in emit]
        mid = rest[:10] + [mats[n] for n in emit] + rest[10:]       # slots 10..12 are what the light panels below use
    else:
        mats = _materials_from_mtl(sc, _classroom_mtl())
        mid = [mats[f"cl_{i:02d}"] for i in range(79)]
    k = lambda n: max(1, int(round(n * np.sqrt(detail))))
    X, Y, Z = 12.0, 3.2, 9.0
    parts = []
    add = lambda geo, m: parts.append((*geo, mid[m]))
    add(grid((0, 0, 0), (0, 0, Z), (X, 0, 0), k(150), k(200), uv_scale=(6, 8)), 0)                        # floor
    add(grid((0, Y, 0), (X, 0, 0), (0, 0, Z), k(100), k(75), uv_scale=(6, 8)), 1)                         # ceiling
    add(grid((0, 0, 0), (X, 0, 0), (0, Y, 0), k(120), k(32)), 2)                                          # walls
    add(grid((0, 0, Z), (0, Y, 0), (X, 0, 0), k(32), k(120)), 2)
    add(grid((0, 0, 0), (0, Y, 0), (0, 0, Z), k(32), k(90)), 3)
    add(grid((X, 0, 0), (0, 0, Z), (0, Y, 0), k(90), k(32)), 3)
    for i in range(6):                                                                                     # light panels (emitters 10..12)
        x, z = 2.0 + (i % 3) * 4.0, 2.5 + (i // 3) * 4.0
        add(grid((x - 0.6, Y - 0.02, z - 0.3), (1.2, 0, 0), (0, 0, 0.6), 4, 2), 10 + i % 3)
    leg = lambda cx, cz, h: revolve(lambda t: 0.02 + 0.004 * np.sin(t * 30), 0.0, h, k(12), k(10), center=(cx, cz))
    m = 13
    for r in range(5):                                                                                     # desks and chairs
        for c in range(6):
            x, z = 1.6 + c * 1.7, 1.5 + r * 1.45
            add(grid((x - 0.6, 0.74, z - 0.3), (0, 0, 0.6), (1.2, 0, 0), k(14), k(28),
                     displace=lambda S, Tt: 0.004 * np.sin(S * 40) * np.sin(Tt * 40)), m % 60 + 13)
            add(grid((x - 0.6, 0.71, z - 0.3), (1.2, 0, 0), (0, 0, 0.6), k(28), k(14)), m % 60 + 13)
            for dx, dz in ((-0.55, -0.25), (0.55, -0.25), (-0.55, 0.25), (0.55, 0.25)):
                add(leg(x + dx, z + dz, 0.71), (m + 1) % 60 + 13)
            cz = z + 0.55
            add(grid((x - 0.22, 0.45, cz - 0.2), (0, 0, 0.4), (0.44, 0, 0), k(26), k(26),
                     displace=lambda S, Tt: -0.03 * np.sin(S * np.pi) * np.sin(Tt * np.pi)), (m + 2) % 60 + 13)
            add(grid((x - 0.22, 0.5, cz + 0.2), (0.44, 0, 0), (0, 0.45, 0), k(26), k(26),
                     displace=lambda S, Tt: 0.04 * np.sin(S * np.pi)), (m + 3) % 60 + 13)
            for dx, dz in ((-0.2, -0.18), (0.2, -0.18), (-0.2, 0.18), (0.2, 0.18)):
                add(leg(x + dx, cz + dz, 0.45), (m + 4) % 60 + 13)
            m += 5
    rng = np.random.RandomState(0xB00C)
    for sh in range(3):                                                                                    # shelves of books along the +x wall
        for row in range(4):
            z = 0.6
            while z < Z - 0.8:
                w, h, dpt = 0.02 + 0.05 * rng.rand(), 0.18 + 0.1 * rng.rand(), 0.15 + 0.06 * rng.rand()
                y0 = 0.3 + row * 0.45
                add(box((X - 0.05 - dpt - sh * 0.0, y0, z), (X - 0.05, y0 + h, z + w)), int(rng.randint(13, 79)))
                z += w + 0.004
    for i in range(4):                                                                                     # window blinds on the z = 0 wall
        add(grid((1.2 + i * 2.6, 1.0, 0.03), (2.0, 0, 0), (0, 1.8, 0), k(100), k(60),
                 displace=lambda S, Tt: 0.015 * np.sin(Tt * 150.0)), 5 + i % 4)
    def gdisp(th, ph):
        return 0.01 * np.sin(9 * th) * np.sin(7 * ph) + 0.004 * np.sin(41 * th) * np.sin(37 * ph)
    add(sphere_displaced(k(200), k(200), 0.25, (X - 1.2, 1.1, Z - 1.0), gdisp), 9)                        # globe
    add(grid((3.0, 0.9, Z - 0.03), (0, 1.4, 0), (6.0, 0, 0), k(80), k(160), displace=lambda S, Tt: -0.003 * (np.sin(S * 60) * np.sin(Tt * 90) > 0.6)), 4)  # board
    sc.add_meshes(parts)
    cam = {"position": (0.6, 1.5, 0.7), "fov": 60.0, "target": (8.0, 1.0, 6.0)}
    return sc, cam


def _san_miguel_mtl(n_mats=287, n_kd=200, n_bump=100, n_pr=23):
    """287 materials / 323 texture files as res/san_miguel: map_Kd on the first 200, map_bump on 100, map_Pr on 23;
    five carry Tf (glass)."""
    rng = np.random.RandomState(0x5A17)
    out = []
    for i in range(n_mats):
        out.append(f"newmtl sm_{i:03d}")
        out += ["Ns %.6f" % (10 + 90 * rng.rand()), "Ni 1.500000", "d 1.000000", "illum 2",
                "Kd %.6f %.6f %.6f" % tuple(0.2 + 0.7 * rng.rand(3)), "Ks 0.000000 0.000000 0.000000", "Ke 0.000000 0.000000 0.000000"]
        if i < n_kd:
            out.append(f"map_Kd kd_{i:03d}.png")
        if 50 <= i < 50 + n_bump:
            out.append(f"map_bump bump_{i:03d}.png")
        if 120 <= i < 120 + n_pr:
            out.append(f"map_Pr pr_{i:03d}.png")
        if i in (240, 250, 260, 270, 280):
            out += ["Tf 1.000000 1.000000 1.000000", "Pr 0.020000"]
        out.append("")
    return "\n".join(out) + "\n"


def san_miguel(env=None, n_plants=2000, leaves_per_plant=2500, tex_size=256, real_materials=None):
    """C5: courtyard shell + n_plants procedural plants of leaves_per_plant leaf quads (2 triangles each; 10 M triangles
    at the defaults), the 287 materials of the reference's real res/san_miguel/san-miguel.mtl (264 map_Kd + 57 map_Bump +
    3 map_Ks files, Pr / Pm / Ps scalars on every material, 5 with Tf) assigned by hash => divergent shading.  The 323 PNGs
    themselves (134 MB) are not vendored: each file name gets a procedural texture (N_* = bump map).
    real_materials=False (or no tests/golden/ref_res) uses a synthetic MTL of the same shape."""
    sc = Scene(env or make_envmap())
    rng = np.random.RandomState(0x5A17)
    real = _ref_mtl("san_miguel/san-miguel.mtl") if real_materials in (None, True) else None

    def decode(path):
        name = path.rsplit("/", 1)[-1]
        kseed = int(name.split("_")[1].split(".")[0])
        if name.startswith("kd"):
            r = np.random.RandomState(kseed)
            return make_albedo_texture(tex_size, tex_size, 1000 + kseed, tuple(0.3 + 0.7 * r.rand(3)))
        if name.startswith("bump"):
            return make_bump_texture(tex_size, tex_size, 2000 + kseed)
        t = make_bump_texture(tex_size // 2, tex_size // 2, 3000 + kseed)
        return t

    def decode_real(path):
        import zlib
        name = path.rsplit("/", 1)[-1]
        kseed = zlib.crc32(name.encode()) % 100000
        if name.startswith("N_") or "bump" in name.lower():
            return make_bump_texture(tex_size, tex_size, 2000 + kseed)
        r = np.random.RandomState(kseed)
        return make_albedo_texture(tex_size, tex_size, 1000 + kseed, tuple(0.3 + 0.7 * r.rand(3)))

    if real is not None:
        mats = _materials_from_mtl(sc, real, decode=decode_real)
        assert len(mats) == 287
        mid = np.array(list(mats.values()), dtype=np.uint32)
    else:
        mats = _materials_from_mtl(sc, _san_miguel_mtl(), decode=decode)
        mid = np.array([mats[f"sm_{i:03d}"] for i in range(287)], dtype=np.uint32)
    Lx, Lz, Hy = 40.0, 30.0, 12.0
    parts = []
    sub = 400 if n_plants >= 500 else 120
    parts.append((*grid((-Lx / 2, 0, -Lz / 2), (0, 0, Lz), (Lx, 0, 0), sub, sub, uv_scale=(10, 10),
                        displace=lambda S, Tt: 0.03 * np.sin(S * 60) * np.sin(Tt * 45)), mid[0]))             # ground
    wsub = sub // 2
    parts.append((*grid((-Lx / 2, 0, -Lz / 2), (Lx, 0, 0), (0, Hy, 0), wsub, wsub // 2, uv_scale=(8, 3)), mid[1]))
    parts.append((*grid((-Lx / 2, 0, Lz / 2), (0, Hy, 0), (Lx, 0, 0), wsub // 2, wsub, uv_scale=(3, 8)), mid[2]))
    parts.append((*grid((-Lx / 2, 0, -Lz / 2), (0, Hy, 0), (0, 0, Lz), wsub // 2, wsub, uv_scale=(3, 8)), mid[3]))
    parts.append((*grid((Lx / 2, 0, -Lz / 2), (0, 0, Lz), (0, Hy, 0), wsub, wsub // 2, uv_scale=(8, 3)), mid[4]))
    # arcade columns around the courtyard
    ncol = 0
    for x in np.linspace(-Lx / 2 + 3, Lx / 2 - 3, 14):
        for z in (-Lz / 2 + 3.5, Lz / 2 - 3.5):
            parts.append((*revolve(lambda t: 0.28 + 0.05 * np.cos(t * np.pi * 2) ** 6, 0.0, 5.0, 24, 24, center=(x, z), uv_scale=(2, 4)), mid[5 + ncol % 20]))
            ncol += 1
    # plants: leaf quads in an ellipsoidal crown above a stem
    side = int(np.ceil(np.sqrt(n_plants)))
    for pi in range(n_plants):
        gx, gz = pi % side, pi // side
        cx = -Lx / 2 + 5 + (gx + rng.rand()) * (Lx - 10) / side
        cz = -Lz / 2 + 5 + (gz + rng.rand()) * (Lz - 10) / side
        hgt = 1.0 + 2.5 * rng.rand()
        n = leaves_per_plant
        u = rng.randn(n, 3).astype(F)
        u /= np.maximum(np.linalg.norm(u, axis=1, keepdims=True), 1e-6)
        rad = (rng.rand(n, 1).astype(F)) ** F(1 / 3)
        c = np.array([cx, hgt, cz], F) + u * rad * np.array([0.7, 0.9, 0.7], F)
        a = rng.randn(n, 3).astype(F); a /= np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-6)
        b = np.cross(a, rng.randn(n, 3).astype(F)); b /= np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-6)
        sz = F(0.03) + F(0.05) * rng.rand(n, 1).astype(F)
        a *= sz; b *= sz * F(0.6)
        P = np.stack([c - a - b, c + a - b, c + a + b, c - a + b], axis=1).reshape(-1, 3)
        nrm = np.cross(a, b); nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-12)
        N = np.repeat(nrm, 4, axis=0)
        UV = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F), (n, 1))
        q = (np.arange(n, dtype=np.uint32) * 4)[:, None]
        idx = (q + np.array([0, 1, 2, 0, 2, 3], np.uint32)[None, :]).reshape(-1)
        # 3 meshes per plant with hashed materials (divergent shading inside one crown)
        cuts = [0, n // 3, 2 * n // 3, n]
        for j in range(3):
            lo, hi = cuts[j], cuts[j + 1]
            mat = mid[(pi * 2654435761 + j * 40503) % 287]
            parts.append((P[lo * 4:hi * 4], N[lo * 4:hi * 4], UV[lo * 4:hi * 4], idx[lo * 6:hi * 6] - np.uint32(lo * 4), mat))
        parts.append((*revolve(lambda t: 0.06 * (1 - 0.5 * t), 0.0, hgt, 8, 6, center=(cx, cz)), mid[30 + pi % 10]))
    sc.add_meshes(parts)
    cam = {"position": (-Lx / 2 + 2.0, 2.2, -Lz / 2 + 2.5), "fov": 60.0, "target": (4.0, 1.8, 3.0)}
    return sc, cam


BUILDERS = {"cornell": cornell, "dragon": dragon, "sponza": sponza, "furnace": furnace, "empty": empty, "lit_room": lit_room,
            "classroom": classroom, "san_miguel": san_miguel}


def make_camera(cam, width, height):
    """Camera::new + steering to the stand-in's pose; returns a ``Camera`` in its steady state."""
    from .camera import Camera
    c = Camera((width, height), cam["fov"], cam["position"])
    c.look_at(cam["target"])
    return c
