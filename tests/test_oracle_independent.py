"""A SECOND restatement of the reference's path — written from the WGSL, sharing no code with oracle/ or csrc/.

oracle/pt_oracle.cpp and the CUDA kernels have one author and one numerics header (include/ptmath.h): a shared misreading
of a WGSL builtin, of the RNG draw order or of the accumulation rule would pass every oracle-vs-kernel test.  This file
restates `main -> pixel_color -> ray_color -> trace_ray / sample_material -> scatter -> disney_bsdf -> sky_color`
(src/shaders/pathtracer/pathtracer.wgsl:21-43, pathtracer/pathtracer.wgsl:76-101,149-278, pathtracer/common/rt_helpers.wgsl:
17-80, common/bsdf.wgsl:39-231, common/sky.wgsl:1-12, common/random.wgsl:7-36) in plain numpy float32 with numpy's own
sin / cos / arctan / arccos / arctan2, a float64 brute-force ray/triangle test and its own bilinear lookup, and checks the
C++ oracle against it pixel by pixel: the same radiance within a stated tolerance AND the same RNG state at the end of the
path (i.e. the same number of random draws in the same order), and the same first-hit depth / normal-mapped normal / albedo
planes (pathtracer/pathtracer.wgsl:208-223, 270-273; a first hit on an emitter leaves them zero).  One fixture is textured:
file-backed RGBA8 albedo and normal maps and a float roughness map, uv from -0.6 to 2.7 (Repeat on both sides of zero).
The reference's dormant light branch (SEND_TO_LIGHT_PROBABILITY > 0) is restated too and compared at 0.5.

Tolerances: fixtures whose radiance is a smooth function of the random numbers must agree to 1e-5 (relative to the pixel's
magnitude) on EVERY pixel; the multi-bounce diffuse fixture is chaotic (an ulp in a transcendental moves a later hit across
a triangle edge), so there >= 95 % of the pixels must end in the identical RNG state and those must agree to 2e-4.
"""
import numpy as np
import pytest

F = np.float32
PI = F(3.14159265359)                      # helpers.wgsl:4
U32 = 0xFFFFFFFF


# ---- random.wgsl:7-36 -------------------------------------------------------------------------------------------------
def jenkins(x):
    x = (x + (x << 10)) & U32
    x ^= x >> 6
    x = (x + (x << 3)) & U32
    x ^= x >> 11
    x = (x + (x << 15)) & U32
    return x


class Rng:
    def __init__(self, x, y, width, frame):
        seed = (((x * 1 + y * width) & U32) * 1) & U32                 # dot(pixel.xy, (1, res.x)) * pixel.z, pixel.z = 1
        self.state = jenkins(seed ^ jenkins(frame))
        self.draws = 0

    def int(self):
        self.state = (self.state * 747796405 + 2891336453) & U32
        s = self.state
        word = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & U32
        self.draws += 1
        return (word >> 22) ^ word

    def float(self):
        return F(F(self.int()) / F(0xFFFFFFFF))                        # f32(u) / f32(0xffffffff); f32(0xffffffff) = 2^32


# ---- WGSL builtins on float32 vectors ---------------------------------------------------------------------------------
def v(*a):
    return np.array(a, dtype=F)


def dot(a, b):
    return F(np.sum(a * b, dtype=F))


def normalize(a):
    return (a / F(np.sqrt(dot(a, a)))).astype(F)


def cross(a, b):
    return v(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def face_forward(e1, e2, e3):
    return e1 if dot(e2, e3) < 0 else -e1


def reflect(e1, e2):
    return (e1 - F(2.0) * dot(e2, e1) * e2).astype(F)


def refract(e1, e2, e3):
    d = dot(e2, e1)
    k = F(1.0) - e3 * e3 * (F(1.0) - d * d)
    if k < 0:
        return v(0, 0, 0)
    return (e3 * e1 - (e3 * d + F(np.sqrt(k))) * e2).astype(F)


def sign(x):
    return F(1.0) if x > 0 else (F(-1.0) if x < 0 else F(0.0))


def mix(a, b, t):
    return F(a * (F(1.0) - t) + b * t)


# ---- scene access ------------------------------------------------------------------------------------------------------
class World:
    def __init__(self, scene, max_bounces, p_light=0.0, spp=1):
        self.sc, self.max_bounces, self.p_light, self.spp = scene, max_bounces, F(p_light), spp
        idx = scene.indices.reshape(-1, 3)
        P = scene.vertices["position"].astype(np.float64)
        self.v0, self.e1, self.e2 = P[idx[:, 0]], P[idx[:, 1]] - P[idx[:, 0]], P[idx[:, 2]] - P[idx[:, 0]]
        self.tri_mesh = np.zeros(len(idx), np.int64)
        for m, mesh in enumerate(scene.meshes):
            self.tri_mesh[mesh["index_offset"] // 3:(mesh["index_offset"] + mesh["index_count"]) // 3] = m
        # pt.rs:206-243: emissive triangles in global order (emissive_texture_idx != 4), one all-zero dummy when there is none
        Pf = scene.vertices["position"].astype(F)
        self.lights = [(Pf[i0], Pf[i1], Pf[i2]) for t, (i0, i1, i2) in enumerate(idx)
                       if scene.materials[scene.meshes[self.tri_mesh[t]]["material_index"]]["emissive_texture_idx"] != 4]
        if not self.lights:
            self.lights = [(v(0, 0, 0), v(0, 0, 0), v(0, 0, 0))]

    def closest(self, ro, rd):
        """rayQuery closest hit, t in (0.01, 10000), two-sided, no culling — float64 Moeller-Trumbore over every triangle."""
        o, d = ro.astype(np.float64), rd.astype(np.float64)
        p = np.cross(d, self.e2)
        det = np.einsum("ij,ij->i", self.e1, p)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = o - self.v0
            u = np.einsum("ij,ij->i", s, p) * inv
            q = np.cross(s, self.e1)
            w = np.einsum("ij,ij->i", q, d[None, :].repeat(len(q), 0)) * inv
            t = np.einsum("ij,ij->i", self.e2, q) * inv
        ok = (det != 0) & (u >= 0) & (u <= 1) & (w >= 0) & (u + w <= 1) & (t > 0.01) & (t < 10000.0)
        if not ok.any():
            return None
        t = np.where(ok, t, np.inf)
        k = int(np.argmin(t))
        return k, F(t[k]), F(u[k]), F(w[k])

    def texture(self, idx, uv):
        """textureSampleLevel(.., uv, 0): Repeat addressing, bilinear (texel centres at (i + 0.5) / n), in float64."""
        tex = self.sc.textures[int(idx)].texels
        h, w = tex.shape[:2]
        tx = tex.astype(np.float64) / (255.0 if tex.dtype == np.uint8 else 1.0)
        if w == 1 and h == 1:
            return tx[0, 0].astype(F)
        x, y = float(uv[0]) * w - 0.5, float(uv[1]) * h - 0.5
        x0, y0 = int(np.floor(x)), int(np.floor(y))
        fx, fy = x - x0, y - y0
        c = lambda i, j: tx[j % h, i % w]
        top = c(x0, y0) * (1 - fx) + c(x0 + 1, y0) * fx
        bot = c(x0, y0 + 1) * (1 - fx) + c(x0 + 1, y0 + 1) * fx
        return (top * (1 - fy) + bot * fy).astype(F)


def sky_color(world, rd):                                              # sky.wgsl:1-12
    d = normalize(rd)
    vv = F(np.arccos(np.clip(d[1], F(-1), F(1)))) / PI
    uu = (F(np.arctan2(d[2], d[0])) + PI) / (F(2.0) * PI)
    return world.texture(9, (uu, vv))[:3]


def sample_material(world, mat, uv):                                   # rt_helpers.wgsl:17-42 (the taps the path reads)
    t = world.texture
    return dict(albedo=t(mat["albedo_texture_idx"], uv)[:3], tsn=(t(mat["normal_texture_idx"], uv)[:3] * F(2.0) - F(1.0)).astype(F),
                roughness=t(mat["roughness_texture_idx"], uv)[0], metallic=t(mat["metallic_texture_idx"], uv)[0],
                emissive=t(mat["emissive_texture_idx"], uv)[:3], transmission=t(mat["transmission_texture_idx"], uv)[0],
                ior=t(mat["ior_texture_idx"], uv)[0], anisotropy=F(mat["anisotropy"]))


def disney_bsdf(rng, p, ray_in, normal, m):                            # bsdf.wgsl:195-231 with :39-87, :155-187, :120-147 inlined
    vdir = (-ray_in).astype(F)
    raw_n = normal
    n = face_forward(raw_n, ray_in, raw_n)
    tangent = normalize(cross(n, vdir))
    bitangent = normalize(cross(n, tangent))
    n = normalize((tangent * m["tsn"][0] + bitangent * m["tsn"][1] + n * m["tsn"][2]).astype(F))
    aspect = F(np.sqrt(F(1.0) - F(0.9) * m["anisotropy"]))
    alpha = max(F(0.0001), F(m["roughness"] / aspect))
    diffuse = (m["albedo"] * (F(1.0) - m["metallic"])).astype(F)
    ndotv = dot(n, vdir)
    rng.float()                                                        # :203 `let u = rand_float();` — drawn, never used
    u1 = rng.float()                                                   # :158
    with np.errstate(divide="ignore"):
        up_down = F(np.arctan(alpha * F(np.sqrt(u1 / (F(1.0) - u1)))))
    left_right = rng.float() * F(2.0) * PI                             # :162
    lx, ly, lz = F(np.sin(up_down)) * F(np.cos(left_right)), F(np.cos(up_down)), F(np.sin(up_down)) * F(np.sin(left_right))
    h = normalize((lx * bitangent + ly * n + lz * tangent).astype(F))
    d = dot(vdir, raw_n)                                               # :124, the RAW normal
    eta = m["ior"] * F(d <= 0) + (F(1.0) / m["ior"]) * F(d > 0)
    r = refract(-vdir, h, eta)
    r0 = (F(1.0) - eta) / (F(1.0) + eta)
    r0sq = r0 * r0
    w = F(1.0) - dot(h, vdir)
    fres = r0sq + (F(1.0) - r0sq) * (w * w * w * w * w)
    u = rng.float()                                                    # :133
    tf = m["transmission"] * (F(1.0) - fres)
    refracted = F(tf >= u) * abs(sign(dot(r, v(1, 1, 1))))             # step(u, tf) * |sign(r.x + r.y + r.z)|
    out = (r * refracted + reflect(-vdir, h) * (F(1.0) - refracted)).astype(F)
    origin = (p + n * mix(F(0.01) * sign(ndotv), F(0.0), refracted)).astype(F)
    return diffuse, F(1.0), out, origin                                # colour, pdf (= 1 on both arms), direction, origin


def ray_color(world, rng, ro, rd):                                     # pathtracer/pathtracer.wgsl:149-238 (SEND_TO_LIGHT_PROBABILITY = 0)
    sc = world.sc
    mask, acc = v(1, 1, 1), v(0, 0, 0)
    rays = 0
    aux = (F(0.0), v(0, 0, 0), v(0, 0, 0))                             # first_t, first_normal, first_albedo (:240-256 defaults)
    for i in range(world.max_bounces + 1):
        rays += 1
        hit = world.closest(ro, rd)
        if hit is None:
            return (acc + sky_color(world, rd) * mask).astype(F), rays, aux  # :229-232
        tri, t, b1, b2 = hit
        w0 = F(1.0) - b1 - b2
        vtx = [sc.vertices[sc.indices[tri * 3 + k]] for k in range(3)]
        uv = (vtx[0]["uv"] * w0 + vtx[1]["uv"] * b1 + vtx[2]["uv"] * b2).astype(F)
        mat = sample_material(world, sc.materials[sc.meshes[world.tri_mesh[tri]]["material_index"]], uv)
        hit_light = dot(mat["emissive"], v(1, 1, 1)) > 0
        p = (rd * t + ro).astype(F)                                    # fma(rd, t, ro)
        n = normalize((vtx[0]["normal"] * w0 + vtx[1]["normal"] * b1 + vtx[2]["normal"] * b2).astype(F))
        to_light = None
        if not (rng.float() < F(1.0) - world.p_light):                 # :170 (always consumed) / :175-191, the dormant light branch
            l0, l1, l2 = world.lights[rng.int() % len(world.lights)]   # :176
            r1 = rng.float(); r2 = rng.float()                         # random_point_on_triangle, helpers.wgsl:7-18
            s1 = F(np.sqrt(r1))
            to_point = (l0 * (F(1.0) - s1) + l1 * s1 * (F(1.0) - r2) + l2 * s1 * r2).astype(F)
            ln = vtx[0]["normal"].astype(F)                            # light_scatter :59 — vertex 0's normal, not the interpolated one
            to_p_dir = normalize((to_point - p).astype(F))
            # triangle_pdf :22 traces a probe ray here whatever should_hit says; its result (`importance`) is never applied
            # to the throughput and it draws no random number, so oracle and kernels omit it — it is not counted here either
            if dot(to_p_dir, ln) > 0:
                to_light = (to_p_dir, ln)
        if to_light is not None:                                       # origin p without offset, attenuation = albedo, pdf unused
            direction, sn_raw = to_light
            colour, pdf, origin = mat["albedo"], F(1.0), p
        else:
            colour, pdf, direction, origin = disney_bsdf(rng, p, rd, n, mat)
            sn_raw = None
        prev_rd = rd
        ro, rd = origin, direction
        acc = (acc + mat["emissive"] * mask).astype(F)                 # :200
        mask = (mask * (colour / pdf)).astype(F)                       # :202
        if hit_light:                                                  # :204-206 — BEFORE the aux block: a first hit on a light leaves the planes 0
            break
        if i == 0:                                                     # :208-223, with scatter's n = faceForward(n, rd, n) (:91)
            sn = face_forward(n, prev_rd, n) if sn_raw is None else sn_raw
            tangent = normalize(cross(sn, (-prev_rd).astype(F)))
            bitangent = normalize(cross(sn, tangent))
            tsn = mat["tsn"]
            aux = (t, normalize((tangent * tsn[0] + bitangent * tsn[1] + sn * tsn[2]).astype(F)), mat["albedo"])
    return acc, rays, aux


def pixel(world, un, x, y):                                            # pathtracer.wgsl:27 + pathtracer/pathtracer.wgsl:240-278, SPP = 1
    cam = un.camera
    rng = Rng(x, y, cam.width, cam.frame)
    pos, fp = v(*cam.position), v(*cam.first_pixel_pos)
    psx, psy = v(*cam.pixel_space_x), v(*cam.pixel_space_y)
    jx = F(x) + F(0.5) * rng.float()
    jy = F(y) + F(0.5) * rng.float()
    rd = normalize(((fp + psx * jx + psy * jy) - pos).astype(F))
    col, rays, aux = v(0, 0, 0), 0, None
    for _ in range(world.spp):                                         # :258-266: the SAME rd for every sample, planes of the last one
        c1, r1, aux = ray_color(world, rng, pos, rd)
        col, rays = (col + c1).astype(F), rays + r1
    return (col / F(world.spp)).astype(F), rng, rays, aux


# ---- fixtures ------------------------------------------------------------------------------------------------------------
def quad(p0, du, dv):
    p0, du, dv = np.array(p0, F), np.array(du, F), np.array(dv, F)
    P = np.stack([p0, p0 + du, p0 + du + dv, p0 + dv])
    n = np.cross(du, dv); n = n / np.linalg.norm(n)
    return P, np.tile(n.astype(F), (4, 1)), np.array([[0, 0], [1, 0], [1, 1], [0, 1]], F), np.array([0, 1, 2, 0, 2, 3], np.uint32)


def scene_from(ptc, standins, mtl, objects, decode=None):
    sc = ptc.Scene(standins.make_envmap(256, 128))
    names = [l.split()[1] for l in mtl.splitlines() if l.startswith("newmtl")]
    if decode is None:
        sc.load_obj(obj_text="mtllib m.mtl\n", mtl_text=mtl)
    else:
        sc.load_obj(obj_text="mtllib m.mtl\n", mtl_text=mtl, decode=decode)
    for geo, name in objects:
        sc.add_mesh(*geo, names.index(name) + 1)
    return sc


def fixture(ptc, standins, which):
    cam = {"position": (-4.0, 1.0, 0.2), "fov": 60.0, "target": (0.0, 0.8, 0.0)}
    if which == "single_triangle":
        mtl = "newmtl a\nKd 0.8 0.6 0.4\nPr 0.35\n"
        P = np.array([[1.0, -1.0, -2.5], [1.0, 3.0, 0.0], [1.0, -1.0, 2.5]], F)
        tri = (P, np.tile(np.array([-1, 0, 0], F), (3, 1)), np.array([[0, 0], [1, 0], [0, 1]], F), np.array([0, 1, 2], np.uint32))
        return scene_from(ptc, standins, mtl, [(tri, "a")]), cam, 4
    if which == "furnace_mirror":
        # closed emissive room (every escaping path ends on a wall: radiance Ke) with a near-perfect mirror inside
        mtl = "newmtl wall\nKd 1 1 1\nKe 0.75 0.5 0.25\n\nnewmtl mirror\nKd 0.9 0.9 0.9\nPr 0.0\nPm 0.0\n"
        objs = [(g, "wall") for g in (quad((-6, -3, -5), (12, 0, 0), (0, 0, 10)), quad((-6, 6, -5), (0, 0, 10), (12, 0, 0)),
                                      quad((-6, -3, -5), (0, 9, 0), (12, 0, 0)), quad((-6, -3, 5), (12, 0, 0), (0, 9, 0)),
                                      quad((-6, -3, -5), (0, 0, 10), (0, 9, 0)), quad((6, -3, -5), (0, 9, 0), (0, 0, 10)))]
        objs.append((quad((1.0, -1.0, -2.0), (0, 3.5, 0), (0.6, 0, 4.0)), "mirror"))
        return scene_from(ptc, standins, mtl, objs), cam, 4
    if which == "glass_slab":
        mtl = "newmtl glass\nKd 0.9 0.95 1.0\nPr 0.05\nNi 1.5\nTf 1 1 1\n"
        objs = [(quad((0.5, -1.5, -2.5), (0, 4.5, 0), (0, 0, 5.0)), "glass"), (quad((0.9, -1.5, -2.5), (0, 0, 5.0), (0, 4.5, 0)), "glass")]
        return scene_from(ptc, standins, mtl, objs), cam, 4
    if which == "textured_wall":
        # file-backed albedo (RGBA8), normal map (RGBA8, consumed as a tangent-space normal) and roughness (float texels) on one
        # wall whose uv run from -0.6 to 2.7: bilinear taps, Repeat addressing on both sides of 0, uv interpolation, the
        # tangent frame of bsdf.wgsl:46-57; the wall is planar, so every path leaves for the sky after one bounce (smooth)
        mtl = "newmtl wall\nKd 1 1 1\nmap_Kd albedo.png\nmap_bump bump.png\nmap_Pr rough.png\nPm 0.4\n"

        def decode(path):
            name = path.rsplit("/", 1)[-1]
            if name.startswith("albedo"):
                return standins.make_albedo_texture(64, 32, 11, (0.9, 0.8, 0.7))
            if name.startswith("rough"):
                t = standins.make_bump_texture(48, 48, 5)
                return ptc.Texture((t.texels.astype(np.float32) / np.float32(255.0) * np.float32(0.5)).astype(np.float32))
            return standins.make_bump_texture(32, 64, 3)
        P, N, UV, I = quad((1.0, -2.0, -3.5), (0, 6.0, 0), (0.8, 0, 7.0))
        UV = (UV * F(3.3) - F(0.6)).astype(F)
        return scene_from(ptc, standins, mtl, [((P, N, UV, I), "wall")], decode=decode), cam, 4
    if which == "diffuse_corner":
        mtl = "newmtl floor\nKd 0.7 0.7 0.7\n\nnewmtl red\nKd 0.8 0.2 0.2\nPr 0.6\n\nnewmtl lamp\nKd 1 1 1\nKe 4 4 4\n"
        objs = [(quad((-6, 0, -6), (0, 0, 12), (12, 0, 0)), "floor"), (quad((2.5, 0, -6), (0, 0, 12), (0, 6, 0)), "red"),
                (quad((-6, 0, 3.0), (12, 0, 0), (0, 6, 0)), "red"), (quad((0, 4.5, -1), (1.5, 0, 0), (0, 0, 1.5)), "lamp")]
        return scene_from(ptc, standins, mtl, objs), cam, 4
    raise KeyError(which)


def compare(ptc, standins, oracle, which, W=40, H=30, frames=(1, 7), stride=5, p_light=0.0, spp=1):
    sc, cam, bounces = fixture(ptc, standins, which)
    cfg = ptc.make_config(W, H, max_bounces=bounces, send_to_light_probability=p_light, samples_per_pixel=spp)
    camera = standins.make_camera(cam, W, H)
    orc = oracle.OracleScene(sc, cfg)
    world = World(sc, bounces, p_light, spp)
    rows = []
    for frame in frames:
        un = camera.uniforms(frame=frame)
        planes = orc.render_frame(un)                                  # depth / normal / albedo of the first hit (pathtracer.wgsl:270-273)
        for k in range(0, W * H, stride):
            x, y = k % W, k // W
            want, state, rays = orc.pixel(un, x, y, brute=True)
            got, rng, my_rays, (t, n1, a1) = pixel(world, un, x, y)
            scale = max(1e-3, float(np.abs(want[:3]).max()))
            aux_err = max(abs(float(t) - float(planes["depth"][y, x])) / max(1.0, float(t)),
                          float(np.abs(n1 - planes["normal"][y, x, :3]).max()), float(np.abs(a1 - planes["albedo"][y, x, :3]).max()))
            rows.append((float(np.abs(got - want[:3]).max()) / scale, rng.state == state, my_rays == rays, rays, aux_err, float(t) > 0))
    return np.array(rows, dtype=np.float64)


@pytest.mark.parametrize("which", ["single_triangle", "furnace_mirror", "glass_slab"])
def test_smooth_fixtures_agree_on_every_pixel(ptc, standins, oracle, which):
    r = compare(ptc, standins, oracle, which)
    assert len(r) >= 200
    assert r[:, 1].all() and r[:, 2].all(), f"{which}: RNG state / ray count differs on {int((1 - r[:, 1]).sum())} pixels"
    assert r[:, 0].max() <= 1e-5, f"{which}: radiance differs by {r[:, 0].max():.2e}"
    assert r[:, 4].max() <= 2e-5, f"{which}: first-hit depth / normal / albedo differ by {r[:, 4].max():.2e}"
    assert r[:, 3].max() >= 2                                          # paths really bounce


def test_textured_normal_mapped_fixture(ptc, standins, oracle):
    """File-backed albedo / normal / roughness textures: every path draws the same random numbers and traces the same number of
    rays; radiance agrees to 1e-5 on >= 98 % of the pixels and to 2e-4 on all (median 3e-7).  The tail is sensitivity, not a
    different formula: the GGX angle atan(alpha sqrt(u1 / (1 - u1))) and the environment map's sun amplify the ulp-level
    differences between this file's float64 bilinear taps / numpy transcendentals and the oracle's fp32 ones."""
    r = compare(ptc, standins, oracle, "textured_wall")
    assert len(r) >= 400
    assert r[:, 1].all() and r[:, 2].all(), f"RNG state / ray count differs on {int((1 - r[:, 1]).sum())} pixels"
    assert np.median(r[:, 0]) <= 2e-6 and (r[:, 0] <= 1e-5).mean() >= 0.98 and r[:, 0].max() <= 2e-4, (np.median(r[:, 0]), r[:, 0].max())
    assert r[:, 4].max() <= 2e-5 and r[:, 5].mean() > 0.3, r[:, 4].max()   # normal-mapped first-hit normal, textured albedo, depth
    assert (r[:, 3] >= 2).mean() > 0.3                                 # a good share of the paths hit the wall and bounce


def test_chaotic_multi_bounce_fixture(ptc, standins, oracle):
    r = compare(ptc, standins, oracle, "diffuse_corner")
    same = r[:, 1].astype(bool) & r[:, 2].astype(bool)
    assert same.mean() >= 0.95, f"only {same.mean():.3f} of the paths end in the same RNG state"
    assert r[same, 0].max() <= 2e-4, f"radiance differs by {r[same, 0].max():.2e} on paths with identical draws"
    assert r[:, 3].mean() >= 2.5                                       # several segments per path on average


def test_samples_per_pixel_loop(ptc, standins, oracle):
    """SAMPLES_PER_PIXEL = 3 (pathtracer/pathtracer.wgsl:258-266): one primary direction, three paths, the mean."""
    r = compare(ptc, standins, oracle, "glass_slab", spp=3, stride=7)
    assert r[:, 1].all() and r[:, 2].all() and r[:, 0].max() <= 1e-5 and r[:, 4].max() <= 2e-5, (r[:, 0].max(), r[:, 4].max())
    assert r[:, 3].min() >= 3


def test_running_mean_accumulation(ptc, standins, oracle):
    """main, pathtracer.wgsl:21-43: prev = max(0, output); new = (prev.rgb * prev.a + color) / (prev.a + 1), a + 1 — or
    (color, 1) while the cursor is grabbed.  Three frames of the oracle's output texture against this file's pixels."""
    W, H = 24, 18
    sc, cam, bounces = fixture(ptc, standins, "single_triangle")
    cfg = ptc.make_config(W, H, max_bounces=bounces)
    camera = standins.make_camera(cam, W, H)
    orc = oracle.OracleScene(sc, cfg)
    world = World(sc, bounces)
    img, mine = None, np.zeros((H, W, 4), F)
    for frame, grabbed in ((1, 0), (2, 0), (3, 0), (4, 1), (5, 0)):
        un = camera.uniforms(frame=frame)
        un.is_grabbed = grabbed
        img = orc.render_frame(un, img)["color"]
        for y in range(0, H, 3):
            for x in range(0, W, 3):
                col = pixel(world, un, x, y)[0]
                prev = np.maximum(F(0.0), mine[y, x])
                if grabbed == 0:
                    mine[y, x, :3] = (prev[:3] * prev[3] + col) / F(prev[3] + F(1.0))
                    mine[y, x, 3] = prev[3] + F(1.0)
                else:
                    mine[y, x] = (col[0], col[1], col[2], F(1.0))
        got, want = mine[::3, ::3], img[::3, ::3]
        assert np.array_equal(got[..., 3], want[..., 3]), frame
        assert np.abs(got[..., :3] - want[..., :3]).max() <= 1e-5 * max(1.0, float(np.abs(want[..., :3]).max())), frame
    assert (img[..., 3] == 2).all()                                    # 3 frames, a grabbed one (-> 1), one more


def test_dormant_light_branch(ptc, standins, oracle):
    """SEND_TO_LIGHT_PROBABILITY = 0.5 (the reference ships 0.0: pathtracer/pathtracer.wgsl:21-73,127-140,170-191 never run
    there): light choice by rand_int() % num_lights, two draws for the point on the light, vertex 0's normal as the
    shading normal, attenuation = albedo, origin without offset, fallback to the regular branch when the light is behind
    the surface.  Chaotic fixture, same bars as above."""
    r = compare(ptc, standins, oracle, "diffuse_corner", p_light=0.5)
    same = r[:, 1].astype(bool) & r[:, 2].astype(bool)
    assert same.mean() >= 0.95, f"only {same.mean():.3f} of the paths end in the same RNG state with the same ray count"
    assert r[same, 0].max() <= 2e-4 and r[same, 4].max() <= 2e-5, (r[same, 0].max(), r[same, 4].max())
    plain = compare(ptc, standins, oracle, "diffuse_corner")
    assert not np.array_equal(plain[:, 3], r[:, 3])                    # the branch really changes the paths


def test_rng_against_the_reference_known_answers():
    """SURVEY 8c vectors, through THIS file's integer restatement (a third derivation next to make_kat.py and the oracle)."""
    assert jenkins(0) == 0 and jenkins(1) == 0x124EA49D
    r = Rng(0, 0, 512, 1)
    assert r.state == 0xC0738807 and [r.int() for _ in range(4)] == [0x3A1AF3C9, 0x0DF48F23, 0x13A1C877, 0xC40A3AEB]
    r = Rng(959, 539, 1920, 2)
    assert r.state == 0x4B476A5D and [r.int() for _ in range(4)] == [0xD7CBBB2E, 0xED1F2465, 0x2FA5F28A, 0x204296CF]
