"""The reference's REAL material libraries through the product's loader (csrc/pt_host.cpp).

Inputs: tests/golden/ref_res/ — byte copies of /root/reference/res/**/*.mtl and of the Sponza textures (made by
tests/golden/fetch_ref_res.py; MANIFEST.txt holds their sha256).  Expected values come from two places that share no code
with the loader:
  * `expected_tables` below — a Python restatement of the MTL -> Material mapping written from src/scene.rs:134-167 and
    275-394 and from tobj 4.0.3's documented MTL rules (typed keys vs `unknown_param`), and
  * hand-derived facts: material counts, `Tf 1 1 1` -> the white default texture 0, `Ks 0 0 0` -> texture
    3 (from_scalar(0.0) carries the colour [0,0,0,1], texture.rs:61-76, and comes first), the first Sponza material's indices.
"""
import hashlib
import os

import numpy as np
import pytest

RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_res")
MTLS = {"sponza/sponza.mtl": 21, "classroom/classroom.mtl": 80, "san_miguel/san-miguel.mtl": 288, "bistro/bistro.mtl": 204,
        "dragon.mtl": 2, "cornell_box.mtl": 6, "cornell_box_pbr.mtl": 5, "glass_cube.mtl": 4, "cube.mtl": 2, "knob/mitsuba.mtl": 4}

TYPED = {"Ka", "Kd", "Ks", "Ns", "Ni", "d", "map_Ka", "map_Kd", "map_Ks", "map_Ns", "map_Bump", "map_bump", "bump", "map_d", "illum"}


def f32(x):
    return float(np.float32(x))


def rust_f32(s):
    """str::parse::<f32>: None unless the WHOLE string is a float literal (no inner or outer spaces)."""
    if not s or s != s.strip() or any(c in s for c in " \t_xX("):
        return None
    try:
        return f32(float(s))
    except ValueError:
        return None


def parse_mtl(text):
    """tobj 4.0.3 MTL rules: typed keys fill Option fields, every other key lands in unknown_param as the trimmed rest of line."""
    mats, cur = [], None
    for raw in text.splitlines():
        line = raw.strip()
        words = line.split()
        if not words or words[0].startswith("#"):
            continue
        key, rest = words[0], line[len(words[0]):].strip()
        if key == "newmtl":
            cur = {"name": rest, "unknown": {}}
            mats.append(cur)
        elif cur is None:
            continue
        elif key in ("Kd", "Ks"):
            cur[key] = [f32(float(w)) for w in words[1:4]]
        elif key == "Ni":
            cur["Ni"] = f32(float(words[1]))
        elif key in ("map_Kd", "map_Ks"):
            cur[key] = rest
        elif key in ("map_Bump", "map_bump", "bump"):
            cur["normal_texture"] = rest
        elif key in TYPED:
            pass
        else:
            cur["unknown"][key] = rest
    return mats


def expected_tables(mats, obj_dir):
    """scene.rs:77-122 (ten default textures, material 0) + :134-167 over `mats`.  A texture is ("color", rgba) /
    ("scalar", v) / ("file", path); dedup looks at `.color` / `.scalar`, and from_scalar(v) also sets color [v,0,0,1]."""
    tex = [("color", [1, 1, 1, 1]), ("scalar", f32(0.99)), ("color", [.5, .5, .5, 1]), ("scalar", 0.0), ("color", [0, 0, 0, 1]),
           ("scalar", 0.0), ("color", [.5, .5, 1, 1]), ("scalar", 0.0), ("scalar", 1.0), ("file", "<envmap>")]

    def color_of(t):
        return [f32(c) for c in t[1]] if t[0] == "color" else ([t[1], 0.0, 0.0, 1.0] if t[0] == "scalar" else None)

    def push(t):
        tex.append(t)
        return len(tex) - 1

    def join(p):
        return os.path.join(obj_dir, p) if obj_dir else p

    def known_vector(path, val):                                              # :364-394
        if path is not None:
            return push(("file", join(path)))
        if val is None:
            return None
        want = [val[0], val[1], val[2], 1.0]
        for i, t in enumerate(tex):
            if color_of(t) == want:
                return i
        return push(("color", want))

    def known_scalar(val):                                                    # :331-362
        if val is None:
            return None
        for i, t in enumerate(tex):
            if t[0] == "scalar" and t[1] == val:
                return i
        return push(("scalar", val))

    def prop(m, tex_key, key, scalar):                                        # :275-329
        u = m["unknown"]
        if tex_key in u:
            return push(("file", join(u[tex_key])))
        if key not in u:
            return None
        if scalar:
            v = rust_f32(u[key])
            if v is None:
                return None
            return known_scalar(v)
        parts = [rust_f32(s) for s in u[key].split(" ")]
        vec = [0.0 if p is None else p for p in parts]
        assert len(vec) == 3, "the reference panics here (scene.rs:305)"
        if key == "Ke" and vec == [0.0, 0.0, 0.0]:
            return None
        return known_vector(None, vec)

    def pick(i, default):
        return default if i is None else i

    out = [dict(albedo_texture_idx=0, roughness_texture_idx=1, specular_texture_idx=2, metallic_texture_idx=3, emissive_texture_idx=4,
                sheen_texture_idx=5, normal_texture_idx=6, transmission_texture_idx=7, ior_texture_idx=8,
                clearcoat_thickness=0.0, clearcoat_roughness=0.0, anisotropy=0.0, anisotropy_rotation=0.0)]
    for m in mats:
        d = {}
        d["albedo_texture_idx"] = pick(known_vector(m.get("map_Kd"), m.get("Kd")), 0)
        d["roughness_texture_idx"] = pick(prop(m, "map_Pr", "Pr", True), 1)
        d["specular_texture_idx"] = pick(known_vector(m.get("map_Ks"), m.get("Ks")), 2)
        d["metallic_texture_idx"] = pick(prop(m, "map_Pm", "Pm", True), 3)
        d["emissive_texture_idx"] = pick(prop(m, "map_Ke", "Ke", False), 4)
        d["sheen_texture_idx"] = pick(prop(m, "map_Ps", "Ps", True), 5)
        d["normal_texture_idx"] = pick(known_vector(m.get("normal_texture"), None), 6)
        d["transmission_texture_idx"] = pick(prop(m, "map_Tf", "Tf", False), 7)
        d["ior_texture_idx"] = pick(known_scalar(m.get("Ni")), 8)
        for field, key in (("clearcoat_thickness", "Pc"), ("clearcoat_roughness", "Pcr"), ("anisotropy", "Pn"), ("anisotropy_rotation", "Pnr")):
            v = rust_f32(m["unknown"].get(key, ""))
            d[field] = 0.0 if v is None else v
        out.append(d)
    return out, tex


def load_through_product(ptc, standins, rel, decode=None):
    text = open(os.path.join(RES, rel), encoding="utf-8", errors="replace").read()
    names = [m["name"] for m in parse_mtl(text)]
    # one triangle per material so every material is referenced; geometry is irrelevant here
    obj = ["mtllib x.mtl", "v 0 0 0", "v 1 0 0", "v 0 1 0", "vn 0 0 1"]
    for n in names:
        obj += [f"usemtl {n}", "f 1//1 2//1 3//1"]
    seen = []

    def dummy(path):
        seen.append(path)
        return ptc.Texture(np.full((2, 2, 4), 128, np.uint8), path=path)

    sc = ptc.Scene(standins.make_envmap(8, 4)).load_obj(obj_text="\n".join(obj) + "\n", mtl_text=text, obj_dir="DIR", decode=decode or dummy, decode_threads=1)
    return sc, text, seen


def test_manifest_matches_the_vendored_files():
    lines = [l.split() for l in open(os.path.join(RES, "MANIFEST.txt")) if not l.startswith("#")]
    assert len(lines) >= 30
    for digest, size, rel in lines:
        data = open(os.path.join(RES, rel), "rb").read()
        assert len(data) == int(size) and hashlib.sha256(data).hexdigest() == digest, rel


@pytest.mark.parametrize("rel", sorted(MTLS))
def test_real_mtl_material_table_and_texture_push_order(ptc, standins, rel):
    sc, text, seen = load_through_product(ptc, standins, rel)
    mats = parse_mtl(text)
    assert len(sc.materials) == MTLS[rel] == len(mats) + 1                     # hand-counted `newmtl` lines + default material 0
    want_m, want_t = expected_tables(mats, "DIR")
    for i, w in enumerate(want_m):
        got = {k: (int(sc.materials[i][k]) if k.endswith("_idx") else float(sc.materials[i][k])) for k in w}
        assert got == w, (rel, i, mats[i - 1]["name"] if i else "<default>")
    assert len(sc.textures) == len(want_t)
    files = [t[1] for t in want_t[10:] if t[0] == "file"]
    assert seen == files                                                       # file textures are pushed (and decoded) in this order, never de-duplicated
    for i, t in enumerate(want_t[:9] + want_t[10:], start=0):
        j = i if i < 9 else i + 1
        if t[0] == "file":
            assert sc.textures[j].path == t[1]
        else:
            c = [f32(x) for x in (t[1] if t[0] == "color" else [t[1], 0.0, 0.0, 1.0])]
            assert sc.textures[j].texels.reshape(-1).tolist() == c, (rel, j)
    # every mesh points at its material: mesh k uses material k + 1 (scene.rs:224, prev_num_materials = 1)
    assert sc.meshes["material_index"].tolist() == list(range(1, len(mats) + 1))


def test_sponza_hand_derived_facts(ptc, standins):
    """SURVEY F10 / Appendix B on the real res/sponza/sponza.mtl."""
    sc, text, seen = load_through_product(ptc, standins, "sponza/sponza.mtl")
    m = sc.materials
    # 10 defaults + 37 files + from_scalar(1.5) + one `Kd` constant (the material without map_Kd) + one non-zero `Ks` constant
    assert len(m) == 21 and len(sc.textures) == 50
    assert (m["transmission_texture_idx"][1:] == 0).all()          # `Tf 1 1 1` on all 20 -> white default texture 0: glass Sponza
    assert (m["emissive_texture_idx"][1:] == 4).all()              # `Ke 0 0 0` -> None -> default 4: no emitters
    assert int((m["specular_texture_idx"][1:] == 3).sum()) == 19   # `Ks 0 0 0` == from_scalar(0.0).color -> texture 3, not 4 (one material has a real Ks)
    assert (m["roughness_texture_idx"][1:] == 1).all() and (m["metallic_texture_idx"][1:] == 3).all() and (m["sheen_texture_idx"][1:] == 5).all()
    first = m[1]                                                    # sp_00_luk_mali
    assert (first["albedo_texture_idx"], first["normal_texture_idx"], first["ior_texture_idx"]) == (10, 11, 12)
    assert seen[0] == "DIR/sp_luk.JPG" and seen[1] == "DIR/sp_luk-bump.JPG"
    assert sc.textures[12].texels.reshape(-1).tolist() == [1.5, 0.0, 0.0, 1.0]
    assert (m["ior_texture_idx"][1:] == 12).all()                   # Ni 1.5 everywhere: one scalar texture, de-duplicated
    assert len(seen) == 37                                          # 19 map_Kd + 18 bump files


def test_classroom_and_san_miguel_hand_derived_facts(ptc, standins):
    sc, text, seen = load_through_product(ptc, standins, "classroom/classroom.mtl")
    m = sc.materials
    assert "Tf" not in text and (m["transmission_texture_idx"] == 7).all()       # no Tf at all -> no transmission
    lights = int((m["emissive_texture_idx"] != 4).sum())
    assert lights == sum(1 for l in text.splitlines() if l.split()[:1] == ["Ke"] and [float(x) for x in l.split()[1:4]] != [0.0, 0.0, 0.0])
    assert lights >= 3 and seen == []                                             # constant colours only
    sc, text, seen = load_through_product(ptc, standins, "san_miguel/san-miguel.mtl")
    assert len(sc.materials) == 288 and len(seen) == sum(1 for l in text.splitlines() if l.split()[:1] and l.split()[0] in
                                                         ("map_Kd", "map_Ks", "map_Pr", "map_Pm", "map_Ke", "map_Ps", "map_Tf")) + \
        sum(1 for mm in parse_mtl(text) if "normal_texture" in mm)


def test_sponza_real_textures_decode_and_sample(ptc, standins, oracle):
    """The real Sponza JPGs through the product's decode path: 8-bit RGBA, value / 255 with no sRGB decode (texture.rs:241);
    the oracle's bilinear lookup of a decoded texel centre returns exactly that texel."""
    sc, _, _ = load_through_product(ptc, standins, "sponza/sponza.mtl", decode=lambda p: ptc.decode_image(os.path.join(RES, "sponza", os.path.basename(p))))
    t = sc.textures[10]
    assert t.texels.dtype == np.uint8 and t.texels.shape[2] == 4 and (t.texels[..., 3] == 255).all()
    h, w = t.texels.shape[:2]
    assert (w, h) == (1021, 150)                                    # sp_luk.JPG as stored in the reference repository
    orc = oracle.OracleScene(sc, ptc.make_config(8, 8))
    xs, ys = np.array([0, 5, w - 1]), np.array([0, 7, h - 1])
    uv = np.stack([(xs + 0.5) / w, (ys + 0.5) / h], axis=1).astype(np.float32)
    got = orc.sample_texture(10, uv)
    want = t.texels[ys, xs].astype(np.float32) / np.float32(255.0)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)
