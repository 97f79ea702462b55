"""Host-side scene ingest (csrc/pt_host.cpp) against the behaviour of src/scene.rs:77-394 + tobj 4.0.3
GPU_LOAD_OPTIONS.  The reference has no loader tests; each case cites the line it restates."""
import numpy as np
import pytest

OBJ = """mtllib m.mtl
o first
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 0.25
vt 0 1
vn 0 0 1
usemtl a
f 1/1/1 2/2/1 3/3/1 4/4/1
usemtl b
f 1/1/1 3/3/1 4/4/1
o second
v 0 0 1
v 1 0 1
v 0 1 1
f -3//1 -2//1 -1//1
"""
MTL = """newmtl a
Kd 0.2 0.4 0.6
Ks 0 0 0
Ni 1.5
Pr 0.25
Ke 0 0 0
Tf 1 1 1
map_bump n.png

newmtl b
\tKd 1 1 1
\tKe 2 2 2
\tmap_Kd d.png
\tPm 0.99
\taniso 0.5
\tPn 0.125
"""


def tiny_decode(ptc):
    def decode(path):
        t = np.zeros((2, 4, 4), np.uint8)
        t[..., 0] = 10; t[..., 3] = 255
        return ptc.Texture(t, path=path)
    return decode


def test_default_scene(ptc, standins):
    """Scene::new (scene.rs:77-122): ten textures in fixed order, default material 0."""
    sc = ptc.Scene(standins.make_envmap(16, 8))
    vals = [t.texels.reshape(-1).tolist() for t in sc.textures[:9]]
    f = np.float32
    assert vals == [[1, 1, 1, 1], [f(0.99), 0, 0, 1], [.5, .5, .5, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 0, 0, 1], [.5, .5, 1, 1], [0, 0, 0, 1], [1, 0, 0, 1]]
    m = sc.materials[0]
    assert [int(m[k]) for k in ("albedo_texture_idx", "roughness_texture_idx", "specular_texture_idx", "metallic_texture_idx", "emissive_texture_idx",
                                "sheen_texture_idx", "normal_texture_idx", "transmission_texture_idx", "ior_texture_idx")] == list(range(9))


def test_load_obj_models_materials_textures(ptc, standins):
    sc = ptc.Scene(standins.make_envmap(16, 8))
    sc.load_obj(obj_text=OBJ, mtl_text=MTL, obj_dir="some/dir", decode=tiny_decode(ptc))
    # tobj: a new model per o/g AND per usemtl change with pending faces -> 3 meshes
    assert len(sc.meshes) == 3
    # quad triangulated as (a,b,c),(a,c,d) with single-index dedup (4 unique vertices, 6 indices)
    assert sc.meshes[0]["vertex_count"] == 4 and sc.meshes[0]["index_count"] == 6
    assert sc.indices[:6].tolist() == [0, 1, 2, 0, 2, 3]
    # second model re-starts its own vertex numbering; indices are rebased by prev_num_vertices (scene.rs:195-199)
    assert sc.meshes[1]["vertex_offset"] == 4 and sc.meshes[1]["vertex_count"] == 3 and sc.indices[6:9].tolist() == [4, 5, 6]
    # negative (relative) indices; no vt on this face -> uv stays 0 (scene.rs:180-183); the material carries over the `o`
    assert sc.meshes[2]["vertex_count"] == 3 and sc.meshes[2]["material_index"] == 2
    np.testing.assert_array_equal(sc.vertices["position"][7:10, 2], [1, 1, 1])
    assert (sc.vertices["uv"][7:10] == 0).all()
    # v flipped: uv.y = 1 - vt.y (scene.rs:182)
    np.testing.assert_array_equal(sc.vertices["uv"][:4], [[0, 1], [1, 1], [1, 0.75], [0, 0]])
    # material_index = mat_id + prev_num_materials (scene.rs:224)
    assert sc.meshes["material_index"].tolist() == [1, 2, 2]
    a, b = sc.materials[1], sc.materials[2]
    # texture push order and constant de-dup (scene.rs:134-167,297-316)
    assert a["albedo_texture_idx"] == 10                        # Kd 0.2 0.4 0.6 -> new colour texture
    assert a["roughness_texture_idx"] == 11                     # Pr 0.25 -> new scalar texture
    assert a["specular_texture_idx"] == 3                       # Ks 0 0 0 == colour of the metallic default [0,0,0,1]
    assert a["metallic_texture_idx"] == 3 and a["sheen_texture_idx"] == 5      # absent -> defaults
    assert a["emissive_texture_idx"] == 4                       # Ke 0 0 0 -> None (scene.rs:307-309)
    assert a["normal_texture_idx"] == 12                        # map_bump -> file texture
    assert a["transmission_texture_idx"] == 0                   # Tf 1 1 1 == colour of texture 0
    assert a["ior_texture_idx"] == 13                           # Ni 1.5 -> new scalar
    assert sc.textures[12].path == "some/dir/n.png"
    assert b["albedo_texture_idx"] == 14 and sc.textures[14].path == "some/dir/d.png"    # map_Kd wins over Kd
    assert b["metallic_texture_idx"] == 1                       # Pm 0.99 == the scalar of the roughness default (exact compare) -> texture 1
    assert b["emissive_texture_idx"] == 15 and b["transmission_texture_idx"] == 7
    assert b["anisotropy"] == np.float32(0.125)                 # "Pn" is read, Blender's "aniso" is not (scene.rs:161)
    np.testing.assert_array_equal(sc.textures[15].texels.reshape(-1), [2, 2, 2, 1])


def test_scalar_dedup_against_default(ptc, standins):
    sc = ptc.Scene(standins.make_envmap(16, 8))
    mtl = "newmtl m\nKd 1 0 0\nPr 0.99\nPm 0\nNi 1\n"
    sc.load_obj(obj_text="mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl m\nf 1 2 3\n", mtl_text=mtl)
    m = sc.materials[1]
    assert m["roughness_texture_idx"] == 1 and m["metallic_texture_idx"] == 3 and m["ior_texture_idx"] == 8
    assert m["albedo_texture_idx"] == 8          # Kd 1 0 0 equals from_scalar(1.0)'s colour [1,0,0,1] (texture.rs:61-76)
    assert len(sc.textures) == 10


def test_generated_normals_reference_bug_and_fix(ptc, standins):
    """scene.rs:202-215: face normals are computed with mesh-LOCAL indices into the GLOBAL vertex array."""
    obj = ("mtllib m.mtl\no a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl m\nf 1 2 3\n"
           "o b\nv 0 0 5\nv 0 1 5\nv 0 0 6\nf 4 5 6\n")
    mtl = "newmtl m\nKd 1 1 1\n"
    sc = ptc.Scene(standins.make_envmap(16, 8)).load_obj(obj_text=obj, mtl_text=mtl)
    n = sc.vertices["normal"]
    # mesh b's loop re-wrote mesh a's vertices (with mesh a's own face normal) and left its own at zero
    np.testing.assert_array_equal(n[:3], [[0, 0, 1]] * 3)
    np.testing.assert_array_equal(n[3:], 0)
    fx = ptc.Scene(standins.make_envmap(16, 8)).load_obj(obj_text=obj, mtl_text=mtl, fix_normals=True)
    np.testing.assert_array_equal(fx.vertices["normal"][:3], [[0, 0, 1]] * 3)
    np.testing.assert_array_equal(fx.vertices["normal"][3:], [[1, 0, 0]] * 3)


def test_polygon_fan_and_points_lines_ignored(ptc, standins):
    obj = "mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 2 1 0\nv 1 2 0\nv 0 1 0\nusemtl m\nf 1 2 3 4 5\nl 1 2\nf 1\n"
    sc = ptc.Scene(standins.make_envmap(16, 8)).load_obj(obj_text=obj, mtl_text="newmtl m\nKd 1 1 1\n")
    assert sc.indices.tolist() == [0, 1, 2, 0, 2, 3, 0, 3, 4]


def test_malformed_inputs_return_errors(ptc, standins):
    sc = ptc.Scene(standins.make_envmap(16, 8))
    with pytest.raises(ptc.PtcError) as e:
        sc.load_obj(obj_text="mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl m\nf 1 2 9\n", mtl_text="newmtl m\nKd 1 1 1\n")
    assert e.value.status == ptc.types.ERR_PARSE
    with pytest.raises(ptc.PtcError):      # `Tf` with a double space: the reference panics at scene.rs:305
        ptc.Scene(standins.make_envmap(16, 8)).load_obj(obj_text="mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl m\nf 1 2 3\n",
                                                       mtl_text="newmtl m\nKd 1 1 1\nTf 1  1 1\n")
    with pytest.raises(ptc.PtcError) as e:
        ptc.Scene(standins.make_envmap(16, 8)).load_obj(path="/nonexistent/file.obj")
    assert e.value.status == ptc.types.ERR_IO


def test_lfs_pointer_is_reported(ptc, standins, tmp_path):
    p = tmp_path / "x.obj"
    p.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:00\nsize 5\n")
    with pytest.raises(ptc.PtcError) as e:
        ptc.Scene(standins.make_envmap(16, 8)).load_obj(path=str(p))
    assert "LFS" in str(e.value)


def test_load_obj_from_files_with_real_texture_decode(ptc, standins, tmp_path):
    import cv2
    img = np.zeros((8, 4, 3), np.uint8); img[..., 2] = 200; img[..., 1] = 100; img[..., 0] = 50      # BGR on disk
    cv2.imwrite(str(tmp_path / "d.png"), img)
    g16 = (np.arange(16, dtype=np.uint16) * 4096).reshape(4, 4)
    cv2.imwrite(str(tmp_path / "r.png"), g16)
    (tmp_path / "m.mtl").write_text("newmtl m\nKd 1 1 1\nmap_Kd d.png\nmap_Pr r.png\n")
    (tmp_path / "s.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl m\nf 1 2 3\n")
    sc = ptc.Scene(standins.make_envmap(16, 8)).load_obj(path=str(tmp_path / "s.obj"))
    t = sc.textures[10]
    assert t.texels.dtype == np.uint8 and t.size == (4, 8)
    assert t.texels[0, 0].tolist() == [200, 100, 50, 255]          # RGB order, alpha filled with 255, no sRGB decode
    r = sc.textures[11]
    assert r.texels.dtype == np.float32
    np.testing.assert_allclose(r.texels[1, 0], [4 * 4096 / 65535.0] * 3 + [1.0], rtol=1e-6)   # 16-bit -> /65535, grey -> rgb


def _scene_arrays(sc):
    return (sc.vertices.tobytes(), sc.indices.tobytes(), sc.meshes.tobytes(), sc.materials.tobytes())


@pytest.mark.parametrize("chunk", [16, 37, 64, 200])
def test_parallel_obj_parse_is_chunk_invariant(ptc, standins, monkeypatch, chunk):
    """The OBJ text is scanned in parallel chunks cut at line ends (pt_host.cpp).  Whatever the cut — here
    every few lines: through object / usemtl boundaries, between a face and the vertices its negative indices refer to,
    through n-gons — the scene equals the one parsed as a single chunk."""
    rng = np.random.RandomState(5)
    lines = ["mtllib m.mtl", "# comment", "o first"]
    nv = 0
    for obj in range(6):
        if obj:
            lines.append(("o" if obj % 2 else "g") + f" part{obj}")
        for blk in range(5):
            k = int(rng.randint(3, 7))
            for _ in range(k):
                lines.append("v %.6g %.6g %.6g" % tuple(rng.uniform(-2, 2, 3)))
                lines.append("vt %.6g %.6g" % tuple(rng.uniform(0, 1, 2)))
                lines.append("vn %.6g %.6g %.6g" % tuple(rng.uniform(-1, 1, 3)))
            nv += k
            if blk % 2 == 0:
                lines.append("usemtl " + "ab"[(obj + blk // 2) % 2])
            lines.append("f " + " ".join(f"{-k + j}/{-k + j}/{-k + j}" for j in range(k)))              # relative n-gon
            lines.append("f " + " ".join(f"{nv - k + j + 1}/{nv - k + j + 1}/{nv - k + j + 1}" for j in range(3)))   # absolute triangle
            lines.append("   \t  ")
            lines.append("l 1 2")
    text = "\r\n".join(lines) + "\n"
    scenes = []
    for c in (None, chunk):
        if c is None:
            monkeypatch.delenv("PTC_OBJ_CHUNK_BYTES", raising=False)
        else:
            monkeypatch.setenv("PTC_OBJ_CHUNK_BYTES", str(c))
        sc = ptc.Scene(standins.make_envmap(8, 4))
        sc.load_obj(obj_text=text, mtl_text=MTL, obj_dir="standin", decode=tiny_decode(ptc))
        scenes.append(sc)
    a, b = scenes
    assert len(a.meshes) > 12 and _scene_arrays(a) == _scene_arrays(b)


def test_obj_zero_and_plus_indices(ptc, standins):
    """`0` in a vt / vn slot reads as absent (as the empty field does); in the v slot it is an error; `+3` is 3."""
    base = "mtllib m.mtl\nusemtl a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\n"
    sc = ptc.Scene(standins.make_envmap(8, 4))
    sc.load_obj(obj_text=base + "f 1/0/1 +2//1 3/0/1\n", mtl_text=MTL, obj_dir="standin", decode=tiny_decode(ptc))
    assert len(sc.indices) == 3
    with pytest.raises(ptc.PtcError, match="bad face element"):
        ptc.Scene(standins.make_envmap(8, 4)).load_obj(obj_text=base + "f 0 1 2\n", mtl_text=MTL, obj_dir="standin", decode=tiny_decode(ptc))


def test_failed_load_rolls_the_scene_back(ptc, standins):
    """A load that fails late (here: a mesh without material in an OBJ that loads no MTL, detected after materials,
    vertices and indices were appended) must leave the scene exactly as it was: the reference panics, so no half-loaded
    Scene exists there either."""
    import ctypes as C
    L, T = ptc.lib(), ptc.types
    h = C.c_void_p()
    assert L.ptc_scene_new(b"<envmap>", C.byref(h)) == 0
    try:
        good = b"mtllib m.mtl\nusemtl a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n"
        assert L.ptc_scene_load_obj_text(h, good, b"newmtl a\nKd 0.25 0.5 0.75\n", b"") == 0
        ntex = L.ptc_scene_num_textures(h)
        texel = np.zeros((1, 1, 4), np.float32)
        assert L.ptc_scene_set_texels(h, 9, 1, 1, T.TEXEL_RGBA32F, texel.ctypes.data_as(C.c_void_p)) == 0
        v0 = T.SceneView(); assert L.ptc_scene_get_view(h, C.byref(v0)) == 0
        before = (v0.num_vertices, v0.num_indices, v0.num_meshes, v0.num_materials, v0.num_textures)
        # second load: first mesh fine (new material + constant texture are pushed), second mesh mixes normals -> PTC_ERR_PARSE
        bad = b"mtllib n.mtl\nusemtl b\nv 0 0 1\nv 1 0 1\nv 0 1 1\nf 1 2 3\no two\nvn 0 0 1\nv 0 0 2\nv 1 0 2\nv 0 1 2\nf 4//1 5//1 6\n"
        rc = L.ptc_scene_load_obj_text(h, bad, b"newmtl b\nKd 0.1 0.2 0.3\nKe 3 3 3\n", b"")
        assert rc == T.ERR_PARSE, L.ptc_last_error()
        assert L.ptc_scene_num_textures(h) == ntex
        v1 = T.SceneView(); assert L.ptc_scene_get_view(h, C.byref(v1)) == 0
        assert (v1.num_vertices, v1.num_indices, v1.num_meshes, v1.num_materials, v1.num_textures) == before
        assert v1.num_meshes == 1 and v1.num_indices == 3
    finally:
        L.ptc_scene_free(h)


def test_mtl_floats_do_not_depend_on_the_locale(ptc, standins):
    """MTL scalars are parsed like Rust's f32::from_str (locale-free): a process running under a comma-decimal LC_NUMERIC
    reads `Pr 0.25` as 0.25, not 0."""
    import locale
    mtl = "newmtl a\nKd 0.5 0.25 0.125\nPr 0.25\nNi 1.5\nPm +0.5\nPs 1e-1\n"
    obj = "mtllib m.mtl\nusemtl a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n"
    def load():
        sc = ptc.Scene(standins.make_envmap(8, 4)).load_obj(obj_text=obj, mtl_text=mtl)
        m = sc.materials[1]
        return [sc.textures[int(m[k])].texels.reshape(-1)[:3].tolist() for k in ("albedo_texture_idx", "roughness_texture_idx", "ior_texture_idx", "metallic_texture_idx", "sheen_texture_idx")]
    want = load()
    assert want[0] == [0.5, 0.25, 0.125] and want[1][0] == 0.25 and want[2][0] == 1.5 and want[3][0] == 0.5 and want[4][0] == np.float32(0.1)
    old = locale.setlocale(locale.LC_NUMERIC)
    switched = False
    for name in ("de_DE.UTF-8", "fr_FR.UTF-8", "de_DE", "fr_FR"):
        try:
            locale.setlocale(locale.LC_NUMERIC, name); switched = True; break
        except locale.Error:
            continue
    try:
        assert load() == want          # (when no comma locale is installed this still checks from_chars == the old result)
    finally:
        locale.setlocale(locale.LC_NUMERIC, old)
