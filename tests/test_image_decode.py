"""Texture decode inside the library (ptc_scene_decode_textures / ptc_decode_image; reference: Texture::load,
src/texture.rs:232-243, run on 32 threads by Scene::load_obj, src/scene.rs:239-254).

PNG, OpenEXR and PFM are decoded on the host: PNG checked bit for bit against OpenCV / Pillow on every colour type and bit depth, on
Adam7 streams written by an independent encoder in this file, and on damaged files; OpenEXR (NONE / ZIPS / ZIP / PIZ, HALF / FLOAT)
against OpenCV's OpenEXR, including the reference's real 71 MB environment map where the reference checkout is present.  JPEG goes through nvJPEG on the GPU
(`-m gpu`): the reference's real Sponza JPGs (tests/golden/ref_res, baseline and progressive) against OpenCV's libjpeg
within the tolerance two conforming IDCTs may differ by, then a textured OBJ rendered by the compiled host (ptrender)."""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_res")


def lib_decode(ptc, data, device=0):
    L = ptc._lib.lib()
    w, h, fmt, px = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_void_p()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    rc = L.ptc_decode_image(buf, len(data), device, C.byref(w), C.byref(h), C.byref(fmt), C.byref(px))
    if rc != 0:
        return rc, L.ptc_last_error().decode()
    n = w.value * h.value
    dt, k = (np.uint8, 4) if fmt.value == ptc.types.TEXEL_RGBA8 else (np.float32, 16)
    arr = np.frombuffer(C.string_at(px, n * k), dtype=dt).reshape(h.value, w.value, 4).copy()
    L.ptc_free(px)
    return 0, arr


def png_bytes(ihdr, rows_of_pass, extra_chunks=()):
    """A PNG written from scratch: rows_of_pass = list (one entry, or the seven Adam7 passes) of lists of raw row bytes."""
    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)
    raw = b"".join(b"\x00" + r for rows in rows_of_pass for r in rows)
    z = zlib.compress(raw, 6)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr)
    for t, body in extra_chunks:
        out += chunk(t, body)
    half = len(z) // 2
    return out + chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b"")      # two IDAT chunks: they must be concatenated


def expected_rgba(im):
    """decode_image's convention (wgpu-pathtracer_b200/scene.py) for an array in RGB(A) / grey order"""
    if im.ndim == 2:
        im = np.repeat(im[:, :, None], 3, axis=2)
    if im.shape[2] == 3:
        one = 255 if im.dtype == np.uint8 else 65535
        im = np.concatenate([im, np.full(im.shape[:2] + (1,), one, dtype=im.dtype)], axis=2)
    if im.dtype == np.uint16:
        im = im.astype(np.float32) / np.float32(65535.0)
    return im


@pytest.mark.parametrize("kind", ["grey8", "rgb8", "rgba8", "grey16", "rgb16", "rgba16"])
def test_png_matches_opencv(ptc, tmp_path, kind):
    import cv2
    rng = np.random.default_rng(7)
    H, W = 37, 53                                             # odd sizes: every filter meets a short last pixel
    dt = np.uint16 if kind.endswith("16") else np.uint8
    ch = {"grey": 0, "rgb": 3, "rgba": 4}[kind.rstrip("0123456789")]
    shape = (H, W) if ch == 0 else (H, W, ch)
    # smooth gradients + noise so that the encoder picks all five filter types
    base = (np.add.outer(np.arange(H), np.arange(W)) * (np.iinfo(dt).max // (H + W))).astype(np.int64)
    im = (base if ch == 0 else base[:, :, None] + np.arange(ch) * 1000) + rng.integers(0, 40, shape)
    im = np.clip(im, 0, np.iinfo(dt).max).astype(dt)
    path = str(tmp_path / f"{kind}.png")
    assert cv2.imwrite(path, im if ch < 3 else im[:, :, [2, 1, 0] + ([3] if ch == 4 else [])])
    rc, got = lib_decode(ptc, open(path, "rb").read())
    assert rc == 0, got
    want = expected_rgba(im)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_png_palette_low_bit_depths_and_transparency(ptc, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    H, W = 19, 45
    # palette with per-entry alpha (tRNS)
    idx = rng.integers(0, 13, (H, W)).astype(np.uint8)
    pal = rng.integers(0, 256, (13, 3)).astype(np.uint8)
    alpha = rng.integers(0, 256, 9).astype(np.uint8)                    # shorter than the palette: the rest is opaque
    for depth in (8, 4):
        if depth == 8:
            rows = [bytes(r) for r in idx]
        else:
            rows = [bytes(((np.pad(r, (0, W % 2))[0::2] << 4) | np.pad(r, (0, W % 2))[1::2]).astype(np.uint8)) for r in idx]
        data = png_bytes(struct.pack(">IIBBBBB", W, H, depth, 3, 0, 0, 0), [rows], [(b"PLTE", pal.tobytes()), (b"tRNS", alpha.tobytes())])
        rc, got = lib_decode(ptc, data)
        assert rc == 0, got
        want = np.concatenate([pal[idx], np.where(idx < 9, alpha[np.minimum(idx, 8)], 255)[:, :, None].astype(np.uint8)], axis=2)
        assert np.array_equal(got, want)
        p = tmp_path / f"pal{depth}.png"; p.write_bytes(data)
        assert np.array_equal(np.asarray(Image.open(p).convert("RGBA")), want)           # the hand-written file is a valid PNG
    # 1 / 2 / 4-bit grey: scaled to 8 bits
    for depth in (1, 2, 4):
        g = rng.integers(0, 1 << depth, (H, W)).astype(np.uint8)
        per = 8 // depth
        rows = []
        for r in g:
            r = np.pad(r, (0, (-W) % per)).reshape(-1, per)
            rows.append(bytes(sum((r[:, k].astype(np.uint32) << (8 - depth * (k + 1))) for k in range(per)).astype(np.uint8)))
        rc, got = lib_decode(ptc, png_bytes(struct.pack(">IIBBBBB", W, H, depth, 0, 0, 0, 0), [rows]))
        assert rc == 0, got
        v = (g.astype(np.uint32) * 255 // ((1 << depth) - 1)).astype(np.uint8)
        assert np.array_equal(got[..., 0], v) and np.array_equal(got[..., 1], v) and (got[..., 3] == 255).all()
    # grey + alpha, and a transparent colour key on RGB
    ga = rng.integers(0, 256, (H, W, 2)).astype(np.uint8)
    rc, got = lib_decode(ptc, png_bytes(struct.pack(">IIBBBBB", W, H, 8, 4, 0, 0, 0), [[bytes(r.reshape(-1)) for r in ga]]))
    assert rc == 0 and np.array_equal(got[..., 0], ga[..., 0]) and np.array_equal(got[..., 2], ga[..., 0]) and np.array_equal(got[..., 3], ga[..., 1])
    rgb = rng.integers(0, 4, (H, W, 3)).astype(np.uint8)
    key = rgb[5, 7]
    data = png_bytes(struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0), [[bytes(r.reshape(-1)) for r in rgb]], [(b"tRNS", struct.pack(">HHH", *map(int, key)))])
    rc, got = lib_decode(ptc, data)
    assert rc == 0 and np.array_equal(got[..., :3], rgb)
    assert np.array_equal(got[..., 3] == 0, (rgb == key).all(axis=2)) and got[5, 7, 3] == 0


@pytest.mark.parametrize("size", [(1, 1), (3, 2), (8, 8), (29, 17)])
def test_png_adam7(ptc, size):
    W, H = size
    rng = np.random.default_rng(W * 100 + H)
    im = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    X0, Y0, DX, DY = (0, 4, 0, 2, 0, 1, 0), (0, 0, 4, 0, 2, 0, 1), (8, 8, 4, 4, 2, 2, 1), (8, 8, 8, 4, 4, 2, 2)
    passes = []
    for k in range(7):
        sub = im[Y0[k]::DY[k], X0[k]::DX[k]]
        passes.append([bytes(r.reshape(-1)) for r in sub] if sub.size else [])
    rc, got = lib_decode(ptc, png_bytes(struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 1), passes))
    assert rc == 0, got
    assert np.array_equal(got[..., :3], im) and (got[..., 3] == 255).all()


def test_png_all_filter_types_by_hand(ptc):
    """rows filtered with each of the five types by an encoder written here (Sub, Up, Average, Paeth arithmetic mod 256)"""
    rng = np.random.default_rng(11)
    H, W, bpp = 10, 21, 4
    im = rng.integers(0, 256, (H, W, 4)).astype(np.uint8)
    flat = im.reshape(H, -1).astype(np.int32)
    raw = b""
    for y in range(H):
        cur, up = flat[y], (flat[y - 1] if y else np.zeros_like(flat[0]))
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]])
        t = y % 5
        if t == 0: f = cur
        elif t == 1: f = cur - left
        elif t == 2: f = cur - up
        elif t == 3: f = cur - ((left + up) >> 1)
        else:
            p = left + up - ul
            pa, pb, pc = abs(p - left), abs(p - up), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
            f = cur - pred
        raw += bytes([t]) + bytes((f % 256).astype(np.uint8))
    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    rc, got = lib_decode(ptc, data)
    assert rc == 0 and np.array_equal(got, im)


def test_damaged_files_are_errors_not_crashes(ptc):
    im = np.arange(5 * 4 * 3, dtype=np.uint8).reshape(5, 4, 3)
    good = png_bytes(struct.pack(">IIBBBBB", 4, 5, 8, 2, 0, 0, 0), [[bytes(r.reshape(-1)) for r in im]])
    assert lib_decode(ptc, good)[0] == 0
    for bad, why in [(good[:40], "truncated"), (good[:30] + bytes([good[30] ^ 1]) + good[31:], "CRC"), (b"GIF89a" + good[6:], "unknown image format"),
                     (good.replace(b"IDAT", b"iDAT"), "CRC")]:
        rc, msg = lib_decode(ptc, bad)
        assert rc != 0 and why.lower() in msg.lower(), (why, msg)
    # image data that inflates to the wrong size
    short = png_bytes(struct.pack(">IIBBBBB", 4, 6, 8, 2, 0, 0, 0), [[bytes(r.reshape(-1)) for r in im]])
    rc, msg = lib_decode(ptc, short)
    assert rc != 0 and "expected size" in msg


def test_pfm(ptc):
    rng = np.random.default_rng(5)
    im = rng.random((6, 9, 3)).astype(np.float32) * 4
    for endian, scale in (("<", b"-1.0"), (">", b"1.0")):
        data = b"PF\n9 6\n" + scale + b"\n" + im[::-1].astype(endian + "f4").tobytes()
        rc, got = lib_decode(ptc, data)
        assert rc == 0 and got.dtype == np.float32 and np.array_equal(got[..., :3], im) and (got[..., 3] == 1).all()
    grey = rng.random((4, 3)).astype(np.float32)
    rc, got = lib_decode(ptc, b"Pf\n3 4\n-1.0\n" + grey[::-1].tobytes())
    assert rc == 0 and np.array_equal(got[..., 0], grey) and np.array_equal(got[..., 2], grey)


def _cv2_exr():
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    return cv2


def _exr_expected(path):
    cv2 = _cv2_exr()
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert ref is not None
    if ref.ndim == 2:
        ref = np.repeat(ref[:, :, None], 3, axis=2)
    else:
        ref = ref[:, :, [2, 1, 0] + ([3] if ref.shape[2] == 4 else [])]
    if ref.shape[2] == 3:
        ref = np.concatenate([ref, np.ones(ref.shape[:2] + (1,), np.float32)], axis=2)
    return ref.astype(np.float32)


@pytest.mark.parametrize("compression", ["none", "zips", "zip", "piz"])
@pytest.mark.parametrize("half", [False, True])
def test_exr_matches_openexr(ptc, tmp_path, compression, half):
    """OpenEXR scan-line files written by OpenCV's OpenEXR (every compression this library reads, HALF and FLOAT, grey / RGB /
    RGBA, sizes that leave partial chunks and odd wavelet levels) decode to the same floats, bit for bit"""
    cv2 = _cv2_exr()
    comp = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS, "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP,
            "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}[compression]
    rng = np.random.default_rng(9)
    for (H, W, ch) in [(1, 1, 3), (37, 53, 3), (70, 33, 4), (64, 64, 0), (45, 130, 3)]:
        shape = (H, W) if ch == 0 else (H, W, ch)
        smooth = np.add.outer(np.arange(H), np.arange(W)).astype(np.float32) / (H + W)
        im = (smooth if ch == 0 else smooth[:, :, None] * np.arange(1, ch + 1)) + rng.random(shape).astype(np.float32) * (0.05 if H > 1 else 1.0)
        im = im.astype(np.float32)
        if H == 45:
            im[:20] = 0.25                                     # constant rows: Huffman run-length codes, and a short value bitmap
        path = str(tmp_path / f"t_{compression}_{int(half)}_{H}x{W}.exr")
        assert cv2.imwrite(path, im, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, comp])
        rc, got = lib_decode(ptc, open(path, "rb").read())
        assert rc == 0, (path, got)
        want = _exr_expected(path)
        assert got.dtype == np.float32 and got.shape == want.shape and np.array_equal(got, want), path


def test_exr_reference_environment_map(ptc):
    """the reference's own res/knob/envmap.exr (4096 x 2048 FLOAT RGBA, PIZ; what Scene::new loads, scene.rs:93): every texel equal
    to OpenEXR's decode.  The file is 71 MB and lives only where /root/reference does."""
    path = "/root/reference/res/knob/envmap.exr"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this box")
    rc, got = lib_decode(ptc, open(path, "rb").read())
    assert rc == 0, got
    want = _exr_expected(path)
    assert got.shape == (2048, 4096, 4) and np.array_equal(got, want)
    assert 0.04 < got[..., :3].min() < 0.05 and 0.9 < got[..., :3].max() <= 1.0    # the range SURVEY records for this map (0.045 .. 1.0)


def test_exr_damaged_or_unsupported(ptc, tmp_path):
    cv2 = _cv2_exr()
    im = np.random.default_rng(2).random((40, 24, 3)).astype(np.float32)
    path = str(tmp_path / "a.exr")
    assert cv2.imwrite(path, im, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
    good = open(path, "rb").read()
    assert lib_decode(ptc, good)[0] == 0
    for bad in (good[:200], good[:len(good) - 50], good[:len(good) // 2] + bytes(len(good) - len(good) // 2)):
        rc, msg = lib_decode(ptc, bad)
        assert rc != 0 and "OpenEXR" in msg, msg
    path = str(tmp_path / "b.exr")
    assert cv2.imwrite(path, im, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_RLE])
    rc, msg = lib_decode(ptc, open(path, "rb").read())
    assert rc != 0 and "compression" in msg


OBJ = """mtllib m.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
usemtl a
f 1/1 2/2 3/3
usemtl b
f 1/1 3/3 4/4
"""


def test_scene_decode_equals_the_python_decoder(ptc, tmp_path):
    """load_obj(decode="library") — PNG files named by an MTL, decoded on the library's threads — gives the same scene,
    texel for texel, as the OpenCV path; a missing file is reported by name and the other textures stay decoded"""
    import cv2
    rng = np.random.default_rng(1)
    names = []
    for i, (dt, ch) in enumerate([(np.uint8, 3), (np.uint8, 4), (np.uint16, 3), (np.uint8, 0)]):
        shape = (16 + i, 24 - i) if ch == 0 else (16 + i, 24 - i, ch)
        im = rng.integers(0, np.iinfo(dt).max, shape).astype(dt)
        names.append(f"t{i}.png"); assert cv2.imwrite(str(tmp_path / names[-1]), im)
    (tmp_path / "m.mtl").write_text(f"newmtl a\nKd 1 1 1\nmap_Kd {names[0]}\nmap_Bump {names[1]}\nnewmtl b\nmap_Kd {names[2]}\nmap_Ks {names[3]}\n")
    (tmp_path / "s.obj").write_text(OBJ)
    env = np.ones((1, 1, 4), np.float32)
    a = ptc.Scene(env).load_obj(path=str(tmp_path / "s.obj"), decode="library", decode_threads=3)
    b = ptc.Scene(env).load_obj(path=str(tmp_path / "s.obj"))
    assert len(a.textures) == len(b.textures) > 10
    nfile = 0
    for ta, tb in zip(a.textures, b.textures):
        assert ta.texels.dtype == tb.texels.dtype and np.array_equal(ta.texels, tb.texels)
        nfile += ta.path is not None
    assert nfile == 4 and np.array_equal(a.materials, b.materials)
    os.remove(tmp_path / names[2])
    with pytest.raises(ptc.PtcError, match=names[2]):
        ptc.Scene(env).load_obj(path=str(tmp_path / "s.obj"), decode="library")


def test_jpeg_without_a_gpu_fails_loudly(ptc):
    """JPEG is decoded by nvJPEG on the GPU: on a box without one the call says so (there is no CPU fallback to hide behind)"""
    if ptc._lib.lib().ptc_device_count() > 0:
        pytest.skip("GPU present")
    data = open(os.path.join(RES, "sponza", "KAMEN.JPG"), "rb").read()
    rc, msg = lib_decode(ptc, data)
    assert rc != 0 and ("CUDA device" in msg or "nvjpeg" in msg.lower())


def sponza_jpgs():
    d = os.path.join(RES, "sponza")
    return sorted(f for f in os.listdir(d) if f.lower().endswith(".jpg"))


def jpeg_sof(b):
    """(marker, components, sampling factors) of the frame header: 0xC0 baseline, 0xC2 progressive"""
    i = 2
    while i + 4 <= len(b):
        if b[i] != 0xFF:
            i += 1; continue
        m = b[i + 1]
        if m in (0xC0, 0xC1, 0xC2):
            n = b[i + 9]
            return m, n, [b[i + 11 + 3 * k] for k in range(n)]
        if m == 0xD8 or 0xD0 <= m <= 0xD7 or m == 0x01 or m == 0xFF:
            i += 2; continue
        i += 2 + ((b[i + 2] << 8) | b[i + 3])
    return None, 0, []


@pytest.mark.gpu
def test_jpeg_nvjpeg_matches_libjpeg_on_the_reference_textures(ptc):
    """every Sponza JPG of the reference (baseline and progressive, grey and 4:2:0 colour): same size, and texels within what
    two conforming decoders differ by.  Measured on B200 (tools/jpg_stats.py): grey files max 1 LSB; colour files mean 0.46-0.71,
    99.9th percentile <= 5.3, isolated pixels up to 17 at chroma edges (libjpeg's triangle-filter chroma upsampling vs nvJPEG's)."""
    import cv2
    files = sponza_jpgs()
    assert len(files) >= 18
    kinds = set()
    for f in files:
        data = open(os.path.join(RES, "sponza", f), "rb").read()
        sof, ncomp, _ = jpeg_sof(data)
        kinds.add((sof, ncomp))
        rc, got = lib_decode(ptc, data)
        assert rc == 0, (f, got)
        ref = cv2.imread(os.path.join(RES, "sponza", f), cv2.IMREAD_UNCHANGED)
        ref = np.repeat(ref[:, :, None], 3, axis=2) if ref.ndim == 2 else ref[:, :, ::-1]
        assert got.dtype == np.uint8 and got.shape[:2] == ref.shape[:2] and (got[..., 3] == 255).all(), f
        diff = np.abs(got[..., :3].astype(np.int32) - ref.astype(np.int32))
        if ncomp == 1:
            assert diff.max() <= 1, (f, int(diff.max()))
        else:
            assert diff.mean() <= 1.0 and np.percentile(diff, 99.9) <= 8 and diff.max() <= 32, (f, int(diff.max()), float(diff.mean()))
    assert (0xC2, 1) in kinds and (0xC0, 3) in kinds, kinds          # progressive and baseline, grey and colour were all exercised


@pytest.mark.gpu
@pytest.mark.parametrize("env_format", ["pfm", "exr"])
def test_ptrender_loads_a_textured_obj(ptc, standins, tmp_path, env_format):
    """the compiled host renders an OBJ whose MTL is the reference's real sponza.mtl (JPG textures decoded by the library) and
    gets, bit for bit, the image the Python mirror renders from the same library-decoded textures"""
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "wgpu-pathtracer_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "ptrender"], stdout=subprocess.DEVNULL)
    d = tmp_path / "sponza"
    shutil.copytree(os.path.join(RES, "sponza"), d)
    mats = [l.split()[1] for l in open(d / "sponza.mtl") if l.startswith("newmtl")]
    # a wall of quads, one per material, facing the camera
    lines = ["mtllib sponza.mtl"]
    n = len(mats)
    for i in range(n):
        x0, x1 = i * 1.0, i * 1.0 + 1.0
        lines += [f"v 4 0 {x0}", f"v 4 0 {x1}", f"v 4 2 {x1}", f"v 4 2 {x0}"]
    lines += ["vt 0 0", "vt 3 0", "vt 3 3", "vt 0 3", "vn -1 0 0"]
    for i, m in enumerate(mats):
        b = i * 4
        lines += [f"usemtl {m}", f"f {b + 1}/1/1 {b + 2}/2/1 {b + 3}/3/1", f"f {b + 1}/1/1 {b + 3}/3/1 {b + 4}/4/1"]
    (d / "wall.obj").write_text("\n".join(lines) + "\n")
    env = standins.make_envmap(64, 32)
    env_path = str(tmp_path / f"sky.{env_format}")
    if env_format == "pfm":
        open(env_path, "wb").write(b"PF\n64 32\n-1.0\n" + env.texels[::-1, :, :3].astype("<f4").tobytes())
    else:                                                      # what Scene::new("res/knob/envmap.exr") loads: FLOAT RGBA, PIZ
        cv2 = _cv2_exr()
        assert cv2.imwrite(env_path, np.ascontiguousarray(env.texels[:, :, [2, 1, 0, 3]].astype(np.float32)),
                           [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
    W, H, frames = 256, 96, 2
    pos, look = (-6.0, 1.0, n / 2), (4.0, 1.0, n / 2)
    r = subprocess.run([os.path.join(csrc, "ptrender"), "--obj", str(d / "wall.obj"), "--env", env_path, "--size", f"{W}x{H}",
                        "--frames", str(frames), "--pos", *map(str, pos), "--look", *map(str, look), "--fov", "90", "--out", str(tmp_path / "cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "texture file(s) decoded" in r.stdout
    got = np.fromfile(tmp_path / "cpp.rgba32f", np.float32).reshape(H, W, 4)
    scene = ptc.Scene(ptc.Texture(env.texels[:, :, :].astype(np.float32)))
    scene.load_obj(path=str(d / "wall.obj"), decode="library")
    assert sum(t.path is not None for t in scene.textures) >= 18
    camera = standins.make_camera({"position": pos, "fov": 90.0, "target": look}, W, H)
    step = ptc.RTStep.create(scene, ptc.make_config(W, H))
    for f in range(1, frames + 1):
        step.update(camera.uniforms(frame=f)); step.render()
    ref = step.read(ptc.types.PLANE_COLOR)
    step.destroy()
    assert np.array_equal(got, ref)
    assert len(np.unique(np.round(got[..., :3], 3).reshape(-1, 3), axis=0)) > 500          # textured, not flat


# ---- hostile headers and the mutation fuzzer (tools/fuzz_decode.cpp) -----------------------------------------------------
def test_headers_that_promise_more_than_the_file_holds(ptc):
    """Scene files are untrusted: a few bytes must not make the library allocate gigabytes, overflow its size arithmetic
    or throw across the C ABI (each case below was found by tools/fuzz.sh or by reading the size arithmetic)."""
    import resource
    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xFFFFFFFF)
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    # PNG: 30000 x 30000 RGBA16 (7.2 GB of scan-lines) behind 11 bytes of image data
    huge_png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 30000, 30000, 16, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 3)) + chunk(b"IEND", b"")
    rc, msg = lib_decode(ptc, huge_png)
    assert rc != 0 and "too short" in msg, msg
    # PFM: dimensions whose product wraps around 2^64 to a small number
    for header in (b"PF\n9223372036854775807 9223372036854775807\n-1\n", b"PF\n4294967296 4294967296\n-1\n", b"PF\n99999999999999999999 3\n-1\n"):
        rc, msg = lib_decode(ptc, header + b"\0" * 64)
        assert rc != 0 and "PFM" in msg, msg
    # OpenEXR: a dataWindow of 2^31 x 2^31 (the product overflows int64 arithmetic), and one of 30000 x 30000 in a 400-byte file
    def exr(dw, compression):
        def attr(name, typ, val):
            return name + b"\0" + typ + b"\0" + struct.pack("<I", len(val)) + val
        chlist = b"".join(n + b"\0" + struct.pack("<IBBBBII", 2, 0, 0, 0, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
        head = struct.pack("<II", 20000630, 2) + attr(b"channels", b"chlist", chlist) + attr(b"compression", b"compression", bytes([compression])) \
            + attr(b"dataWindow", b"box2i", struct.pack("<iiii", *dw)) + b"\0"
        return head + b"\0" * 256
    rc, msg = lib_decode(ptc, exr((-2 ** 31, -2 ** 31, 2 ** 31 - 1, 2 ** 31 - 1), 0))
    assert rc != 0 and "bad OpenEXR header" in msg, msg
    for compression in (0, 3, 4):
        rc, msg = lib_decode(ptc, exr((0, 0, 29999, 29999), compression))
        assert rc != 0 and ("too short" in msg or "truncated" in msg), msg
    # none of the refusals allocated the promised image (ru_maxrss is in KB)
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 512 * 1024


def _fuzz_build(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fuzz_decode")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", f"-I{cuda}/include",
           "-x", "c++", os.path.join(root, "wgpu-pathtracer_b200", "csrc", "img_decode.cu"), "-x", "c++", os.path.join(root, "tools", "fuzz_decode.cpp"),
           "-o", exe, f"-L{cuda}/lib64", "-lcudart", "-lz", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no AddressSanitizer build of the decoders here: " + r.stderr[-300:])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:max_allocation_size_mb=2048", LD_LIBRARY_PATH=f"{cuda}/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    return root, exe, env


def test_decoders_under_the_mutation_fuzzer(ptc, tmp_path):
    """The seed corpus (tools/fuzz_seeds.py: every PNG colour type / bit depth / Adam7, EXR NONE / ZIPS / ZIP / PIZ in HALF and
    FLOAT with both wavelets, both PFM flavours) decodes through the built library, and 30 000 structure-aware mutations of
    it run through an AddressSanitizer + UBSan build of the same decoders without a report.  Long runs: tools/fuzz.sh."""
    _cv2_exr()
    import importlib.util
    root, exe, env = _fuzz_build(tmp_path)
    spec = importlib.util.spec_from_file_location("fuzz_seeds", os.path.join(root, "tools", "fuzz_seeds.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    names = mod.main(str(tmp_path / "seeds"))
    assert len(names) >= 47
    for n in names:
        rc, got = lib_decode(ptc, open(tmp_path / "seeds" / n, "rb").read())
        assert rc == 0, (n, got)
    files = [str(tmp_path / "seeds" / n) for n in names]
    r = subprocess.run([exe, "30000", "7"] + files, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "0 sanitizer reports" in r.stdout and " 30000 inputs" in r.stdout
    decoded = int(r.stdout.split(" inputs, ")[1].split(" decoded")[0])
    assert 1000 < decoded < 29000        # the mutations neither all bounce off a checksum nor all pass through unchanged


def test_host_allocation_failure_is_a_status_not_an_exception(ptc):
    """No C++ exception crosses the C ABI ('never abort/throw across the ABI'): with the address space capped, a
    valid 16000 x 16000 PNG whose 1 GB RGBA8 image cannot be allocated comes back as PTC_ERR_OOM with a message, and the
    process lives on.  Run in a child process (ctypes only) so the limit does not touch the test session."""
    import sys
    lib = ptc._lib.lib()._name
    script = r'''
import ctypes as C, resource, struct, zlib, sys
L = C.CDLL(sys.argv[1]); L.ptc_last_error.restype = C.c_char_p
def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
W = H = 16000
raw = (b"\0" + b"\0" * W) * H
png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b"")
del raw
vm = int(open("/proc/self/statm").read().split()[0]) * resource.getpagesize()
resource.setrlimit(resource.RLIMIT_AS, (vm + (700 << 20), vm + (700 << 20)))
w, h, f, p = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_void_p()
rc = L.ptc_decode_image(png, len(png), 0, C.byref(w), C.byref(h), C.byref(f), C.byref(p))
print("RC", rc, L.ptc_last_error().decode())
small = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 1, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0\x10\x20")) + chunk(b"IEND", b"")
rc = L.ptc_decode_image(small, len(small), 0, C.byref(w), C.byref(h), C.byref(f), C.byref(p))
print("AFTER", rc, w.value, h.value)
'''
    r = subprocess.run([sys.executable, "-c", script, lib], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"RC {ptc.types.ERR_OOM} ptc_decode_image: out of host memory" in r.stdout, r.stdout
    assert "AFTER 0 2 1" in r.stdout                                   # the library still works afterwards


def test_png_reference_texture_sets(ptc):
    """The reference's own PNG textures (res/san_miguel: 323 files, res/bistro: 622, and the small scenes; 1 GB in all, present
    only where /root/reference is): 8-bit rgb / rgba / grey / palette and 1-bit palette / grey, two of them with a tRNS colour
    key.  Every 9th file plus both colour-keyed ones decode to exactly OpenCV's (libpng) pixels — Pillow's for the colour-keyed
    files, since OpenCV ignores tRNS on grey images while the reference's `image` crate expands it to alpha.  The whole set
    was compared once the same way: 955 of 955 identical."""
    import glob
    import cv2
    from PIL import Image
    files = sorted(glob.glob("/root/reference/res/**/*.png", recursive=True))
    if len(files) < 900:
        pytest.skip("the reference checkout with its texture sets is not present")
    keyed = [f for f in files if os.path.basename(f) in ("Paris_LiquorBottle_01_Labels_BaseColor.png", "Paris_Table_cloth_01_BaseColor.png")]
    assert len(keyed) == 2
    kinds = set()
    for f in files[::9] + keyed:
        data = open(f, "rb").read()
        rc, got = lib_decode(ptc, data)
        assert rc == 0, (f, got)
        kinds.add((data[24], data[25]))
        if f in keyed:
            exp = np.asarray(Image.open(f).convert("RGBA"))
            assert (exp[..., 3] == 0).any()
        else:
            im = cv2.imread(f, cv2.IMREAD_UNCHANGED)                  # BGR(A) or grey
            exp = expected_rgba(im if im.ndim == 2 else im[..., [2, 1, 0] + ([3] if im.shape[2] == 4 else [])])
        assert exp.shape == got.shape and np.array_equal(exp, got), f
    assert {(8, 2), (8, 6), (8, 3), (1, 3)} <= kinds


def test_radiance_hdr(ptc, tmp_path):
    """Radiance RGBE (.hdr): run-length encoded and flat scan-lines written by OpenCV decode to exactly OpenCV's floats
    (value = mantissa * 2^(e - 136), the `image` crate's rule too); the 1991 repeat marker by hand; damaged files are errors."""
    import cv2
    rng = np.random.default_rng(3)
    for (h, w) in [(5, 3), (7, 8), (9, 40), (3, 200)]:
        img = (rng.random((h, w, 3), dtype=np.float32) ** 4 * 50).astype(np.float32)
        img[0, :min(w, 20)] = 0.25                                          # a run
        path = str(tmp_path / f"t_{h}x{w}.hdr")
        assert cv2.imwrite(path, img)
        data = open(path, "rb").read()
        rc, got = lib_decode(ptc, data)
        assert rc == 0, got
        ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1]
        assert got.dtype == np.float32 and np.array_equal(got[..., :3], ref) and (got[..., 3] == 1).all()
    # flat pixels with an old-style repeat: (9, 9, 9, 130) followed by (1, 1, 1, 3) = that pixel four times in all
    flat = b"".join(bytes([10 + x, 20, 30 + y, 128]) for y in range(3) for x in range(4))
    data = b"#?RGBE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 4\n" + flat + bytes([9, 9, 9, 130, 1, 1, 1, 3])
    rc, got = lib_decode(ptc, data)
    assert rc == 0, got
    assert np.array_equal(got[3, :, :3], np.full((4, 3), 9 * 2.0 ** (130 - 136), np.float32))
    assert np.array_equal(got[1, 2, :3], np.array([12, 20, 31], np.float32) * np.float32(2.0 ** (128 - 136)))
    for bad, why in [(data[:-3], "truncated"), (data.replace(b"-Y 4 +X 4", b"+Y 4 +X 4"), "orientation"),
                     (data.replace(b"-Y 4 +X 4", b"-Y 30000 +X 30000"), "too short"), (data.replace(b"-Y 4 +X 4", b"-Y 99999 +X 99999"), "image size"), (data.replace(b"32-bit_rle_rgbe", b"32-bit_rle_xyze"), "pixel format"),
                     (data[:-4] + bytes([1, 1, 1, 9]), "repeat")]:
        rc, msg = lib_decode(ptc, bad)
        assert rc != 0 and why in msg, (why, msg)


def test_radiance_hdr_reference_environment_map(ptc):
    """The reference ships one Radiance file, res/bistro/san_giuseppe_bridge_4k.hdr (4096 x 2048, run-length RGBE, 24 MB; only
    where /root/reference is present): it decodes to exactly OpenCV's floats."""
    import cv2
    path = "/root/reference/res/bistro/san_giuseppe_bridge_4k.hdr"
    if not os.path.exists(path):
        pytest.skip("the reference checkout is not present")
    rc, got = lib_decode(ptc, open(path, "rb").read())
    assert rc == 0, got
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1]
    assert got.shape == (2048, 4096, 4) and np.array_equal(got[..., :3], ref) and float(got[..., :3].max()) == 98304.0
