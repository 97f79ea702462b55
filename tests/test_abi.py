"""The C-ABI shared library loads without a GPU and exports every symbol include/ptcore.h declares."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptcore.h")).read()
    return sorted(set(re.findall(r"PTC_API\s+[\w\s\*]+?\b(ptc_\w+)\s*\(", text)))


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "ptcore.h"\nint main(void){ ptc_config c; ptc_uniforms u; (void)c; (void)u; return sizeof(ptc_uniforms) == 112 && sizeof(ptc_camera) == 96 '
                   '&& sizeof(ptc_material) == 52 && sizeof(ptc_vertex) == 48 && sizeof(ptc_triangle) == 64 && sizeof(ptc_mesh) == 20 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_library_exports_every_declared_symbol(ptc):
    L = ptc.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ptcore.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", ptc.LIB_PATH], text=True)
    exported = set(re.findall(r" T (ptc_\w+)", out))
    assert exported == set(names), (exported ^ set(names))


def test_no_oracle_dependency_in_product(ptc):
    """The product library must not link or reference the oracle."""
    out = subprocess.check_output(["ldd", ptc.LIB_PATH], text=True)
    assert "oracle" not in out and "hostemu" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wgpu-pathtracer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pt_oracle" not in text and "oracle/" not in text.replace("the oracle/", ""), os.path.join(dirpath, f)


def test_version_and_error_strings(ptc):
    L = ptc.lib()
    assert b"sm_100a" in L.ptc_version()
    assert isinstance(L.ptc_last_error(), bytes)


def test_create_without_gpu_fails_loudly(ptc, standins):
    """No CPU fallback: without a CUDA device ptc_create returns PTC_ERR_NO_DEVICE."""
    if ptc.lib().ptc_device_count() > 0:
        pytest.skip("a GPU is present")
    scene, _ = standins.cornell()
    with pytest.raises(ptc.PtcError) as e:
        ptc.RTStep.create(scene, ptc.make_config(16, 16))
    assert e.value.status == ptc.types.ERR_NO_DEVICE and "no CPU path" in str(e.value)


def test_null_arguments_return_invalid(ptc):
    L = ptc.lib()
    T = ptc.types
    assert L.ptc_create(None, None, 0, None) == T.ERR_INVALID
    assert L.ptc_update(None, None) == T.ERR_INVALID
    assert L.ptc_render(None) == T.ERR_INVALID
    assert L.ptc_poll(None, 1) == T.ERR_INVALID
    assert L.ptc_read(None, 0, None, 0) == T.ERR_INVALID
    assert L.ptc_device_ptr(None, 0) is None
    L.ptc_destroy(None)


def test_scene_view_roundtrip(ptc, standins):
    scene, _ = standins.cornell()
    v = scene.view()
    assert v.num_vertices == len(scene.vertices) and v.num_indices == 36 * 3 and v.num_textures == len(scene.textures)
    raw = np.frombuffer(C.string_at(v.vertices, 48 * 3), dtype=ptc.types.VERTEX_DT)
    np.testing.assert_array_equal(raw["position"], scene.vertices["position"][:3])


def test_cpp_harness_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """harness/render_main.cpp + include/ptcore.hpp: the compiled-code host over the C ABI builds, and on a box without a
    CUDA device it reports the library's error instead of rendering anything (exit code 3 = PTC_ERR_NO_DEVICE)."""
    import subprocess
    csrc = os.path.join(ROOT, "wgpu-pathtracer_b200", "csrc")
    subprocess.check_call(["make", "-s", "-C", csrc, "ptrender"], stdout=subprocess.DEVNULL)
    exe = os.path.join(csrc, "ptrender")
    assert subprocess.run([exe, "--help"], capture_output=True).returncode == 0
    obj = tmp_path / "tri.obj"
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 1 1 1\n")
    obj.write_text("mtllib m.mtl\nusemtl a\nv 2 -1 -1\nv 2 1 -1\nv 2 0 1\nf 1 2 3\n")
    r = subprocess.run([exe, "--obj", str(obj), "--size", "32x32", "--frames", "1", "--out", str(tmp_path / "o")], capture_output=True, text=True)
    import ctypes
    has_gpu = ctypes.CDLL(os.path.join(csrc, "libptcore.so")).ptc_device_count() > 0
    if has_gpu:
        assert r.returncode == 0 and (tmp_path / "o.pfm").exists()
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr


def test_config_fields_without_meaning_are_refused(ptc, standins):
    """ptc_create validates the whole ptc_config before it looks for a device: options this build gives no meaning to
    (unknown compat flags, non-zero reserved words, an unsupported bvh_width, samples_per_pixel 0) are errors, never ignored."""
    scene, _ = standins.cornell()
    import ctypes as C
    L, T = ptc.lib(), ptc.types
    view = scene.view()
    def create(**kw):
        cfg = ptc.make_config(16, 16)
        for k, v in kw.items():
            if k == "reserved":
                cfg.reserved[v] = 1
            else:
                setattr(cfg, k, v)
        h = C.c_void_p()
        rc = L.ptc_create(C.byref(view), C.byref(cfg), 0, C.byref(h))
        if rc == 0:
            L.ptc_destroy(h)
        return rc, L.ptc_last_error().decode()
    for kw in (dict(compat_flags=1), dict(reserved=0), dict(reserved=3), dict(ploc_window=4), dict(ploc_window=65, bvh_build=2), dict(bvh_width=3), dict(samples_per_pixel=0), dict(accum_mode=7)):
        rc, msg = create(**kw)
        assert rc == T.ERR_INVALID, (kw, rc, msg)
