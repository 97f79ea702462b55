"""rust/ is text (no Rust toolchain in this image), so it cannot be compiled here — these tests keep it from drifting:
build.rs must compile exactly the sources the Makefile links, with the same link libraries and numerics flags; ffi.rs must
declare every PTC_API symbol of include/ptcore.h with the same number of parameters."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(*p):
    return open(os.path.join(ROOT, *p)).read()


def test_build_rs_matches_the_makefile():
    mk, rs = read("wgpu-pathtracer_b200", "csrc", "Makefile"), read("rust", "build.rs")
    objs = re.search(r"^OBJS\s*=\s*(.+)$", mk, re.M).group(1).split()
    stems = sorted(o[:-2] for o in objs)
    srcs = re.findall(r'"(\w+\.(?:cu|cpp))"', rs)
    assert sorted(os.path.splitext(s)[0] for s in set(srcs)) == stems
    for s in set(srcs):
        assert os.path.exists(os.path.join(ROOT, "wgpu-pathtracer_b200", "csrc", s)), s
    link = re.search(r"^libptcore\.so:.*\n\t(.+)$", mk, re.M).group(1)
    for lib in re.findall(r"-l\w+", link):
        assert f'"{lib}"' in rs, f"{lib} is on the Makefile's link line but not in build.rs"
    for flag in ("-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-ffp-contract=off", "arch=compute_100a,code=sm_100a"):
        assert flag in mk and flag in rs, flag
    assert "-mfma" in rs and re.search(r"scene_prep\.o:.*\n\t.*-mfma", mk)
    for h in re.search(r"^HDRS\s*=\s*(.+)$", mk, re.M).group(1).split():
        if h.startswith("../"):
            continue
        assert f'"{h}"' in rs, f"{h}: rerun-if-changed missing in build.rs"


def c_decls():
    hdr = read("include", "ptcore.h")
    out = {}
    for m in re.finditer(r"PTC_API\s+[\w\s\*]+?\b(ptc_\w+)\s*\(([^)]*)\)\s*;", hdr):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_ffi_rs_declares_every_symbol():
    rs = read("rust", "ffi.rs")
    block = rs[rs.index('extern "C" {'):]
    got = {}
    for m in re.finditer(r"pub fn (ptc_\w+)\s*\(([^)]*)\)", block):
        args = m.group(2).strip()
        got[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    want = c_decls()
    assert len(want) >= 45
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    assert got == want, {k: (got[k], want[k]) for k in want if got[k] != want[k]}
    assert list(got) == list(want), "ffi.rs lists the symbols in the header's order"


def test_ffi_rs_struct_layouts_match_the_c_header():
    """field counts of the #[repr(C)] mirrors vs include/ptcore_types.h / ptcore.h (names are the same words)"""
    types, hdr, rs = read("include", "ptcore_types.h"), read("include", "ptcore.h"), read("rust", "ffi.rs")

    def c_fields(text, name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1]).lstrip("*"))
        return names

    def rs_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\}" % name, rs, re.S).group(1)
        return re.findall(r"pub (\w+):", body)

    for name, text in (("ptc_config", types), ("ptc_stats", types), ("ptc_mesh", types), ("ptc_texture_desc", types),
                       ("ptc_host_camera", hdr), ("ptc_scene_texture_info", hdr)):
        assert rs_fields(name) == c_fields(text, name), name
    view = [f for f in c_fields(types, "ptc_scene_view")]
    assert rs_fields("ptc_scene_view") == view
    for k, v in re.findall(r"pub const (PTC_PLANE_\w+): c_int = (\d+);", rs):
        assert re.search(r"%s\s*=\s*%s\b" % (k, v), types), k


def test_cuda_rt_step_only_calls_declared_functions():
    step, rs = read("rust", "cuda_rt_step.rs"), read("rust", "ffi.rs")
    for fn in set(re.findall(r"ffi::(ptc_\w+)\(", step)):
        assert f"pub fn {fn}" in rs, fn
    assert "rgba32f()" in step and "pub fn rgba32f(&mut self)" in step        # the accessor the maintainer adds is spelled out
    assert "pub fn rgba32f(&mut self)" in read("INTEGRATION.md")
