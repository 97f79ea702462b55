"""Short runs of the mutation fuzzers over the host code that reads untrusted input, built with AddressSanitizer + UBSan
(tools/fuzz.sh runs them for millions of inputs; profiles/r02zn_fuzz.log is the last long run):
  tools/fuzz_obj.cpp   — OBJ / MTL text through ptc_scene_load_obj_text (csrc/pt_host.cpp; reference Scene::load_obj,
                         src/scene.rs:124-257 over tobj 4.0.3): accepted scenes satisfy every invariant ptc_create relies on,
                         refused ones leave the scene untouched
  tools/fuzz_scene.cpp — hostile scene views through prepare_scene (csrc/scene_prep.cpp; reference RTStep::create,
                         src/render_steps/pt.rs:341-433) and the three BVH builders, then rays through the traversal of
                         pt_device.cuh compiled for the host, compared with brute force over pt_ray_tri
The decoders' fuzzer runs from tests/test_image_decode.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wgpu-pathtracer_b200", "csrc")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SAN = ["-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-pthread"]


def build(tmp_path, name, args):
    exe = str(tmp_path / name)
    r = subprocess.run(["g++"] + SAN + args + ["-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        if "sanitize" in r.stderr or "asan" in r.stderr or "ubsan" in r.stderr:
            pytest.skip("no AddressSanitizer / UBSan runtime here: " + r.stderr[-300:])
        raise AssertionError(r.stderr[-3000:])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:max_allocation_size_mb=2048", LD_LIBRARY_PATH=f"{CUDA}/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_obj_mtl_ingest_under_the_mutation_fuzzer(tmp_path):
    exe, env = build(tmp_path, "fuzz_obj", [f"-I{CUDA}/include", os.path.join(CSRC, "pt_host.cpp"), "-x", "c++", os.path.join(CSRC, "img_decode.cu"),
                                            "-x", "c++", os.path.join(ROOT, "tools", "fuzz_obj.cpp"), f"-L{CUDA}/lib64", "-lcudart", "-lz", "-ldl"])
    r = subprocess.run([exe, "8000", "5"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "0 sanitizer reports" in r.stdout
    accepted = int(r.stdout.split("fuzz_obj: ")[1].split(" loads accepted")[0])
    refused = int(r.stdout.split("), ")[1].split(" refused")[0])
    assert accepted > 500 and refused > 500          # both sides of the parser are exercised


def test_scene_preparation_and_traversal_under_the_fuzzer(tmp_path):
    exe, env = build(tmp_path, "fuzz_scene", ["-ffp-contract=off", "-mavx2", "-mfma", "-Wno-unused-function", os.path.join(ROOT, "tools", "fuzz_scene.cpp"),
                                              os.path.join(CSRC, "scene_prep.cpp"), os.path.join(CSRC, "bvh_build.cpp")])
    r = subprocess.run([exe, "250", "9"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "0 sanitizer reports" in r.stdout and "all equal to brute force" in r.stdout
    accepted = int(r.stdout.split("fuzz_scene: ")[1].split(" scene x builder")[0])
    nonfinite = int(r.stdout.split("accepted (")[1].split(" of them")[0])
    refused = int(r.stdout.split("), ")[1].split(" refused")[0])
    hits = int(r.stdout.split("rays traced, ")[1].split(" hits")[0])
    assert accepted > 300 and nonfinite > 5 and refused > 50 and hits > 300
