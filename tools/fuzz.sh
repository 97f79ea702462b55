#!/bin/bash
# fuzz.sh <decode|obj|scene|all|tsan|hostemu> [iterations-per-seed] [rng-seeds...] — AddressSanitizer + UBSan builds of the library's
# host code that reads untrusted input, driven by the mutation fuzzers next to this script.  No GPU needed.
#   decode : csrc/img_decode.cu (PNG, OpenEXR, PFM; compiled as host C++ — the file holds no device code; nvJPEG parses JPEG
#            itself and is not exercised)                                      tools/fuzz_decode.cpp + tools/fuzz_seeds.py
#   obj    : csrc/pt_host.cpp (OBJ / MTL text -> scene arrays, tobj-4.0.3 semantics)            tools/fuzz_obj.cpp
#   scene  : csrc/scene_prep.cpp + csrc/bvh_build.cpp + the traversal of pt_device.cuh on the host (hostile scene views:
#            broken tables, degenerate / coincident / non-finite geometry, three builders)       tools/fuzz_scene.cpp
#   tsan   : the same three fuzzers (short runs; the OBJ scan cut into 64-byte chunks, the > 256 KB OpenEXR seeds) and
#            tools/prepare_threads.cpp (120 000 triangles through prepare_scene and the three host builders) under
#            ThreadSanitizer instead
#   hostemu: tests/hostemu (the device functions of pt_device.cuh compiled for the host) AND the oracle built with ASan +
#            UBSan and run through tests/test_hostemu_parity.py, test_display_path.py and test_oracle_fixtures.py
# A sanitizer report or a violated invariant ends the run with a non-zero exit code.  Last long run: profiles/r02zn_fuzz.log
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"; root="$(dirname "$here")"; csrc="$root/wgpu-pathtracer_b200/csrc"
work="${FUZZ_DIR:-/tmp/ptc_fuzz}"; mkdir -p "$work"
what="${1:-all}"; shift || true
iters="${1:-}"; shift || true
seeds=("$@"); [ ${#seeds[@]} -eq 0 ] && seeds=(1 2 3)
cuda="${CUDA_HOME:-/usr/local/cuda}"
san="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -pthread"
export LD_LIBRARY_PATH="$cuda/lib64:${LD_LIBRARY_PATH:-}" ASAN_OPTIONS=detect_leaks=1:max_allocation_size_mb=2048

if [ "$what" = decode ] || [ "$what" = all ]; then
    g++ $san -I"$cuda/include" -x c++ "$csrc/img_decode.cu" -x c++ "$here/fuzz_decode.cpp" -o "$work/fuzz_decode" -L"$cuda/lib64" -lcudart -lz -ldl
    python "$here/fuzz_seeds.py" "$work/seeds"
    for s in "${seeds[@]}"; do "$work/fuzz_decode" "${iters:-300000}" "$s" "$work"/seeds/*; done
fi
if [ "$what" = obj ] || [ "$what" = all ]; then
    g++ $san -I"$cuda/include" "$csrc/pt_host.cpp" -x c++ "$csrc/img_decode.cu" -x c++ "$here/fuzz_obj.cpp" -o "$work/fuzz_obj" -L"$cuda/lib64" -lcudart -lz -ldl
    for s in "${seeds[@]}"; do "$work/fuzz_obj" "${iters:-200000}" "$s"; done
fi
if [ "$what" = scene ] || [ "$what" = all ]; then
    # -ffp-contract=off -mavx2 -mfma: the flags of tests/hostemu (explicit fmaf only, as in the device build)
    g++ $san -ffp-contract=off -mavx2 -mfma -Wno-unused-function "$here/fuzz_scene.cpp" "$csrc/scene_prep.cpp" "$csrc/bvh_build.cpp" -o "$work/fuzz_scene"
    for s in "${seeds[@]}"; do "$work/fuzz_scene" "${iters:-5000}" "$s"; done
fi
if [ "$what" = tsan ]; then
    tsan="-std=c++17 -O1 -g -fsanitize=thread -pthread"
    export TSAN_OPTIONS=halt_on_error=1
    g++ $tsan -I"$cuda/include" -x c++ "$csrc/img_decode.cu" -x c++ "$here/fuzz_decode.cpp" -o "$work/fuzz_decode_tsan" -L"$cuda/lib64" -lcudart -lz -ldl
    python "$here/fuzz_seeds.py" "$work/seeds"
    "$work/fuzz_decode_tsan" "${iters:-3000}" "${seeds[0]}" "$work"/seeds/big_*
    g++ $tsan -I"$cuda/include" "$csrc/pt_host.cpp" -x c++ "$csrc/img_decode.cu" -x c++ "$here/fuzz_obj.cpp" -o "$work/fuzz_obj_tsan" -L"$cuda/lib64" -lcudart -lz -ldl
    PTC_OBJ_CHUNK_BYTES=64 "$work/fuzz_obj_tsan" "${iters:-3000}" "${seeds[0]}"
    g++ $tsan -ffp-contract=off -mavx2 -mfma -Wno-unused-function "$here/fuzz_scene.cpp" "$csrc/scene_prep.cpp" "$csrc/bvh_build.cpp" -o "$work/fuzz_scene_tsan"
    "$work/fuzz_scene_tsan" 150 "${seeds[0]}"
    g++ $tsan -ffp-contract=off -mavx2 -mfma -Wno-unused-function "$here/prepare_threads.cpp" "$csrc/scene_prep.cpp" "$csrc/bvh_build.cpp" -o "$work/prepare_threads_tsan"
    "$work/prepare_threads_tsan"
fi
if [ "$what" = hostemu ]; then
    g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -mavx2 -mfma -Wno-unused-function -fvisibility=hidden -pthread \
        -fsanitize=address,undefined -fno-sanitize-recover=undefined -shared -o "$work/libpt_hostemu_asan.so" \
        "$root/tests/hostemu/hostemu.cpp" "$root/tests/hostemu/warpsim.cpp" "$csrc/scene_prep.cpp" "$csrc/bvh_build.cpp"
    g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -mavx2 -mfma -Wno-unused-function -fvisibility=hidden -pthread \
        -fsanitize=address,undefined -fno-sanitize-recover=undefined -shared -o "$work/libpt_oracle_asan.so" "$root/oracle/pt_oracle.cpp"
    PTC_ORACLE_LIB="$work/libpt_oracle_asan.so" PTC_HOSTEMU_LIB="$work/libpt_hostemu_asan.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
        ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \
        python -m pytest "$root/tests/test_hostemu_parity.py" "$root/tests/test_display_path.py" "$root/tests/test_oracle_fixtures.py" -x -q -m "not gpu" -p no:cacheprovider
fi
