#!/bin/bash
# Development aid: build a named variant of libptcore.so under build/<name>/ (git-ignored, travels to the GPU box)
# so several kernel variants can be A/B-timed in ONE gpurun call:  PTC_LIB=build/<name>/pkg/csrc/libptcore.so python bench.py
# usage: tools/build_variant.sh <name> ["extra nvcc flags"] [git-rev-for-csrc]
set -e
cd "$(dirname "$0")/.."
name="$1"; extra="$2"; rev="$3"
dst="build/$name"
rm -rf "$dst"; mkdir -p "$dst/pkg" "$dst/include"
if [ -n "$rev" ]; then
  git archive "$rev" wgpu-pathtracer_b200/csrc include | tar -x -C "$dst/tmp_$$" 2>/dev/null || { mkdir -p "$dst/tmp"; git archive "$rev" wgpu-pathtracer_b200/csrc include | tar -x -C "$dst/tmp"; }
  cp -r "$dst/tmp/wgpu-pathtracer_b200/csrc" "$dst/pkg/csrc"; cp "$dst/tmp/include/"* "$dst/include/"; rm -rf "$dst/tmp"
else
  cp -r wgpu-pathtracer_b200/csrc "$dst/pkg/csrc"; cp include/* "$dst/include/"
fi
rm -f "$dst"/pkg/csrc/*.o "$dst"/pkg/csrc/*.so
make -s -C "$dst/pkg/csrc" libptcore.so NVCC="nvcc $extra" 2>&1 | grep -E "error|warning: v" || true
ls -la "$dst/pkg/csrc/libptcore.so"
