#!/bin/bash
# Rebuild every native library, then run a command on the GPU box through gpurun.
# usage: tools/gpu.sh <timeout-seconds> '<command>'
set -e
cd "$(dirname "$0")/.."
make -s -C wgpu-pathtracer_b200/csrc libptcore.so 2>&1 | grep -E "error|warning" || true
make -s -C oracle libpt_oracle.so
make -s -C tests/hostemu libpt_hostemu.so
mkdir -p gpurun_out
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "mkdir -p gpurun_out; $2"
