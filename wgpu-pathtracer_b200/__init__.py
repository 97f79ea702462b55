"""wgpu-pathtracer_b200 — B200-native path-tracing core behind the reference's RenderStep interface.

Python here is host-side plumbing only (ctypes over the C ABI in include/ptcore.h); the hot path is the CUDA
library csrc/libptcore.so.  Import with ``importlib.import_module("wgpu-pathtracer_b200")`` (the directory
name carries the reference's hyphen).
"""
from . import _types as types
from ._lib import LIB_PATH, PtcError, lib
from .camera import Camera
from .render_step import RTStep, make_config
from .scene import Scene, Texture, decode_image

__all__ = ["types", "lib", "LIB_PATH", "PtcError", "Camera", "RTStep", "make_config", "Scene", "Texture", "decode_image"]
