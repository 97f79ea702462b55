"""Regenerates tests/golden/kat.json: known-answer vectors for random.wgsl and camera.rs.

The reference has no tests of its own; these vectors were first derived in SURVEY.md §8c by restating
src/shaders/common/random.wgsl:7-36 in exact u32 arithmetic, and are re-derived here in pure Python integers
(no oracle, no product code) so the fixture is independent of both implementations.
"""
import json
import math
import os
import struct

M = 0xFFFFFFFF


def jenkins(x):
    x = (x + (x << 10)) & M; x ^= x >> 6; x = (x + (x << 3)) & M; x ^= x >> 11; x = (x + (x << 15)) & M
    return x


def pcg(state):
    state = (state * 747796405 + 2891336453) & M
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M
    return state, ((word >> 22) ^ word) & M


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def stream(x, y, w, frame, n=4):
    st = jenkins((((x + y * w) & M) * 1 & M) ^ jenkins(frame))
    s0, ints = st, []
    for _ in range(n):
        st, v = pcg(st)
        ints.append(v)
    return s0, ints


def main():
    out = {"jenkins": [["0x0", "0x%08x" % jenkins(0)], ["0x1", "0x%08x" % jenkins(1)], ["0xdeadbeef", "0x%08x" % jenkins(0xDEADBEEF)]],
           "rand_streams": [], "camera": []}
    for (x, y, w, f) in [(0, 0, 512, 1), (1, 0, 512, 1), (0, 1, 512, 1), (255, 255, 512, 1), (959, 539, 1920, 1), (959, 539, 1920, 2),
                         (3839, 2159, 3840, 4096)]:
        s0, ints = stream(x, y, w, f)
        case = {"x": x, "y": y, "width": w, "frame": f, "state": "0x%08x" % s0, "ints": ["0x%08x" % v for v in ints]}
        case["floats"] = [f32(f32(float(v)) / 4294967296.0) for v in ints[:2]]
        out["rand_streams"].append(case)
    # camera at the origin, fov 60, steady state looking down +X (camera.rs:54-79); fp32 via struct round trips
    for (w, h) in [(512, 512), (1920, 1080), (3840, 2160)]:
        vh = f32(2.0 * f32(math.tan(f32(f32(60.0 * f32(math.pi / 180.0)) / 2.0))))
        vw = f32(vh * f32(w / h))
        psx = f32(vw / w); psy = f32(vh / h)
        first = [1.0, f32(f32(vh / 2.0) - f32(0.5 * psy)), f32(f32(-vw / 2.0) + f32(0.5 * psx))]
        out["camera"].append({"width": w, "height": h, "psx_z": psx, "first_pixel": first})
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "kat.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
