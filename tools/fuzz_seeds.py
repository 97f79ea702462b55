"""fuzz_seeds.py <dir> — the seed corpus of tools/fuzz_decode.cpp: small PNG / OpenEXR / PFM files that together reach
every branch of the decoders in csrc/img_decode.cu (all PNG colour types and bit depths, tRNS, Adam7, the five row filters;
EXR NONE / ZIPS / ZIP / PIZ with HALF and FLOAT channels, the 14-bit and the 16-bit wavelet; Radiance HDR run-length, flat
and old-style repeats; both PFM flavours).
PNGs are framed by hand over zlib so the corpus does not depend on an encoder's choices; EXRs come from OpenCV's OpenEXR."""
import os
import struct
import sys
import zlib

import numpy as np

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"


def chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))


def png(w, h, depth, ctype, rows, interlace=0, extra=(), split=1):
    """rows: list of already packed scan-lines (bytes) in stream order, each gets a filter byte cycling through 0..4"""
    raw = b"".join(bytes([k % 5]) + r for k, r in enumerate(rows))
    # filters 1..4 are applied by the DECODER to whatever bytes are there: any byte string is a valid filtered row
    z = zlib.compress(raw, 6)
    cut = [len(z) * k // split for k in range(split + 1)]
    idat = b"".join(chunk(b"IDAT", z[cut[k]:cut[k + 1]]) for k in range(split))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
            + b"".join(extra) + idat + chunk(b"IEND", b""))


def packed_rows(rng, w, h, depth, channels):
    nbytes = (w * depth * channels + 7) // 8
    return [rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes() for _ in range(h)]


def adam7_rows(rng, w, h, depth, channels):
    rows = []
    for x0, y0, dx, dy in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        pw = (w - x0 + dx - 1) // dx if w > x0 else 0
        ph = (h - y0 + dy - 1) // dy if h > y0 else 0
        if pw and ph:
            rows += packed_rows(rng, pw, ph, depth, channels)
    return rows


def main(out):
    import cv2
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(20261018)

    def put(name, data):
        with open(os.path.join(out, name), "wb") as f:
            f.write(data)

    chans = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
    plte = chunk(b"PLTE", rng.integers(0, 256, 3 * 16, dtype=np.uint8).tobytes())
    for ctype, depths in {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}.items():
        for depth in depths:
            w, h = 29, 17
            extra = []
            if ctype == 3:
                extra = [plte, chunk(b"tRNS", bytes([0, 128, 255]))]
                rows = packed_rows(rng, w, h, depth, 1)
                if depth == 8:      # the row filters turn the bytes into any index: a full palette keeps every one in range
                    extra[0] = chunk(b"PLTE", rng.integers(0, 256, 3 * 256, dtype=np.uint8).tobytes())
            else:
                rows = packed_rows(rng, w, h, depth, chans[ctype])
                if ctype == 0:
                    extra = [chunk(b"tRNS", struct.pack(">H", 1))]
                if ctype == 2:
                    extra = [chunk(b"tRNS", struct.pack(">HHH", 1, 2, 3))]
            put(f"c{ctype}d{depth}.png", png(w, h, depth, ctype, rows, extra=extra, split=1 + depth % 3))
    put("grey8_plain.png", png(29, 17, 8, 0, packed_rows(rng, 29, 17, 8, 1)))          # no colour key: the decoder's straight 8-bit loops
    put("rgb8_plain.png", png(29, 17, 8, 2, packed_rows(rng, 29, 17, 8, 3), split=2))
    put("adam7_rgb8.png", png(29, 17, 8, 2, adam7_rows(rng, 29, 17, 8, 3), interlace=1))
    put("adam7_g1.png", png(13, 9, 1, 0, adam7_rows(rng, 13, 9, 1, 1), interlace=1))
    put("adam7_rgba16.png", png(5, 3, 16, 6, adam7_rows(rng, 5, 3, 16, 4), interlace=1))
    yy, xx = np.mgrid[0:40, 0:64]
    smooth = ((np.sin(xx / 5) + np.cos(yy / 7)) * 60 + 128).astype(np.uint8)
    put("smooth_rgb8.png", png(64, 40, 8, 2, [np.dstack([smooth, smooth // 2, 255 - smooth])[y].tobytes() for y in range(40)]))

    noise = (rng.random((24, 40, 3), dtype=np.float32) * 4).astype(np.float32)
    ramp = np.zeros((45, 64, 4), np.float32)      # 32 + 13 rows: the odd block reaches the wavelet's single-row tail
    ramp[..., 0] = np.tile(np.arange(64, dtype=np.float32) / 64, (45, 1))
    ramp[..., 1] = np.tile((np.arange(45, dtype=np.float32) / 45)[:, None], (1, 64))
    ramp[..., 2] = 0.5
    ramp[..., 3] = 1.0
    # PIZ picks its 16-bit wavelet when a block holds more than 2^14 distinct 16-bit values; noise like that is stored raw
    # (the compressor loses), so the values are a smooth ramp over 31 000 half-float bit patterns, 32 768 pixels per block
    bits = (np.arange(40 * 1024, dtype=np.int64) % 32768 * 31000 // 32768).astype(np.uint16)
    wide = bits.view(np.float16).astype(np.float32).reshape(40, 1024)
    comp = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS, "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP, "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}
    for cname, c in comp.items():
        for half in (0, 1):
            t = cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT
            for iname, img in (("noise", noise), ("ramp", ramp)):
                assert cv2.imwrite(os.path.join(out, f"{iname}_{cname}{half}.exr"), img, [cv2.IMWRITE_EXR_TYPE, t, cv2.IMWRITE_EXR_COMPRESSION, c])
    assert cv2.imwrite(os.path.join(out, "wide_piz1.exr"), wide, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, comp["piz"]])
    # 160 rows x 128 x RGBA float = 320 KB of scan-lines: above the size from which decode_exr spreads its chunks over threads
    big = np.zeros((160, 128, 4), np.float32)
    big[..., 0] = np.tile(np.arange(128, dtype=np.float32) / 128, (160, 1))
    big[..., 1] = np.tile((np.arange(160, dtype=np.float32) / 160)[:, None], (1, 128))
    big[..., 2] = big[..., 0] * big[..., 1]
    big[..., 3] = 1.0
    for cname in ("piz", "zip", "zips"):
        assert cv2.imwrite(os.path.join(out, f"big_{cname}0.exr"), big, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_COMPRESSION, comp[cname]])
    assert cv2.imwrite(os.path.join(out, "grey_zip1.exr"), noise[..., 0].copy(), [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, comp["zip"]])

    # Radiance HDR: run-length encoded scan-lines (OpenCV's writer), a width below 8 (flat pixels), and a hand-made flat file
    # with the 1991 repeat marker (1, 1, 1, count)
    hdr = (rng.random((12, 40, 3), dtype=np.float32) ** 4 * 50).astype(np.float32)
    hdr[0, :25] = 0.25
    assert cv2.imwrite(os.path.join(out, "rle.hdr"), hdr)
    assert cv2.imwrite(os.path.join(out, "narrow.hdr"), hdr[:, :5].copy())
    flat = b"".join(bytes([10 + x, 20, 30 + y, 128]) for y in range(3) for x in range(4))
    put("flat_repeat.hdr", b"#?RGBE\nFORMAT=32-bit_rle_rgbe\n\n-Y 4 +X 4\n" + flat + bytes([9, 9, 9, 130, 1, 1, 1, 3]))
    put("rgb_le.pfm", b"PF\n5 3\n-1.0\n" + rng.random(45, dtype=np.float32).tobytes())
    put("grey_be.pfm", b"Pf\n5 3\n1.0\n" + rng.random(15, dtype=np.float32).astype(">f4").tobytes())
    return sorted(os.listdir(out))


if __name__ == "__main__":
    print(len(main(sys.argv[1])), "seed files in", sys.argv[1])
