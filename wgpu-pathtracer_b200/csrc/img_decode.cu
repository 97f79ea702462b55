// img_decode.cu — see img_decode.h.  Host code (compiled by nvcc only for the CUDA runtime and nvjpeg.h declarations).
#include "img_decode.h"
#include "../../include/ptcore_types.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace ptc {

ImageKind sniff_image(const uint8_t* d, size_t n) {
    static const uint8_t png_sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n >= 8 && std::memcmp(d, png_sig, 8) == 0) return IMG_PNG;
    if (n >= 3 && d[0] == 0xFF && d[1] == 0xD8 && d[2] == 0xFF) return IMG_JPEG;
    if (n >= 3 && d[0] == 'P' && (d[1] == 'F' || d[1] == 'f') && (d[2] == '\n' || d[2] == ' ' || d[2] == '\r')) return IMG_PFM;
    if (n >= 8 && d[0] == 0x76 && d[1] == 0x2F && d[2] == 0x31 && d[3] == 0x01) return IMG_EXR;
    if (n >= 10 && (std::memcmp(d, "#?RADIANCE", 10) == 0 || std::memcmp(d, "#?RGBE", 6) == 0)) return IMG_HDR;
    return IMG_UNKNOWN;
}

// ---- PNG (ISO/IEC 15948) -----------------------------------------------------------------------------------------
namespace {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// reverses the row filters of one (sub-)image in place: `raw` holds h rows of (1 filter byte + rowbytes).
// The first bpp bytes of a row have no left neighbour and the first row no upper one; those cases are peeled off so the
// long loops carry no per-byte tests (a 2048 x 2048 RGB texture is 12 M filtered bytes, most of them Paeth).  A pixel-per-step
// SSE2 form of Sub / Average / Paeth (channels in 16-bit lanes) was measured on the reference's 955 PNGs and is SLOWER than
// these byte loops (61 s vs 51 s for 4.2 GB of scan-lines): consecutive bytes belong to different channels, so the scalar
// loop already runs bpp independent chains in flight.
bool unfilter(uint8_t* raw, uint32_t h, size_t rowbytes, uint32_t bpp) {
    const size_t stride = rowbytes + 1;
    const size_t head = std::min<size_t>(bpp, rowbytes);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* row = raw + (size_t)y * stride + 1;
        const uint8_t* up = y ? raw + (size_t)(y - 1) * stride + 1 : nullptr;
        const int type = raw[(size_t)y * stride];
        switch (type) {
            case 0: break;
            case 1: for (size_t i = bpp; i < rowbytes; ++i) row[i] = (uint8_t)(row[i] + row[i - bpp]); break;
            case 2: if (up) for (size_t i = 0; i < rowbytes; ++i) row[i] = (uint8_t)(row[i] + up[i]); break;
            case 3:
                if (up) {
                    for (size_t i = 0; i < head; ++i) row[i] = (uint8_t)(row[i] + (up[i] >> 1));
                    for (size_t i = head; i < rowbytes; ++i) row[i] = (uint8_t)(row[i] + ((row[i - bpp] + up[i]) >> 1));
                } else for (size_t i = head; i < rowbytes; ++i) row[i] = (uint8_t)(row[i] + (row[i - bpp] >> 1));
                break;
            case 4:
                if (up) {
                    for (size_t i = 0; i < head; ++i) row[i] = (uint8_t)(row[i] + up[i]);                    // paeth(0, b, 0) = b
                    for (size_t i = head; i < rowbytes; ++i) {
                        const int a = row[i - bpp], b = up[i], c = up[i - bpp];
                        const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
                        const int pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        row[i] = (uint8_t)(row[i] + pred);
                    }
                } else for (size_t i = head; i < rowbytes; ++i) row[i] = (uint8_t)(row[i] + row[i - bpp]);   // paeth(a, 0, 0) = a
                break;
            default: return false;
        }
    }
    return true;
}

struct PngInfo { uint32_t w, h, depth, ctype, channels; bool interlace; };

// sample k (0-based, in units of `depth` bits) of a row
inline uint32_t sample_at(const uint8_t* row, uint32_t depth, size_t k) {
    switch (depth) {
        case 8: return row[k];
        case 16: return ((uint32_t)row[k * 2] << 8) | row[k * 2 + 1];
        case 4: return (row[k >> 1] >> ((1 - (k & 1)) * 4)) & 15u;
        case 2: return (row[k >> 2] >> ((3 - (k & 3)) * 2)) & 3u;
        default: return (row[k >> 3] >> (7 - (k & 7))) & 1u;
    }
}

}  // namespace

int decode_png(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    if (sniff_image(d, n) != IMG_PNG) { err = "not a PNG"; return 1; }
    PngInfo I{}; bool have_ihdr = false, done = false;
    std::vector<uint8_t> idat, plte, trns;
    size_t p = 8;
    while (!done && p + 12 <= n) {
        const uint32_t len = be32(d + p);
        const uint8_t* type = d + p + 4, *body = d + p + 8;
        if ((size_t)len > n - p - 12) { err = "truncated PNG chunk"; return 1; }
        if (crc32(crc32(0L, type, 4), body, len) != be32(body + len)) { err = "PNG chunk CRC mismatch"; return 1; }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return 1; }
            I.w = be32(body); I.h = be32(body + 4); I.depth = body[8]; I.ctype = body[9]; I.interlace = body[12] != 0;
            if (body[10] != 0 || body[11] != 0 || body[12] > 1) { err = "unsupported PNG compression / filter / interlace method"; return 1; }
            static const int ch_of[7] = {1, 0, 3, 1, 2, 0, 4};
            if (I.ctype > 6 || ch_of[I.ctype] == 0) { err = "bad PNG colour type"; return 1; }
            I.channels = (uint32_t)ch_of[I.ctype];
            const bool depth_ok = I.ctype == 0 ? (I.depth == 1 || I.depth == 2 || I.depth == 4 || I.depth == 8 || I.depth == 16)
                                : I.ctype == 3 ? (I.depth == 1 || I.depth == 2 || I.depth == 4 || I.depth == 8) : (I.depth == 8 || I.depth == 16);
            if (!depth_ok || I.w == 0 || I.h == 0 || (uint64_t)I.w * I.h > (1ull << 30)) { err = "bad PNG dimensions or bit depth"; return 1; }
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) done = true;
        p += 12 + (size_t)len;
    }
    if (!done) { err = "truncated PNG (no IEND chunk)"; return 1; }
    if (!have_ihdr || idat.empty()) { err = "PNG without IHDR / IDAT"; return 1; }
    if (I.ctype == 3 && plte.size() < 3) { err = "palette PNG without PLTE"; return 1; }

    const uint32_t bits_px = I.depth * I.channels, bpp = bits_px >= 8 ? bits_px / 8 : 1;
    auto rowbytes_of = [&](uint32_t w) { return ((size_t)w * bits_px + 7) / 8; };
    // the (sub-)images of the stream: one, or the seven Adam7 passes
    struct Pass { uint32_t x0, y0, dx, dy, w, h; size_t offset; };
    std::vector<Pass> passes;
    size_t raw_bytes = 0;
    if (!I.interlace) { passes.push_back({0, 0, 1, 1, I.w, I.h, 0}); raw_bytes = (size_t)I.h * (rowbytes_of(I.w) + 1); }
    else {
        static const uint32_t X0[7] = {0, 4, 0, 2, 0, 1, 0}, Y0[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int k = 0; k < 7; ++k) {
            const uint32_t w = I.w > X0[k] ? (I.w - X0[k] + DX[k] - 1) / DX[k] : 0, h = I.h > Y0[k] ? (I.h - Y0[k] + DY[k] - 1) / DY[k] : 0;
            if (w == 0 || h == 0) continue;
            passes.push_back({X0[k], Y0[k], DX[k], DY[k], w, h, raw_bytes});
            raw_bytes += (size_t)h * (rowbytes_of(w) + 1);
        }
    }
    // deflate never expands by more than 1032 : 1: an IHDR that promises more than the IDAT bytes can hold is refused before
    // anything of the promised size is allocated (a 100-byte file cannot ask for gigabytes)
    if (raw_bytes / 1040 > idat.size()) { err = "PNG image data too short for its IHDR"; return 1; }
    std::vector<uint8_t> raw(raw_bytes);
    {
        z_stream zs; std::memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) { err = "zlib inflateInit failed"; return 1; }
        if (idat.size() > 0xFFFFFFFFull || raw.size() > 0xFFFFFFFFull) { inflateEnd(&zs); err = "PNG too large"; return 1; }
        zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size();
        zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
        const int rc = inflate(&zs, Z_FINISH);
        const bool full = zs.avail_out == 0;
        inflateEnd(&zs);
        if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || !full) { err = "PNG image data does not inflate to the expected size"; return 1; }
    }
    const bool wide = I.depth == 16;
    out.width = I.w; out.height = I.h; out.format = wide ? PTC_TEXEL_RGBA32F : PTC_TEXEL_RGBA8;
    out.texels.assign((size_t)I.w * I.h * (wide ? 16 : 4), 0);
    const uint32_t maxv = (1u << I.depth) - 1u;
    // a single transparent colour (tRNS of grey / rgb images): samples compared at the file's bit depth
    uint32_t key[3] = {0, 0, 0}; bool have_key = false;
    if (I.ctype == 0 && trns.size() >= 2) { key[0] = ((uint32_t)trns[0] << 8) | trns[1]; have_key = true; }
    if (I.ctype == 2 && trns.size() >= 6) { for (int k = 0; k < 3; ++k) key[k] = ((uint32_t)trns[k * 2] << 8) | trns[k * 2 + 1]; have_key = true; }
    for (const Pass& P : passes) {
        const size_t rb = rowbytes_of(P.w);
        if (!unfilter(raw.data() + P.offset, P.h, rb, bpp)) { err = "bad PNG filter type"; return 1; }
        // the shapes real texture sets are made of (8-bit grey / rgb / rgba / palette, not interlaced, no colour key) get
        // straight loops; everything else takes the general per-sample path below
        const bool fast = !I.interlace && I.depth == 8 && !have_key;
        for (uint32_t y = 0; y < P.h; ++y) {
            const uint8_t* row = raw.data() + P.offset + (size_t)y * (rb + 1) + 1;
            if (fast) {
                uint8_t* o = out.texels.data() + (size_t)y * I.w * 4;
                const uint32_t w = I.w;
                if (I.ctype == 6) { std::memcpy(o, row, (size_t)w * 4); continue; }
                if (I.ctype == 2) { for (uint32_t x = 0; x < w; ++x) { o[x * 4] = row[x * 3]; o[x * 4 + 1] = row[x * 3 + 1]; o[x * 4 + 2] = row[x * 3 + 2]; o[x * 4 + 3] = 255; } continue; }
                if (I.ctype == 0) { for (uint32_t x = 0; x < w; ++x) { o[x * 4] = o[x * 4 + 1] = o[x * 4 + 2] = row[x]; o[x * 4 + 3] = 255; } continue; }
                if (I.ctype == 4) { for (uint32_t x = 0; x < w; ++x) { o[x * 4] = o[x * 4 + 1] = o[x * 4 + 2] = row[x * 2]; o[x * 4 + 3] = row[x * 2 + 1]; } continue; }
                if (I.ctype == 3 && plte.size() >= 768) {               // a full palette: no index can leave it
                    for (uint32_t x = 0; x < w; ++x) { const uint32_t idx = row[x]; o[x * 4] = plte[idx * 3]; o[x * 4 + 1] = plte[idx * 3 + 1]; o[x * 4 + 2] = plte[idx * 3 + 2]; o[x * 4 + 3] = idx < trns.size() ? trns[idx] : 255u; }
                    continue;
                }
            }
            for (uint32_t x = 0; x < P.w; ++x) {
                uint32_t c[4] = {0, 0, 0, maxv};
                switch (I.ctype) {
                    case 0: c[0] = c[1] = c[2] = sample_at(row, I.depth, x); if (have_key && c[0] == key[0]) c[3] = 0; break;
                    case 2: for (int k = 0; k < 3; ++k) c[k] = sample_at(row, I.depth, (size_t)x * 3 + k);
                            if (have_key && c[0] == key[0] && c[1] == key[1] && c[2] == key[2]) c[3] = 0; break;
                    case 3: { const uint32_t idx = sample_at(row, I.depth, x);
                              if ((size_t)idx * 3 + 2 >= plte.size()) { err = "PNG palette index out of range"; return 1; }
                              c[0] = plte[idx * 3]; c[1] = plte[idx * 3 + 1]; c[2] = plte[idx * 3 + 2]; c[3] = idx < trns.size() ? trns[idx] : 255u; break; }
                    case 4: c[0] = c[1] = c[2] = sample_at(row, I.depth, (size_t)x * 2); c[3] = sample_at(row, I.depth, (size_t)x * 2 + 1); break;
                    default: for (int k = 0; k < 4; ++k) c[k] = sample_at(row, I.depth, (size_t)x * 4 + k); break;
                }
                const size_t px = (size_t)(P.y0 + y * P.dy) * I.w + (P.x0 + x * P.dx);
                if (wide) {
                    float* o = reinterpret_cast<float*>(out.texels.data()) + px * 4;
                    for (int k = 0; k < 4; ++k) o[k] = (float)c[k] / 65535.0f;
                } else {
                    uint8_t* o = out.texels.data() + px * 4;
                    if (I.ctype == 3 || I.depth == 8) for (int k = 0; k < 4; ++k) o[k] = (uint8_t)c[k];
                    else for (int k = 0; k < 4; ++k) o[k] = (uint8_t)((c[k] * 255u + maxv / 2u) / maxv);     // 1/2/4-bit grey scaled to 8 bits
                }
            }
        }
    }
    return 0;
}

// ---- PFM ("PF" = rgb, "Pf" = grey; rows bottom-up; negative scale = little endian) ---------------------------------------
int decode_pfm(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    if (sniff_image(d, n) != IMG_PFM) { err = "not a PFM"; return 1; }
    const int ch = d[1] == 'F' ? 3 : 1;
    size_t p = 2; long vals[2] = {0, 0}; double scale = 0.0;
    auto token = [&](std::string& t) { while (p < n && std::isspace(d[p])) ++p; t.clear(); while (p < n && !std::isspace(d[p])) t.push_back((char)d[p++]); return !t.empty(); };
    std::string t;
    for (int k = 0; k < 2; ++k) { if (!token(t)) { err = "bad PFM header"; return 1; } vals[k] = std::atol(t.c_str()); }
    if (!token(t)) { err = "bad PFM header"; return 1; }
    scale = std::atof(t.c_str());
    ++p;                                              // the single whitespace byte after the scale
    if (vals[0] <= 0 || vals[1] <= 0 || vals[0] > (1l << 30) || vals[1] > (1l << 30) || (uint64_t)vals[0] * (uint64_t)vals[1] > (1ull << 30) || scale == 0.0) { err = "bad PFM header"; return 1; }
    const size_t W = (size_t)vals[0], H = (size_t)vals[1];
    if (n < p || n - p < W * H * ch * 4) { err = "truncated PFM"; return 1; }
    const bool little = scale < 0.0;
    out.width = (uint32_t)W; out.height = (uint32_t)H; out.format = PTC_TEXEL_RGBA32F;
    out.texels.resize(W * H * 16);
    float* o = reinterpret_cast<float*>(out.texels.data());
    for (size_t y = 0; y < H; ++y)
        for (size_t x = 0; x < W; ++x) {
            float v[3] = {0, 0, 0};
            for (int k = 0; k < ch; ++k) {
                const uint8_t* s = d + p + (((H - 1 - y) * W + x) * ch + k) * 4;
                uint32_t u = little ? ((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24)) : be32(s);
                std::memcpy(&v[k], &u, 4);
            }
            if (ch == 1) v[1] = v[2] = v[0];
            float* q = o + (y * W + x) * 4; q[0] = v[0]; q[1] = v[1]; q[2] = v[2]; q[3] = 1.0f;
        }
    return 0;
}

// ---- OpenEXR: single-part scan-line images, HALF / FLOAT channels, compression NONE / ZIPS / ZIP / PIZ ---------------------
// (the reference's environment map res/knob/envmap.exr is 4096 x 2048 FLOAT RGBA, PIZ: Scene::new, scene.rs:93.)  Restated from
// the published file format: chunk = y, byte count, data; rows inside a chunk hold the channels in name order; ZIP = zlib +
// byte-delta predictor + even/odd de-interleave; PIZ = bitmap of used 16-bit values -> LUT, canonical Huffman with a run-length
// symbol, 2-D Haar-like wavelet per channel and per 16-bit half, LUT back.
namespace {

struct ExrChan { std::string name; int type; };          // type: 1 HALF, 2 FLOAT (0 UINT is not accepted)

inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

float half_to_float(uint16_t h) {
    const uint32_t sgn = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = sgn;
        else { int k = 0; uint32_t mm = m; while (!(mm & 1024u)) { mm <<= 1; ++k; } u = sgn | ((uint32_t)(113 - k) << 23) | ((mm & 1023u) << 13); }
    } else if (e == 31) u = sgn | 0x7F800000u | (m << 13);
    else u = sgn | ((e + 112u) << 23) | (m << 13);
    float f; std::memcpy(&f, &u, 4); return f;
}

bool zip_block(const uint8_t* src, size_t n, uint8_t* dst, size_t raw) {
    std::vector<uint8_t> tmp(raw);
    uLongf got = (uLongf)raw;
    if (uncompress(tmp.data(), &got, src, (uLong)n) != Z_OK || got != raw) return false;
    for (size_t i = 1; i < raw; ++i) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);       // predictor
    const size_t half = (raw + 1) / 2;                                                    // de-interleave
    for (size_t i = 0; i < raw; ++i) dst[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
    return true;
}

// --- PIZ: Huffman ---
const int kHufEncSize = (1 << 16) + 1;
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint64_t c = 0; int lc = 0;
    uint32_t get(int n) { while (lc < n) { c = (c << 8) | (p < end ? *p : 0); ++p; lc += 8; } lc -= n; return (uint32_t)((c >> lc) & ((1ull << n) - 1)); }
};
bool huf_uncompress(const uint8_t* src, size_t n, uint16_t* out, size_t nout) {
    if (n < 20) return nout == 0;
    const uint32_t im = le32(src), iM = le32(src + 4), nbits = le32(src + 12);
    if (im >= (uint32_t)kHufEncSize || iM >= (uint32_t)kHufEncSize || im > iM) return false;
    std::vector<uint8_t> len((size_t)kHufEncSize, 0);
    BitReader br{src + 20, src + n};
    for (uint32_t i = im; i <= iM; ++i) {                      // packed code lengths: 6 bits each, 59..62 short zero runs, 63 long
        const uint32_t l = br.get(6);
        if (l == 63u) { uint32_t z = br.get(8) + 6u; if (i + z > iM + 1) return false; i += z - 1; }
        else if (l >= 59u) { uint32_t z = l - 59u + 2u; if (i + z > iM + 1) return false; i += z - 1; }
        else len[i] = (uint8_t)l;
    }
    // the table was read by whole bytes: the bit stream starts at the next byte boundary
    const uint8_t* data = br.p - (br.lc / 8);
    // canonical codes, longest length first (hufCanonicalCodeTable)
    uint64_t cnt[59] = {0}, base[59] = {0};
    for (int i = 0; i < kHufEncSize; ++i) cnt[len[i]]++;
    { uint64_t c = 0; for (int l = 58; l > 0; --l) { const uint64_t nc = (c + cnt[l]) >> 1; base[l] = c; c = nc; } }
    std::vector<uint32_t> first(59, 0), sorted;                // symbols of each length in symbol order
    { uint32_t k = 0; for (int l = 1; l < 59; ++l) { first[l] = k; k += (uint32_t)cnt[l]; } sorted.resize(k); }
    { std::vector<uint32_t> pos(first); for (int i = 0; i < kHufEncSize; ++i) if (len[i]) sorted[pos[len[i]]++] = (uint32_t)i; }
    // fast table for codes of up to 12 bits: entry = (symbol << 8) | length, 0 = longer code
    const int kFast = 12;
    std::vector<uint32_t> fast((size_t)1 << kFast, 0);
    for (int l = 1; l <= kFast; ++l)
        for (uint64_t k = 0; k < cnt[l]; ++k) {
            const uint64_t code = base[l] + k;
            if (code >> l) return false;
            const uint32_t sym = sorted[first[l] + (uint32_t)k];
            const uint64_t lo = code << (kFast - l), hi = lo + ((uint64_t)1 << (kFast - l));
            for (uint64_t x = lo; x < hi; ++x) fast[x] = (sym << 8) | (uint32_t)l;
        }
    const uint32_t rlc = iM;                                   // the run-length symbol
    BitReader d{data, src + n};
    uint64_t bits_left = nbits;
    size_t o = 0;
    auto peek = [&](int k) -> uint32_t { while (d.lc < k) { d.c = (d.c << 8) | (d.p < d.end ? *d.p : 0); ++d.p; d.lc += 8; } return (uint32_t)((d.c >> (d.lc - k)) & ((1ull << k) - 1)); };
    while (bits_left > 0) {
        uint32_t sym; int l;
        const uint32_t e = fast[peek(kFast)];
        if (e) { sym = e >> 8; l = (int)(e & 255u); }
        else {
            l = kFast + 1; sym = 0xFFFFFFFFu;
            for (; l < 59; ++l) {
                if (!cnt[l]) continue;
                const uint64_t code = peek(l);
                if (code >= base[l] && code - base[l] < cnt[l]) { sym = sorted[first[l] + (uint32_t)(code - base[l])]; break; }
            }
            if (sym == 0xFFFFFFFFu) return false;
        }
        if ((uint64_t)l > bits_left) break;                    // padding bits at the end of the stream
        d.lc -= l; bits_left -= (uint64_t)l;
        if (sym == rlc) {
            if (bits_left < 8 || o == 0) return false;
            const uint32_t run = peek(8); d.lc -= 8; bits_left -= 8;
            if (o + run > nout) return false;
            const uint16_t v = out[o - 1];
            for (uint32_t k = 0; k < run; ++k) out[o++] = v;
        } else {
            if (o >= nout) return false;
            out[o++] = (uint16_t)sym;
        }
    }
    return o == nout;
}

// --- PIZ: wavelet ---
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int16_t ls = (int16_t)l, hs = (int16_t)h;
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai; b = (uint16_t)(int16_t)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h, bb = (m - (d >> 1)) & 0xFFFF, aa = (d + bb - 0x8000) & 0xFFFF;
    b = (uint16_t)bb; a = (uint16_t)aa;
}
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1; p2 = p; p >>= 1;
    while (p >= 1) {
        uint16_t* py = in; uint16_t* const ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py; uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) { uint16_t* p10 = px + oy1; if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10); *px = i00; }
        }
        if (ny & p) {
            uint16_t* px = py; uint16_t* const ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) { uint16_t* p01 = px + ox1; if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01); *px = i00; }
        }
        p2 = p; p >>= 1;
    }
}
bool piz_block(const uint8_t* src, size_t n, uint8_t* dst, size_t raw, int width, int rows, const std::vector<ExrChan>& ch) {
    if (n < 4) return false;
    const uint32_t minnz = src[0] | (src[1] << 8), maxnz = src[2] | (src[3] << 8);
    std::vector<uint8_t> bitmap(8192, 0);
    size_t p = 4;
    if (minnz <= maxnz) {
        if (maxnz >= 8192 || p + (maxnz - minnz + 1) > n) return false;
        std::memcpy(bitmap.data() + minnz, src + p, maxnz - minnz + 1); p += maxnz - minnz + 1;
    }
    std::vector<uint16_t> lut(65536, 0);
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; ++i) if (i == 0 || (bitmap[i >> 3] & (1u << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t maxv = (uint16_t)(k - 1);
    if (p + 4 > n) return false;
    const uint32_t hlen = le32(src + p); p += 4;
    if (p + hlen > n) return false;
    const size_t nvals = raw / 2;
    std::vector<uint16_t> tmp(nvals);
    if (!huf_uncompress(src + p, hlen, tmp.data(), nvals)) return false;
    // per channel: rows x (width * size) values; wavelet per 16-bit half
    size_t off = 0;
    std::vector<size_t> start(ch.size());
    for (size_t c = 0; c < ch.size(); ++c) {
        const int size = ch[c].type == 1 ? 1 : 2;
        start[c] = off;
        for (int j = 0; j < size; ++j) wav2_decode(tmp.data() + off + j, width, size, rows, width * size, maxv);
        off += (size_t)width * rows * size;
    }
    if (off != nvals) return false;
    for (size_t i = 0; i < nvals; ++i) tmp[i] = lut[tmp[i]];
    // back to scan-line order: row, then channel
    uint8_t* o = dst;
    for (int y = 0; y < rows; ++y)
        for (size_t c = 0; c < ch.size(); ++c) {
            const int size = ch[c].type == 1 ? 1 : 2;
            const uint16_t* r = tmp.data() + start[c] + (size_t)y * width * size;
            for (int x = 0; x < width * size; ++x) { *o++ = (uint8_t)(r[x] & 255u); *o++ = (uint8_t)(r[x] >> 8); }
        }
    return (size_t)(o - dst) == raw;
}

}  // namespace

int decode_exr(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    if (sniff_image(d, n) != IMG_EXR) { err = "not an OpenEXR file"; return 1; }
    const uint32_t version = le32(d + 4);
    if ((version & 0xFFu) != 2 || (version & 0x1A00u)) { err = "unsupported OpenEXR flavour (tiled, deep or multi-part)"; return 1; }
    std::vector<ExrChan> ch;
    int compression = -1; int dw[4] = {0, 0, -1, -1};
    size_t p = 8;
    for (;;) {                                                  // header: name\0 type\0 size value ... \0
        if (p >= n) { err = "truncated OpenEXR header"; return 1; }
        if (d[p] == 0) { ++p; break; }
        const char* name = (const char*)d + p; const size_t nl = strnlen(name, n - p); p += nl + 1;
        if (p >= n) { err = "truncated OpenEXR header"; return 1; }
        const char* type = (const char*)d + p; const size_t tl = strnlen(type, n - p); p += tl + 1;
        if (p + 4 > n) { err = "truncated OpenEXR header"; return 1; }
        const uint32_t sz = le32(d + p); p += 4;
        if ((size_t)sz > n - p) { err = "truncated OpenEXR header"; return 1; }
        const uint8_t* v = d + p; p += sz;
        if (!std::strcmp(name, "channels")) {
            size_t q = 0;
            while (q < sz && v[q] != 0) {
                ExrChan c; c.name = (const char*)v + q; q += c.name.size() + 1;
                if (q + 16 > sz) { err = "bad OpenEXR channel list"; return 1; }
                c.type = (int)le32(v + q);
                const uint32_t xs = le32(v + q + 8), ys = le32(v + q + 12); q += 16;
                if (c.type != 1 && c.type != 2) { err = "OpenEXR UINT channels are not supported"; return 1; }
                if (xs != 1 || ys != 1) { err = "sub-sampled OpenEXR channels are not supported"; return 1; }
                ch.push_back(c);
            }
        } else if (!std::strcmp(name, "compression") && sz >= 1) compression = v[0];
        else if (!std::strcmp(name, "dataWindow") && sz >= 16) for (int k = 0; k < 4; ++k) dw[k] = (int)le32(v + 4 * k);
    }
    const int64_t W = (int64_t)dw[2] - dw[0] + 1, H = (int64_t)dw[3] - dw[1] + 1;
    if (ch.empty() || W <= 0 || H <= 0 || W > (1ll << 30) || H > (1ll << 30) || W * H > (1ll << 30)) { err = "bad OpenEXR header (channels / dataWindow)"; return 1; }
    int lines;
    switch (compression) { case 0: case 2: lines = 1; break; case 3: lines = 16; break; case 4: lines = 32; break;
        default: err = "unsupported OpenEXR compression (only NONE, ZIPS, ZIP and PIZ)"; return 1; }
    size_t row_bytes = 0;
    for (const ExrChan& c : ch) row_bytes += (size_t)W * (c.type == 1 ? 2 : 4);
    const int64_t nchunks = (H + lines - 1) / lines;
    if (p + (size_t)nchunks * 8 > n) { err = "truncated OpenEXR offset table"; return 1; }
    // where R, G, B, A live inside a row (a lone Y channel is grey)
    int idx[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < ch.size(); ++c) {
        const std::string& nm = ch[c].name;
        if (nm == "R") idx[0] = (int)c; else if (nm == "G") idx[1] = (int)c; else if (nm == "B") idx[2] = (int)c; else if (nm == "A") idx[3] = (int)c;
        else if (nm == "Y" && idx[0] < 0) idx[0] = idx[1] = idx[2] = (int)c;
    }
    if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0) { err = "OpenEXR image without R, G, B (or Y) channels"; return 1; }
    std::vector<size_t> chan_off(ch.size());
    { size_t o = 0; for (size_t c = 0; c < ch.size(); ++c) { chan_off[c] = o; o += (size_t)W * (ch[c].type == 1 ? 2 : 4); } }
    // same plausibility bound as for PNG: stored rows need their own size in the file, deflate stays below 1032 : 1 and PIZ's
    // run-length codes below 408 : 1, so a dataWindow the file cannot fill is refused before the image is allocated
    if ((uint64_t)row_bytes * (uint64_t)H / (compression == 0 ? 1u : 1040u) > n) { err = "OpenEXR file too short for its dataWindow"; return 1; }
    out.width = (uint32_t)W; out.height = (uint32_t)H; out.format = PTC_TEXEL_RGBA32F;
    out.texels.assign((size_t)W * H * 16, 0);
    float* px = reinterpret_cast<float*>(out.texels.data());
    // chunks (1, 16 or 32 scan-lines) are independent: the large ones — the reference's environment map is 64 PIZ blocks of
    // 4096 x 32 x 4 floats — are decoded on several host threads, each with its own scratch block
    auto decode_chunk = [&](int64_t k, std::vector<uint8_t>& raw) -> const char* {
        const uint64_t at = le64(d + p + (size_t)k * 8);
        if (at > n || n - at < 8) return "bad OpenEXR chunk offset";
        const int64_t y0 = (int64_t)(int32_t)le32(d + at) - dw[1];
        const uint32_t csz = le32(d + at + 4);
        if (y0 < 0 || y0 >= H || (uint64_t)csz > n - at - 8) return "bad OpenEXR chunk";
        const int rows = (int)std::min<int64_t>(lines, H - y0);
        const size_t want = row_bytes * (size_t)rows;
        raw.resize(want);
        const uint8_t* src = d + at + 8;
        bool ok = true;
        if (csz == want) std::memcpy(raw.data(), src, want);                          // stored: the compressor did not win
        else if (compression == 2 || compression == 3) ok = zip_block(src, csz, raw.data(), want);
        else if (compression == 4) ok = piz_block(src, csz, raw.data(), want, (int)W, rows, ch);
        else ok = false;
        if (!ok) return "OpenEXR chunk does not decompress";
        for (int r = 0; r < rows; ++r) {
            const uint8_t* row = raw.data() + (size_t)r * row_bytes;
            float* o = px + ((size_t)(y0 + r) * W) * 4;
            for (int c4 = 0; c4 < 4; ++c4) {
                if (idx[c4] < 0) { for (int64_t x = 0; x < W; ++x) o[x * 4 + c4] = 1.0f; continue; }
                const uint8_t* cp = row + chan_off[idx[c4]];
                if (ch[idx[c4]].type == 1) for (int64_t x = 0; x < W; ++x) o[x * 4 + c4] = half_to_float((uint16_t)(cp[x * 2] | (cp[x * 2 + 1] << 8)));
                else for (int64_t x = 0; x < W; ++x) { const uint32_t u = le32(cp + x * 4); std::memcpy(&o[x * 4 + c4], &u, 4); }
            }
        }
        return nullptr;
    };
    // two chunks that name the same rows would be written by two threads at once: such a file is decoded by one thread
    // (last chunk wins, as before); a well-formed file names every block of rows exactly once
    int nthreads = 1;
    if (nchunks >= 4 && (uint64_t)row_bytes * (uint64_t)H >= (256u << 10)) {
        std::vector<uint8_t> seen((size_t)nchunks, 0); bool distinct = true;
        for (int64_t k = 0; k < nchunks && distinct; ++k) {
            const uint64_t at = le64(d + p + (size_t)k * 8);
            if (at > n || n - at < 8) { distinct = false; break; }
            const int64_t y0 = (int64_t)(int32_t)le32(d + at) - dw[1];
            if (y0 < 0 || y0 >= H || y0 % lines != 0 || seen[(size_t)(y0 / lines)]++) distinct = false;
        }
        if (distinct) nthreads = (int)std::min<int64_t>(std::min<int64_t>(nchunks, 16), std::max(1u, std::thread::hardware_concurrency()));
    }
    std::atomic<int64_t> next{0};
    std::atomic<const char*> failure{nullptr};
    auto worker = [&] {
        try {
            std::vector<uint8_t> raw;
            for (;;) {
                const int64_t k = next.fetch_add(1);
                if (k >= nchunks || failure.load() != nullptr) break;
                if (const char* why = decode_chunk(k, raw)) { failure.store(why); break; }
            }
        } catch (const std::exception&) { failure.store("out of memory while decoding an OpenEXR chunk"); }
    };
    {
        std::vector<std::thread> th;
        try { for (int t = 1; t < nthreads; ++t) th.emplace_back(worker); } catch (const std::exception&) {}   // fewer threads: still correct
        worker();
        for (auto& t : th) t.join();
    }
    if (const char* why = failure.load()) { err = why; return 1; }
    return 0;
}

// ---- Radiance RGBE (.hdr) -----------------------------------------------------------------------------------------------
// Text header lines up to an empty line, then "-Y <height> +X <width>", then per scan-line either four bytes (2, 2, w >> 8,
// w & 255) followed by the four channels run-length encoded one after the other (count > 128: a run of count - 128 copies
// of the next byte; else count literal bytes), or — widths below 8 or above 32767, and old files — flat RGBE pixels, in
// which (1, 1, 1, k) repeats the previous pixel k << shift times (the 1991 run-length form).
int decode_hdr(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    if (sniff_image(d, n) != IMG_HDR) { err = "not a Radiance HDR file"; return 1; }
    size_t p = 0; bool blank = false;
    while (p < n && !blank) {                                       // header: lines until an empty one
        size_t e = p; while (e < n && d[e] != '\n') ++e;
        if (e >= n) { err = "truncated Radiance header"; return 1; }
        const std::string line((const char*)d + p, e - p);
        if (line.empty()) blank = true;
        else if (line.compare(0, 7, "FORMAT=") == 0) {
            if (line != "FORMAT=32-bit_rle_rgbe") { err = "unsupported Radiance pixel format (only 32-bit_rle_rgbe)"; return 1; }
        }                                                           // files without a FORMAT line are RGBE by default
        p = e + 1;
    }
    if (!blank) { err = "truncated Radiance header"; return 1; }
    size_t e = p; while (e < n && d[e] != '\n') ++e;
    if (e >= n) { err = "truncated Radiance header"; return 1; }
    const std::string res((const char*)d + p, e - p); p = e + 1;
    long H = 0, W = 0; char ys = 0, xs = 0, ya = 0, xa = 0;
    if (std::sscanf(res.c_str(), "%c%c %ld %c%c %ld", &ys, &ya, &H, &xs, &xa, &W) != 6 || ya != 'Y' || xa != 'X' || ys != '-' || xs != '+') {
        err = "unsupported Radiance orientation (only -Y h +X w)"; return 1;
    }
    if (H <= 0 || W <= 0 || H > (1l << 30) || W > (1l << 30) || (uint64_t)H * (uint64_t)W > (1ull << 30)) { err = "bad Radiance image size"; return 1; }
    // a scan-line costs at least 4 bytes however well it compresses (run-length: 8 per 127 pixels and channel): a header that
    // promises more rows than the file has bytes for is refused before the image is allocated
    if ((uint64_t)H * 4u > n - p || (uint64_t)W * (uint64_t)H / 60u > n - p) { err = "Radiance file too short for its size"; return 1; }
    out.width = (uint32_t)W; out.height = (uint32_t)H; out.format = PTC_TEXEL_RGBA32F;
    out.texels.assign((size_t)W * (size_t)H * 16, 0);
    float* px = reinterpret_cast<float*>(out.texels.data());
    std::vector<uint8_t> line((size_t)W * 4);
    for (long y = 0; y < H; ++y) {
        if (n - p >= 4 && W >= 8 && W <= 32767 && d[p] == 2 && d[p + 1] == 2 && !(d[p + 2] & 0x80)) {        // new run-length form
            if ((((long)d[p + 2]) << 8 | d[p + 3]) != W) { err = "Radiance scan-line length differs from the image width"; return 1; }
            p += 4;
            for (int c = 0; c < 4; ++c) {
                long x = 0;
                while (x < W) {
                    if (p >= n) { err = "truncated Radiance scan-line"; return 1; }
                    int count = d[p++];
                    if (count > 128) {
                        count -= 128;
                        if (p >= n || x + count > W) { err = "bad Radiance run"; return 1; }
                        const uint8_t v = d[p++];
                        for (int k = 0; k < count; ++k) line[(size_t)(x + k) * 4 + c] = v;
                    } else {
                        if (count == 0 || (size_t)count > n - p || x + count > W) { err = "bad Radiance run"; return 1; }
                        for (int k = 0; k < count; ++k) line[(size_t)(x + k) * 4 + c] = d[p + k];
                        p += (size_t)count;
                    }
                    x += count;
                }
            }
        } else {                                                    // flat pixels, with the old repeat markers
            long x = 0; int shift = 0;
            while (x < W) {
                if (n - p < 4) { err = "truncated Radiance scan-line"; return 1; }
                const uint8_t* q = d + p; p += 4;
                if (q[0] == 1 && q[1] == 1 && q[2] == 1) {
                    if (x == 0 && y == 0) { err = "bad Radiance repeat marker"; return 1; }
                    if (shift > 24) { err = "bad Radiance repeat marker"; return 1; }
                    const long rep = (long)q[3] << shift;
                    if (rep <= 0 || x + rep > W) { err = "bad Radiance repeat marker"; return 1; }
                    uint8_t prev[4];
                    if (x > 0) std::memcpy(prev, &line[(size_t)(x - 1) * 4], 4);
                    else {                                          // repeats the last pixel of the previous scan-line: re-encode is not
                        err = "bad Radiance repeat marker"; return 1;   // possible from floats, and no writer emits it
                    }
                    for (long k = 0; k < rep; ++k) std::memcpy(&line[(size_t)(x + k) * 4], prev, 4);
                    x += rep; shift += 8;
                } else { std::memcpy(&line[(size_t)x * 4], q, 4); ++x; shift = 0; }
            }
        }
        float* o = px + (size_t)y * (size_t)W * 4;
        for (long x = 0; x < W; ++x) {
            const uint8_t* q = &line[(size_t)x * 4];
            const float f = q[3] ? std::ldexp(1.0f, (int)q[3] - 136) : 0.0f;
            o[x * 4] = (float)q[0] * f; o[x * 4 + 1] = (float)q[1] * f; o[x * 4 + 2] = (float)q[2] * f; o[x * 4 + 3] = 1.0f;
        }
    }
    return 0;
}

// ---- JPEG through nvJPEG (bound at run time) ---------------------------------------------------------------------------
namespace {
struct NvjpegApi {
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*StateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*StateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
    bool ok = false; std::string why;
};
const NvjpegApi& nvjpeg_api() {
    static NvjpegApi api; static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12"})
            if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
        if (!h) { api.why = "libnvjpeg.so.12 not found (JPEG textures are decoded by nvJPEG)"; return; }
        api.CreateSimple = (decltype(api.CreateSimple))dlsym(h, "nvjpegCreateSimple");
        api.Destroy = (decltype(api.Destroy))dlsym(h, "nvjpegDestroy");
        api.StateCreate = (decltype(api.StateCreate))dlsym(h, "nvjpegJpegStateCreate");
        api.StateDestroy = (decltype(api.StateDestroy))dlsym(h, "nvjpegJpegStateDestroy");
        api.GetImageInfo = (decltype(api.GetImageInfo))dlsym(h, "nvjpegGetImageInfo");
        api.Decode = (decltype(api.Decode))dlsym(h, "nvjpegDecode");
        api.ok = api.CreateSimple && api.Destroy && api.StateCreate && api.StateDestroy && api.GetImageInfo && api.Decode;
        if (!api.ok) api.why = "libnvjpeg lacks an expected entry point";
    });
    return api;
}
}  // namespace

JpegDecoder::JpegDecoder(int cuda_device) : device_(cuda_device) {
    const NvjpegApi& A = nvjpeg_api();
    if (!A.ok) { why_ = A.why; return; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); why_ = "no CUDA device: JPEG textures are decoded on the GPU (nvJPEG)"; return; }
    if (cuda_device < 0 || cuda_device >= ndev || cudaSetDevice(cuda_device) != cudaSuccess) { why_ = "bad CUDA device for the JPEG decoder"; return; }
    nvjpegHandle_t h = nullptr; nvjpegJpegState_t st = nullptr; cudaStream_t s = nullptr;
    if (A.CreateSimple(&h) != NVJPEG_STATUS_SUCCESS) { why_ = "nvjpegCreateSimple failed"; return; }
    if (A.StateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) { A.Destroy(h); why_ = "nvjpegJpegStateCreate failed"; return; }
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { A.StateDestroy(st); A.Destroy(h); why_ = "stream creation failed"; return; }
    handle_ = h; state_ = st; stream_ = s; ok_ = true;
}
JpegDecoder::~JpegDecoder() {
    const NvjpegApi& A = nvjpeg_api();
    if (ok_) { cudaSetDevice(device_); cudaFree(d_buf_); A.StateDestroy((nvjpegJpegState_t)state_); A.Destroy((nvjpegHandle_t)handle_); cudaStreamDestroy((cudaStream_t)stream_); }
}
int JpegDecoder::decode(const uint8_t* data, size_t n, DecodedImage& out, std::string& err) {
    if (!ok_) { err = why_; return 1; }
    const NvjpegApi& A = nvjpeg_api();
    cudaSetDevice(device_);
    int ncomp = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0}; nvjpegChromaSubsampling_t sub;
    if (A.GetImageInfo((nvjpegHandle_t)handle_, data, n, &ncomp, &sub, w, h) != NVJPEG_STATUS_SUCCESS || w[0] <= 0 || h[0] <= 0) { err = "nvjpegGetImageInfo failed (not a JPEG nvJPEG can read)"; return 1; }
    const size_t W = (size_t)w[0], H = (size_t)h[0], bytes = W * H * 3;
    if (d_bytes_ < bytes) {
        cudaFree(d_buf_); d_buf_ = nullptr; d_bytes_ = 0;
        if (cudaMalloc(&d_buf_, bytes) != cudaSuccess) { cudaGetLastError(); err = "JPEG decode: device allocation failed"; return 1; }
        d_bytes_ = bytes;
    }
    nvjpegImage_t img; std::memset(&img, 0, sizeof(img));
    img.channel[0] = d_buf_; img.pitch[0] = W * 3;
    // interleaved RGB, grey images replicated, as image::open(..).to_rgba32f() sees them
    const nvjpegStatus_t rc = A.Decode((nvjpegHandle_t)handle_, (nvjpegJpegState_t)state_, data, n, NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream_);
    if (rc != NVJPEG_STATUS_SUCCESS) { err = "nvjpegDecode failed (status " + std::to_string((int)rc) + ")"; return 1; }
    std::vector<uint8_t> rgb(bytes);
    if (cudaMemcpyAsync(rgb.data(), d_buf_, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream_) != cudaSuccess || cudaStreamSynchronize((cudaStream_t)stream_) != cudaSuccess) {
        err = std::string("JPEG decode: ") + cudaGetErrorString(cudaGetLastError()); return 1;
    }
    out.width = (uint32_t)W; out.height = (uint32_t)H; out.format = PTC_TEXEL_RGBA8;
    out.texels.resize(W * H * 4);
    for (size_t i = 0; i < W * H; ++i) { out.texels[i * 4] = rgb[i * 3]; out.texels[i * 4 + 1] = rgb[i * 3 + 1]; out.texels[i * 4 + 2] = rgb[i * 3 + 2]; out.texels[i * 4 + 3] = 255; }
    return 0;
}

}  // namespace ptc
