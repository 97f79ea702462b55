// fuzz_decode.cpp — mutation fuzzer for the library's image decoders (csrc/img_decode.cu: PNG, OpenEXR, Radiance HDR, PFM), built with
// AddressSanitizer + UBSan by tools/fuzz.sh.  Test infrastructure, not part of libptcore.
//
// Scene files are untrusted input (Texture::load decodes whatever the MTL names, reference src/texture.rs:232-243), so a
// damaged file must come back as an error string, never as a crash, an out-of-bounds access or an unbounded allocation.
// Mutations are structure-aware where a checksum or a compressor would otherwise stop them at the door:
//   PNG : chunks are re-framed with a fresh CRC after the body is mutated; IDAT is inflated, mutated (filter bytes, pixel
//         bytes, length) and deflated again half of the time, and mutated in its compressed form the other half
//   EXR : header attributes, the offset table, chunk headers and chunk payloads are mutated in place (no checksums);
//         ZIP / ZIPS payloads are also re-deflated from a mutated raw block so the predictor / interleave code sees them
//   HDR : header and resolution line (hostile sizes), scan-line markers, run counts
//   PFM : text header and payload length
// usage: fuzz_decode <iterations> <seed> file...      exit code 0 = no sanitizer report (a report aborts the process)
#include "../wgpu-pathtracer_b200/csrc/img_decode.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

namespace {

typedef std::vector<uint8_t> Bytes;

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(Bytes& b, uint32_t v) { for (int k = 3; k >= 0; --k) b.push_back((uint8_t)(v >> (8 * k))); }

// interesting 32-bit values for size / offset / dimension fields
uint32_t magic32(Rng& r) {
    static const uint32_t M[] = {0, 1, 2, 7, 8, 9, 15, 16, 17, 255, 256, 257, 4095, 4096, 65535, 65536, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu, 0x00ffffffu, 0x01000000u, 0x3fffffffu, 0x40000000u};
    return M[r.below(sizeof(M) / sizeof(M[0]))];
}

void mutate_bytes(Bytes& b, Rng& r, size_t lo = 0, size_t hi = (size_t)-1) {
    if (hi > b.size()) hi = b.size();
    if (lo >= hi) return;
    const uint32_t rounds = 1 + r.below(4);
    for (uint32_t k = 0; k < rounds; ++k) {
        const size_t at = lo + r.below((uint32_t)(hi - lo));
        switch (r.below(7)) {
            case 0: b[at] ^= (uint8_t)(1u << r.below(8)); break;
            case 1: b[at] = (uint8_t)r.next(); break;
            case 2: b[at] = (uint8_t)(r.below(2) ? 0xff : 0x00); break;
            case 3: if (at + 4 <= hi) { const uint32_t v = magic32(r); const bool be = r.below(2); for (int j = 0; j < 4; ++j) b[at + j] = (uint8_t)(v >> (8 * (be ? 3 - j : j))); } break;
            case 4: if (at + 2 <= hi) { b[at] = (uint8_t)r.next(); b[at + 1] = (uint8_t)r.next(); } break;
            case 5: { const size_t len = 1 + r.below(16); for (size_t j = at; j < hi && j < at + len; ++j) b[j] = (uint8_t)r.next(); } break;
            default: { const size_t len = 1 + r.below(64); const uint8_t v = (uint8_t)r.next(); for (size_t j = at; j < hi && j < at + len; ++j) b[j] = v; } break;
        }
    }
}

void resize_random(Bytes& b, Rng& r) {
    switch (r.below(4)) {
        case 0: b.resize(r.below((uint32_t)b.size() + 1)); break;                          // truncate
        case 1: b.resize(b.size() + 1 + r.below(64), (uint8_t)r.next()); break;            // grow
        case 2: if (!b.empty()) b.erase(b.begin() + r.below((uint32_t)b.size())); break;   // drop one byte
        default: b.insert(b.begin() + r.below((uint32_t)b.size() + 1), (uint8_t)r.next()); break;
    }
}

bool inflate_all(const Bytes& in, Bytes& out) {
    out.clear();
    z_stream zs; std::memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(in.data()); zs.avail_in = (uInt)in.size();
    uint8_t buf[1 << 14]; int rc = Z_OK;
    while (rc == Z_OK) {
        zs.next_out = buf; zs.avail_out = sizeof(buf);
        rc = inflate(&zs, Z_NO_FLUSH);
        out.insert(out.end(), buf, buf + (sizeof(buf) - zs.avail_out));
        if (out.size() > (64u << 20)) break;
    }
    inflateEnd(&zs);
    return rc == Z_STREAM_END;
}
Bytes deflate_all(const Bytes& in, Rng& r) {
    uLongf cap = compressBound((uLong)in.size());
    Bytes out(cap);
    compress2(out.data(), &cap, in.data(), (uLong)in.size(), (int)r.below(10));
    out.resize(cap);
    return out;
}

// ---- PNG ----------------------------------------------------------------------------------------------------------------
struct Chunk { char type[4]; Bytes body; };
bool png_split(const Bytes& f, std::vector<Chunk>& out) {
    size_t p = 8;
    while (p + 12 <= f.size()) {
        const uint32_t len = be32(&f[p]);
        if (len > f.size() - p - 12) return false;
        Chunk c; std::memcpy(c.type, &f[p + 4], 4); c.body.assign(f.begin() + p + 8, f.begin() + p + 8 + len);
        out.push_back(c); p += 12 + len;
    }
    return !out.empty();
}
Bytes png_join(const std::vector<Chunk>& cs, Rng& r, bool break_crc) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    Bytes f(sig, sig + 8);
    for (const Chunk& c : cs) {
        put_be32(f, (uint32_t)c.body.size());
        f.insert(f.end(), c.type, c.type + 4);
        f.insert(f.end(), c.body.begin(), c.body.end());
        uint32_t crc = (uint32_t)crc32(crc32(0L, (const Bytef*)c.type, 4), c.body.data(), (uInt)c.body.size());
        if (break_crc && r.below(8) == 0) crc ^= 1u << r.below(32);
        put_be32(f, crc);
    }
    return f;
}
Bytes mutate_png(const Bytes& seed, Rng& r) {
    std::vector<Chunk> cs;
    if (!png_split(seed, cs)) { Bytes b = seed; mutate_bytes(b, r); return b; }
    const uint32_t edits = 1 + r.below(3);
    for (uint32_t e = 0; e < edits; ++e) {
        Chunk& c = cs[r.below((uint32_t)cs.size())];
        const bool idat = !std::memcmp(c.type, "IDAT", 4), ihdr = !std::memcmp(c.type, "IHDR", 4);
        const uint32_t what = r.below(10);
        if (ihdr && what < 7 && c.body.size() == 13) {
            switch (r.below(6)) {
                case 0: { const uint32_t v = r.below(2) ? magic32(r) : 1 + r.below(64); for (int j = 0; j < 4; ++j) c.body[j] = (uint8_t)(v >> (8 * (3 - j))); } break;
                case 1: { const uint32_t v = r.below(2) ? magic32(r) : 1 + r.below(64); for (int j = 0; j < 4; ++j) c.body[4 + j] = (uint8_t)(v >> (8 * (3 - j))); } break;
                case 2: { static const uint8_t D[] = {1, 2, 4, 8, 16, 0, 3, 32}; c.body[8] = D[r.below(8)]; } break;
                case 3: { static const uint8_t T[] = {0, 2, 3, 4, 6, 1, 5, 7}; c.body[9] = T[r.below(8)]; } break;
                case 4: c.body[12] = (uint8_t)r.below(3); break;
                default: mutate_bytes(c.body, r); break;
            }
        } else if (idat && what < 5) {
            Bytes raw;
            Bytes all;                                            // a stream split over several IDATs inflates as one
            for (const Chunk& k : cs) if (!std::memcmp(k.type, "IDAT", 4)) all.insert(all.end(), k.body.begin(), k.body.end());
            if (inflate_all(all, raw)) {
                if (r.below(3) == 0) resize_random(raw, r); else mutate_bytes(raw, r);
                if (r.below(4) == 0 && !raw.empty()) raw[0] = (uint8_t)r.below(8);                // filter type incl. invalid ones
                Bytes z = deflate_all(raw, r);
                bool first = true;
                for (size_t k = 0; k < cs.size();) {
                    if (std::memcmp(cs[k].type, "IDAT", 4)) { ++k; continue; }
                    if (first) { cs[k].body = z; first = false; ++k; } else cs.erase(cs.begin() + k);
                }
                continue;                                         // `c` may dangle after the erase
            }
            mutate_bytes(c.body, r);
        } else if (what == 7) {
            cs.erase(cs.begin() + r.below((uint32_t)cs.size()));  // drop a chunk
            if (cs.empty()) break;
        } else if (what == 8) {
            const Chunk dup = cs[r.below((uint32_t)cs.size())];   // duplicate / reorder a chunk
            cs.insert(cs.begin() + r.below((uint32_t)cs.size() + 1), dup);
        } else if (what == 9) resize_random(c.body, r);
        else mutate_bytes(c.body, r);
    }
    Bytes f = png_join(cs, r, true);
    if (r.below(16) == 0) resize_random(f, r);
    return f;
}

// ---- EXR ----------------------------------------------------------------------------------------------------------------
uint32_t le32(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint64_t le64(const uint8_t* p) { return le32(p) | ((uint64_t)le32(p + 4) << 32); }
size_t exr_header_end(const Bytes& f) {                           // offset of the byte after the header's terminating 0
    size_t p = 8;
    while (p < f.size() && f[p] != 0) {
        while (p < f.size() && f[p]) ++p; ++p;                    // name
        while (p < f.size() && f[p]) ++p; ++p;                    // type
        if (p + 4 > f.size()) return f.size();
        const uint32_t sz = le32(&f[p]); p += 4;
        if (sz > f.size() - p) return f.size();
        p += sz;
    }
    return p + 1 <= f.size() ? p + 1 : f.size();
}
Bytes mutate_exr(const Bytes& seed, Rng& r) {
    Bytes f = seed;
    const size_t hend = exr_header_end(f);
    switch (r.below(8)) {
        case 0: case 1: mutate_bytes(f, r, 0, hend); break;       // magic, version flags, attributes (channels, windows, compression)
        case 2: mutate_bytes(f, r, hend, hend + 64 < f.size() ? hend + 64 : f.size()); break;   // offset table
        case 3: {                                                 // a chunk header: y, payload size
            if (hend + 8 <= f.size()) {
                const uint64_t at = le64(&f[hend]);
                if (at + 8 <= f.size()) mutate_bytes(f, r, (size_t)at, (size_t)at + 8);
            }
        } break;
        case 4: {                                                 // re-deflate a mutated raw block into the first chunk (ZIP / ZIPS)
            if (hend + 8 <= f.size()) {
                const uint64_t at = le64(&f[hend]);
                if (at + 8 <= f.size()) {
                    const uint32_t csz = le32(&f[at + 4]);
                    if (csz <= f.size() - at - 8) {
                        Bytes z(f.begin() + at + 8, f.begin() + at + 8 + csz), raw;
                        if (inflate_all(z, raw)) {
                            if (r.below(2)) resize_random(raw, r); else mutate_bytes(raw, r);
                            Bytes nz = deflate_all(raw, r);
                            if (nz.size() <= csz) {               // keep the file layout: pad, and say the new size
                                std::memcpy(&f[at + 8], nz.data(), nz.size());
                                const uint32_t v = (uint32_t)nz.size(); for (int j = 0; j < 4; ++j) f[at + 4 + j] = (uint8_t)(v >> (8 * j));
                                break;
                            }
                        }
                    }
                }
            }
            mutate_bytes(f, r, hend);
        } break;
        case 5: resize_random(f, r); break;
        default: mutate_bytes(f, r, hend); break;                 // payloads: PIZ bitmap / Huffman table / code stream, raw rows
    }
    return f;
}

Bytes mutate_pfm(const Bytes& seed, Rng& r) {
    Bytes f = seed;
    if (r.below(3) == 0) resize_random(f, r);
    else if (r.below(2)) mutate_bytes(f, r, 0, 16 < f.size() ? 16 : f.size());
    else mutate_bytes(f, r);
    if (r.below(8) == 0) {                                         // huge declared sizes
        const char* H[] = {"PF\n99999999 99999999\n-1\n", "PF\n4294967295 2\n-1\n", "Pf\n-5 3\n1\n", "PF\n0 0\n-1\n", "PF\n1 1\nnan\n", "PF\n70000 70000\n-1.0\n"};
        const char* h = H[r.below(6)];
        f.assign(h, h + std::strlen(h)); f.resize(f.size() + r.below(64), 0x3f);
    }
    return f;
}

// Radiance HDR: text header (FORMAT / resolution line) and run-length encoded scan-lines, no checksums
Bytes mutate_hdr(const Bytes& seed, Rng& r) {
    Bytes f = seed;
    size_t hend = 0;                                                // end of the resolution line
    for (size_t i = 1, blank = 0; i < f.size(); ++i) if (f[i] == '\n') { if (blank) { hend = i + 1; break; } if (f[i - 1] == '\n') blank = 1; }
    switch (r.below(6)) {
        case 0: mutate_bytes(f, r, 0, hend); break;
        case 1: {                                                    // hostile sizes in the resolution line
            const char* R[] = {"-Y 99999999 +X 99999999\n", "-Y 1 +X 1073741824\n", "+Y 4 +X 4\n", "-Y 0 +X 0\n", "-Y 3 +X 32768\n", "-Y -5 +X 7\n", "-Y 2 +X 8\n", "-Y 40000 +X 9\n"};
            const char* t = R[r.below(8)];
            size_t a = 0; for (size_t i = 1; i < f.size(); ++i) if (f[i] == '\n' && f[i - 1] == '\n') { a = i + 1; break; }
            if (a && hend > a) { f.erase(f.begin() + a, f.begin() + hend); f.insert(f.begin() + a, t, t + std::strlen(t)); }
        } break;
        case 2: resize_random(f, r); break;
        default: mutate_bytes(f, r, hend); break;                    // scan-line markers, run counts, literals
    }
    return f;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s <iterations> <seed> file...\n", argv[0]); return 2; }
    const long iters = std::atol(argv[1]);
    Rng r{(uint64_t)std::strtoull(argv[2], nullptr, 0) * 2654435761ull + 1};
    std::vector<Bytes> seeds; std::vector<ptc::ImageKind> kinds;
    for (int k = 3; k < argc; ++k) {
        std::ifstream in(argv[k], std::ios::binary);
        Bytes b((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        const ptc::ImageKind kind = ptc::sniff_image(b.data(), b.size());
        if (kind != ptc::IMG_PNG && kind != ptc::IMG_EXR && kind != ptc::IMG_PFM && kind != ptc::IMG_HDR) { std::fprintf(stderr, "skipping %s (not PNG / EXR / HDR / PFM)\n", argv[k]); continue; }
        // every seed must decode as it is: a fuzzer over files the decoder refuses explores nothing
        ptc::DecodedImage img; std::string err;
        const int rc = kind == ptc::IMG_PNG ? ptc::decode_png(b.data(), b.size(), img, err) : kind == ptc::IMG_EXR ? ptc::decode_exr(b.data(), b.size(), img, err)
                     : kind == ptc::IMG_HDR ? ptc::decode_hdr(b.data(), b.size(), img, err) : ptc::decode_pfm(b.data(), b.size(), img, err);
        if (rc != 0) { std::fprintf(stderr, "seed %s does not decode: %s\n", argv[k], err.c_str()); return 2; }
        seeds.push_back(b); kinds.push_back(kind);
    }
    if (seeds.empty()) { std::fprintf(stderr, "no usable seed\n"); return 2; }
    long ok = 0, bad = 0; uint64_t texel_bytes = 0;
    for (long it = 0; it < iters; ++it) {
        const uint32_t s = r.below((uint32_t)seeds.size());
        const Bytes f = kinds[s] == ptc::IMG_PNG ? mutate_png(seeds[s], r) : kinds[s] == ptc::IMG_EXR ? mutate_exr(seeds[s], r) : kinds[s] == ptc::IMG_HDR ? mutate_hdr(seeds[s], r) : mutate_pfm(seeds[s], r);
        // exact-size heap copy: AddressSanitizer then sees a read one byte past the file
        uint8_t* copy = (uint8_t*)std::malloc(f.size() ? f.size() : 1);
        std::memcpy(copy, f.data(), f.size());
        ptc::DecodedImage img; std::string err;
        int rc;
        switch (ptc::sniff_image(copy, f.size())) {
            case ptc::IMG_PNG: rc = ptc::decode_png(copy, f.size(), img, err); break;
            case ptc::IMG_EXR: rc = ptc::decode_exr(copy, f.size(), img, err); break;
            case ptc::IMG_PFM: rc = ptc::decode_pfm(copy, f.size(), img, err); break;
            case ptc::IMG_HDR: rc = ptc::decode_hdr(copy, f.size(), img, err); break;
            default: rc = 1; err = "unrecognised"; break;
        }
        std::free(copy);
        if (rc == 0) {
            const size_t px = (size_t)img.width * img.height, bytes = px * (img.format == 1 ? 16 : 4);
            if (img.texels.size() != bytes || px == 0) { std::fprintf(stderr, "iteration %ld: success with %zu texel bytes for %u x %u (format %u)\n", it, img.texels.size(), img.width, img.height, img.format); return 1; }
            texel_bytes += bytes; ++ok;
        } else {
            if (err.empty()) { std::fprintf(stderr, "iteration %ld: failure without a message\n", it); return 1; }
            ++bad;
        }
    }
    std::printf("fuzz_decode: %ld inputs, %ld decoded (%.1f MB of texels), %ld refused with a message, 0 sanitizer reports\n", iters, ok, texel_bytes / 1e6, bad);
    return 0;
}
