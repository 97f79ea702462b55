// fuzz_obj.cpp — mutation fuzzer for the OBJ / MTL ingest (csrc/pt_host.cpp: parse_obj, ingest; reference Scene::load_obj,
// src/scene.rs:124-257 over tobj 4.0.3), built with AddressSanitizer + UBSan by tools/fuzz.sh.  Test infrastructure.
//
// Every mutated text goes through ptc_scene_load_obj_text.  Whatever the parser accepts must be a scene ptc_create can take
// without reading out of bounds, so on success the view is checked: indices inside the vertex array, meshes inside the index
// and vertex arrays and disjoint in order, material and texture indices in range, index counts multiples of three.  On failure
// the scene must be exactly as it was (the ingest is transactional) and a message must be set.
// usage: fuzz_obj <iterations> <seed>
#include "../include/ptcore.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace ptc {
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }          // pt_step.cu defines it in the library; this build has no pt_step
}

namespace {

struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

const char* kObjSeeds[] = {
    // quads, negative indices, all four face-vertex forms, groups and objects, two materials
    "mtllib a.mtl\no first\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl red\nf 1/1/1 2/2/1 3/3/1 4/4/1\n"
    "g second\nusemtl glass\nv 0 0 1\nv 1 0 1\nv 1 1 1\nf -3/1/1 -2/2/1 -1/3/1\nf 5/1/1 6/2/1 7/3/1 1/4/1 2/1/1\n",
    "mtllib a.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nusemtl red\nf 1 2 3\nf 1 3 4\no b\nf 2 3 4\nusemtl lamp\nf 1 2 4\nl 1 2\np 3\n",
    "mtllib a.mtl\nv 0 0 0 1\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvn 0 1 0\nusemtl lamp\nf 1//1 2//1 3//2\nf 3//2 2//1 1//1\n# comment\n\ns off\n",
    "mtllib a.mtl\r\no crlf\r\nv 0.5 -1e-3 +2.25E2\r\nv 1 0 0\r\nv 0 1 0\r\nvt 0.5 0.5 0\r\nusemtl red\r\nf 1/1 2/1 3/1\r\n",
    "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n",
};
const char* kMtlSeeds[] = {
    "newmtl red\nKd 0.8 0.1 0.1\nKa 0 0 0\nKs 0.5 0.5 0.5\nNs 96\nNi 1.45\nd 1\nillum 2\nmap_Kd tex/red.png\nmap_Bump tex/red_n.png\nPr 0.3\nPm 0.1\n"
    "newmtl glass\nKd 1 1 1\nTf 0.9 0.9 0.9\nNi 1.5\nPr 0.01\nmap_Pr tex/rough.png\nmap_Pm tex/metal.png\nPs 0.2\naniso 0.3\nanisor 0.1\nPc 0.5\nPcr 0.2\n"
    "newmtl lamp\nKd 0 0 0\nKe 5 7 15\nmap_Ke tex/glow.png\nmap_Ks tex/spec.png\nmap_d tex/alpha.png\nmap_Ns tex/ns.png\nmap_Ka tex/amb.png\nbump tex/b.png\n",
    "newmtl red\nKd 1 0 0\nnewmtl glass\nTf 1 1 1\nnewmtl lamp\nKe 1 1 1\n",
};
const char* kTokens[] = {"v", "vn", "vt", "f", "o", "g", "usemtl", "mtllib", "newmtl", "Kd", "Ke", "Tf", "Pr", "Pm", "Ni", "map_Kd", "map_Bump", "s", "l", "p",
    "0", "1", "-1", "2", "3", "4", "5", "-0", "+1", "1e38", "1e39", "-1e-46", "nan", "inf", "-inf", "0x10", "1.", ".5", "1e", "4294967295", "4294967296", "2147483648",
    "-2147483649", "9223372036854775807", "18446744073709551616", "99999999999999999999999", "/", "//", "1/", "1//", "/1", "1/2/3/4", "1/-1/1", "-1/-1/-1", "0/0/0",
    " ", "\t", "\n", "\r\n", "\\\n", "#", "red", "glass", "lamp", "nosuch", ""};

std::string mutate(const std::string& in, Rng& r) {
    std::string s = in;
    const uint32_t rounds = 1 + r.below(4);
    for (uint32_t k = 0; k < rounds; ++k) {
        const size_t at = r.below((uint32_t)s.size() + 1);
        switch (r.below(9)) {
            case 0: if (at < s.size()) s[at] = (char)(32 + r.below(96)); break;
            case 1: if (at < s.size()) s.erase(at, 1 + r.below(8)); break;
            case 2: s.insert(at, kTokens[r.below(sizeof(kTokens) / sizeof(kTokens[0]))]); break;
            case 3: { const char* t = kTokens[r.below(sizeof(kTokens) / sizeof(kTokens[0]))]; s.insert(at, std::string(" ") + t + " "); } break;
            case 4: {                                                       // duplicate a line somewhere else
                const size_t b = s.rfind('\n', at ? at - 1 : 0), e = s.find('\n', at);
                const std::string line = s.substr(b == std::string::npos ? 0 : b + 1, (e == std::string::npos ? s.size() : e + 1) - (b == std::string::npos ? 0 : b + 1));
                s.insert(r.below((uint32_t)s.size() + 1), line);
            } break;
            case 5: {                                                       // replace a number by a hostile one
                size_t p = at; while (p < s.size() && !(s[p] >= '0' && s[p] <= '9')) ++p;
                size_t q = p; while (q < s.size() && ((s[q] >= '0' && s[q] <= '9') || s[q] == '.' || s[q] == '-')) ++q;
                if (p < s.size()) s.replace(p, q - p, kTokens[20 + r.below(35)]);
            } break;
            case 6: s.resize(at); break;                                    // truncate (also mid-line, mid-token)
            case 7: if (at < s.size()) s[at] = (char)r.next(); break;       // any byte, including NUL and > 127
            default: if (at < s.size()) s.insert(at, 1 + r.below(40), s[at]); break;
        }
    }
    return s;
}

#define FAIL(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n--- OBJ ---\n%s\n--- MTL ---\n%s\n", obj.c_str(), mtl.c_str()); return 1; } while (0)

struct Sizes { uint64_t v, i, m, k, t; };
Sizes sizes_of(ptc_scene* sc) {
    // textures added by a load are file-backed and undecoded, so get_view refuses them: give each one a texel first
    const uint64_t nt = ptc_scene_num_textures(sc);
    static const float one[4] = {1, 1, 1, 1};
    for (uint64_t t = 0; t < nt; ++t) { ptc_scene_texture_info info; if (ptc_scene_texture_info_at(sc, t, &info) == 0 && info.source == PTC_TEX_FILE && info.width == 0) ptc_scene_set_texels(sc, t, 1, 1, PTC_TEXEL_RGBA32F, one); }
    ptc_scene_view v; std::memset(&v, 0, sizeof(v));
    if (ptc_scene_get_view(sc, &v) != 0) return {~0ull, 0, 0, 0, 0};
    return {v.num_vertices, v.num_indices, v.num_meshes, v.num_materials, v.num_textures};
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <iterations> <seed>\n", argv[0]); return 2; }
    const long iters = std::atol(argv[1]);
    Rng r{(uint64_t)std::strtoull(argv[2], nullptr, 0) * 2654435761ull + 1};
    long ok = 0, bad = 0; uint64_t tris = 0;
    for (long it = 0; it < iters; ++it) {
        std::string obj = kObjSeeds[r.below(sizeof(kObjSeeds) / sizeof(kObjSeeds[0]))], mtl = kMtlSeeds[r.below(sizeof(kMtlSeeds) / sizeof(kMtlSeeds[0]))];
        const uint32_t which = r.below(4);
        if (which != 1) obj = mutate(obj, r);
        if (which != 0) mtl = mutate(mtl, r);
        // NUL-terminated C strings are the interface: a NUL inside simply ends the text there
        ptc_scene* sc = nullptr;
        if (ptc_scene_new("env.pfm", &sc) != 0 || !sc) FAIL("ptc_scene_new failed");
        if (r.below(4) == 0) ptc_scene_set_fix_normals(sc, 1);
        const int loads = 1 + (int)r.below(2);                             // a second load appends to the first (scene.rs:124: load_obj on a scene that has meshes)
        for (int l = 0; l < loads; ++l) {
            const Sizes before = sizes_of(sc);
            ptc::g_err.clear();
            const bool with_mtl = r.below(8) != 0;
            const int rc = ptc_scene_load_obj_text(sc, obj.c_str(), with_mtl ? mtl.c_str() : nullptr, r.below(2) ? "some/dir" : nullptr);
            const Sizes after = sizes_of(sc);
            if (after.v == ~0ull) FAIL("iteration %ld: ptc_scene_get_view failed after a load (rc %d): %s", it, rc, ptc::g_err.c_str());
            if (rc != 0) {
                if (ptc::g_err.empty()) FAIL("iteration %ld: failure %d without a message", it, rc);
                if (after.v != before.v || after.i != before.i || after.m != before.m || after.k != before.k || after.t != before.t)
                    FAIL("iteration %ld: a failed load changed the scene (%llu/%llu/%llu/%llu/%llu -> %llu/%llu/%llu/%llu/%llu): %s", it,
                         (unsigned long long)before.v, (unsigned long long)before.i, (unsigned long long)before.m, (unsigned long long)before.k, (unsigned long long)before.t,
                         (unsigned long long)after.v, (unsigned long long)after.i, (unsigned long long)after.m, (unsigned long long)after.k, (unsigned long long)after.t, ptc::g_err.c_str());
                ++bad; continue;
            }
            ++ok;
            ptc_scene_view v; std::memset(&v, 0, sizeof(v));
            if (ptc_scene_get_view(sc, &v) != 0) FAIL("iteration %ld: get_view failed", it);
            uint64_t next_index = 0, next_vertex = 0;
            for (uint64_t m = 0; m < v.num_meshes; ++m) {
                const ptc_mesh& M = v.meshes[m];
                if (M.index_count % 3 != 0) FAIL("iteration %ld: mesh %llu has %u indices", it, (unsigned long long)m, M.index_count);
                if (M.index_offset != next_index || M.vertex_offset != next_vertex) FAIL("iteration %ld: mesh %llu does not follow its predecessor", it, (unsigned long long)m);
                if ((uint64_t)M.index_offset + M.index_count > v.num_indices || (uint64_t)M.vertex_offset + M.vertex_count > v.num_vertices) FAIL("iteration %ld: mesh %llu leaves the arrays", it, (unsigned long long)m);
                if (M.material_index >= v.num_materials) FAIL("iteration %ld: mesh %llu names material %u of %llu", it, (unsigned long long)m, M.material_index, (unsigned long long)v.num_materials);
                for (uint32_t k = 0; k < M.index_count; ++k) {
                    const uint32_t ix = v.indices[M.index_offset + k];
                    if (ix < M.vertex_offset || ix >= M.vertex_offset + M.vertex_count) FAIL("iteration %ld: mesh %llu index %u = %u outside its vertices [%u, %u)", it, (unsigned long long)m, k, ix, M.vertex_offset, M.vertex_offset + M.vertex_count);
                }
                next_index += M.index_count; next_vertex += M.vertex_count; tris += M.index_count / 3;
            }
            if (next_index != v.num_indices || next_vertex != v.num_vertices) FAIL("iteration %ld: %llu indices / %llu vertices belong to no mesh", it, (unsigned long long)(v.num_indices - next_index), (unsigned long long)(v.num_vertices - next_vertex));
            for (uint64_t k = 0; k < v.num_materials; ++k) {
                const ptc_material& K = v.materials[k];
                const uint32_t idx[9] = {K.albedo_texture_idx, K.roughness_texture_idx, K.metallic_texture_idx, K.emissive_texture_idx, K.specular_texture_idx, K.sheen_texture_idx, K.normal_texture_idx, K.transmission_texture_idx, K.ior_texture_idx};
                for (uint32_t t : idx) if (t >= v.num_textures) FAIL("iteration %ld: material %llu names texture %u of %llu", it, (unsigned long long)k, t, (unsigned long long)v.num_textures);
            }
        }
        ptc_scene_free(sc);
    }
    std::printf("fuzz_obj: %ld loads accepted (%llu triangles, every scene invariant held), %ld refused with a message and no change to the scene, 0 sanitizer reports\n", ok, (unsigned long long)tris, bad);
    return 0;
}
