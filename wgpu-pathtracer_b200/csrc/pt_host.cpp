// pt_host.cpp — host side of the drop-in: camera math and OBJ/MTL scene ingest.
//
// Restates src/camera.rs:31-79 and src/scene.rs:77-394 of the reference, including the OBJ/MTL
// semantics of its parser dependency tobj 4.0.3 (Cargo.lock:1979-1981; not vendored in the
// reference, restated from its documented behaviour with GPU_LOAD_OPTIONS = triangulate +
// single_index + ignore_points + ignore_lines).  Texture files are recorded by path only, as the
// reference does (texture.rs:186-231: `from_file_*` store the path; decode happens in load());
// the caller decodes and hands texels in with ptc_scene_set_texels.
#include "../../include/ptcore.h"
#include "pt_error.h"
#include "img_decode.h"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using ptc::set_error;

// =============================================================================================
// camera.rs
// =============================================================================================
namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 mulf(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 divf(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 addv(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
// glam 0.30.3 Vec3 (scalar): cross and normalize (v * length_recip)
inline V3 crossv(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normv(V3 a) { float len = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return mulf(a, 1.0f / len); }
inline V3 ld(const float* p) { return {p[0], p[1], p[2]}; }
inline void st(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace

extern "C" {

PTC_API void ptc_camera_update(ptc_host_camera* c) {                       // camera.rs:54-79
    const float kPi = 3.14159265358979323846f;                             // std::f32::consts::PI
    float viewport_height = 2.0f * std::tan(c->fov * (kPi / 180.0f) / 2.0f);
    float viewport_width = viewport_height * ((float)c->width / (float)c->height);
    V3 pos = ld(c->position);
    V3 vz = normv(sub(pos, ld(c->lookat)));
    V3 vx = normv(crossv(V3{0.0f, 1.0f, 0.0f}, vz));
    V3 vy = crossv(vz, vx);
    V3 vxv = mulf(vx, viewport_width);
    V3 vyv = mulf(V3{-vy.x, -vy.y, -vy.z}, viewport_height);
    V3 psx = divf(vxv, (float)c->width), psy = divf(vyv, (float)c->height);
    V3 top_left = sub(sub(sub(pos, vz), divf(vxv, 2.0f)), divf(vyv, 2.0f));
    V3 first = addv(top_left, mulf(addv(psx, psy), 0.5f));
    st(c->pixel_space_x, psx); st(c->pixel_space_y, psy); st(c->first_pixel_pos, first);
    // the look-at point is refreshed AFTER the basis was derived from the old one (one-call lag, as in the reference)
    c->lookat[0] = c->position[0] + std::cos(c->theta_x) * std::cos(c->theta_y);
    c->lookat[1] = c->position[1] + std::sin(c->theta_y);
    c->lookat[2] = c->position[2] + std::sin(c->theta_x) * std::cos(c->theta_y);
}

PTC_API void ptc_camera_new(ptc_host_camera* c, uint32_t width, uint32_t height, float fov_deg, const float position[3]) {   // camera.rs:31-52
    std::memset(c, 0, sizeof(*c));
    c->width = width; c->height = height; c->fov = fov_deg;
    for (int a = 0; a < 3; ++a) { c->position[a] = position[a]; c->lookat[a] = position[a]; }
    c->lookat[2] = position[2] + 1.0f;
    c->samples_per_pixel = 1;   // shader_definitions::SAMPLES_PER_PIXEL
    c->max_bounces = 4;         // shader_definitions::MAX_BOUNCES
    c->camera_speed = 1.0f;
    ptc_camera_update(c);
}

PTC_API void ptc_camera_uniforms(const ptc_host_camera* c, ptc_uniforms* out) {   // uniforms.rs:30-50,64-71
    std::memset(out, 0, sizeof(*out));
    ptc_camera& u = out->camera;
    for (int a = 0; a < 3; ++a) {
        u.position[a] = c->position[a]; u.lookat[a] = c->lookat[a];
        u.pixel_space_x[a] = c->pixel_space_x[a]; u.pixel_space_y[a] = c->pixel_space_y[a]; u.first_pixel_pos[a] = c->first_pixel_pos[a];
    }
    u.width = c->width; u.height = c->height; u.fov = c->fov; u.frame = c->frame;
    u.samples_per_pixel = c->samples_per_pixel; u.max_bounces = c->max_bounces;
}

}  // extern "C"

// =============================================================================================
// scene.rs
// =============================================================================================
namespace {

struct SceneTexture {
    uint32_t source = PTC_TEX_CONSTANT;
    uint32_t width = 1, height = 1, format = PTC_TEXEL_RGBA32F;
    bool has_color = false, has_scalar = false;      // texture.rs:20-21 `color`, `scalar` (dedup keys)
    float color[4] = {0, 0, 0, 0};
    float scalar = 0.0f;
    std::string path;
    std::vector<uint8_t> texels;                      // owned copy for TEXELS / decoded FILE / constants
};

struct TobjMaterial {                                  // the subset of tobj::Material scene.rs reads
    std::string name;
    bool has_diffuse = false, has_specular = false, has_ni = false;
    float diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0}, optical_density = 0.0f;
    std::string diffuse_texture, specular_texture, normal_texture;
    bool has_diffuse_texture = false, has_specular_texture = false, has_normal_texture = false;
    std::map<std::string, std::string> unknown_param;
};

struct VertKey { size_t v, vt, vn; bool operator==(const VertKey& o) const { return v == o.v && vt == o.vt && vn == o.vn; } };
struct VertKeyHash { size_t operator()(const VertKey& k) const { return k.v * 73856093u ^ k.vt * 19349663u ^ k.vn * 83492791u; } };
const size_t kMissing = (size_t)-1;

struct TobjModel {
    std::vector<float> positions, normals, texcoords;
    std::vector<uint32_t> indices;
    bool has_material = false; size_t material_id = 0;
    std::string name;
};

// Rust's f32::from_str accepts decimal / exponent / inf / nan with optional sign and nothing else.
bool parse_f32(const std::string& s, float* out) {
    if (s.empty()) return false;
    // std::from_chars is locale-independent (strtof follows LC_NUMERIC: "0.8" would stop at the '.' under a comma locale)
    // and correctly rounded.  It takes no leading '+', Rust does; "inf" / "nan" / "infinity" in any case are accepted by both.
    const char* a = s.data(); const char* b = a + s.size();
    if (*a == '+') { ++a; if (a == b || *a == '+' || *a == '-') return false; }
    for (const char* c = a; c < b; ++c) if (*c == 'x' || *c == 'X' || *c == ' ' || *c == '\t' || *c == '(') return false;
    float v = 0.0f;
    const std::from_chars_result r = std::from_chars(a, b, v);
    if (r.ptr != b) return false;
    if (r.ec == std::errc::result_out_of_range) {
        // Rust (and strtof) round overflow to infinity and underflow to zero / a subnormal instead of failing: take the value
        // through double, whose range covers every f32 overflow / underflow a file could plausibly hold
        double dv = 0.0;
        const std::from_chars_result rd = std::from_chars(a, b, dv);
        if (rd.ptr != b) return false;
        if (rd.ec == std::errc()) v = (float)dv;
        else { const bool neg = *a == '-'; bool tiny = false; for (const char* c = a; c < b; ++c) if ((*c == 'e' || *c == 'E') && c + 1 < b && c[1] == '-') tiny = true; v = tiny ? 0.0f : HUGE_VALF; if (neg) v = -v; }
    } else if (r.ec != std::errc()) return false;
    *out = v; return true;
}

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out; std::istringstream is(s); std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
std::string parent_dir(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string("") : p.substr(0, k);
}
std::string join_path(const std::string& dir, const std::string& f) {       // Path::new(obj).parent().join(f)
    if (!f.empty() && f[0] == '/') return f;
    if (dir.empty()) return f;
    return dir + "/" + f;
}

}  // namespace

struct ptc_scene {
    std::vector<ptc_mesh> meshes;
    std::vector<ptc_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<ptc_material> materials;
    std::vector<SceneTexture> textures;
    std::vector<ptc_texture_desc> descs;   // filled by get_view
    bool fix_normals = false;
};

namespace {

void set_const(SceneTexture& t, float r, float g, float b, float a) {
    t.source = PTC_TEX_CONSTANT; t.width = t.height = 1; t.format = PTC_TEXEL_RGBA32F;
    t.color[0] = r; t.color[1] = g; t.color[2] = b; t.color[3] = a; t.has_color = true;
    t.texels.resize(16); std::memcpy(t.texels.data(), t.color, 16);
}
uint32_t push_color(ptc_scene* s, float r, float g, float b, float a) {      // Texture::from_color, texture.rs:42-59
    SceneTexture t; set_const(t, r, g, b, a); s->textures.push_back(t); return (uint32_t)s->textures.size() - 1;
}
uint32_t push_scalar(ptc_scene* s, float v) {                                 // Texture::from_scalar, texture.rs:61-76
    SceneTexture t; set_const(t, v, 0.0f, 0.0f, 1.0f); t.has_scalar = true; t.scalar = v; s->textures.push_back(t); return (uint32_t)s->textures.size() - 1;
}
uint32_t push_file(ptc_scene* s, const std::string& path) {                   // Texture::from_file_*, texture.rs:186-231
    SceneTexture t; t.source = PTC_TEX_FILE; t.width = t.height = 0; t.path = path; s->textures.push_back(t); return (uint32_t)s->textures.size() - 1;
}

// scene.rs:364-394
bool known_vector(ptc_scene* s, bool has_tex, const std::string& tex, bool has_prop, const float* prop, const std::string& dir, uint32_t* out) {
    if (has_tex) { *out = push_file(s, join_path(dir, tex)); return true; }
    if (!has_prop) return false;
    for (size_t i = 0; i < s->textures.size(); ++i) {
        const SceneTexture& t = s->textures[i];
        if (t.has_color && t.color[0] == prop[0] && t.color[1] == prop[1] && t.color[2] == prop[2] && t.color[3] == 1.0f) { *out = (uint32_t)i; return true; }
    }
    *out = push_color(s, prop[0], prop[1], prop[2], 1.0f); return true;
}
// scene.rs:331-362 (tex_prop is always None at its only call site)
bool known_scalar(ptc_scene* s, bool has_prop, float prop, uint32_t* out) {
    if (!has_prop) return false;
    for (size_t i = 0; i < s->textures.size(); ++i)
        if (s->textures[i].has_scalar && s->textures[i].scalar == prop) { *out = (uint32_t)i; return true; }
    *out = push_scalar(s, prop); return true;
}
// scene.rs:259-273
bool scalar_prop(const TobjMaterial& m, const char* key, float* out) {
    auto it = m.unknown_param.find(key);
    if (it == m.unknown_param.end()) return false;
    return parse_f32(it->second, out);
}
// returns 0 = None, 1 = Some, -1 = would panic in the reference (scene.rs:305 unwrap on a malformed vector)
int vec_prop(const TobjMaterial& m, const char* key, float out[3]) {
    auto it = m.unknown_param.find(key);
    if (it == m.unknown_param.end()) return 0;
    std::vector<float> vals; const std::string& s = it->second;
    size_t a = 0;
    for (;;) {                                                                // str::split(" ")
        size_t b = s.find(' ', a);
        std::string tok = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
        float v = 0.0f; if (!parse_f32(tok, &v)) v = 0.0f;                     // unwrap_or(0.0)
        vals.push_back(v);
        if (b == std::string::npos) break;
        a = b + 1;
    }
    if (vals.size() != 3) return -1;
    out[0] = vals[0]; out[1] = vals[1]; out[2] = vals[2];
    return 1;
}
// scene.rs:275-329; returns 1 = Some(*out), 0 = None, <0 = error
int texture_prop(ptc_scene* s, const TobjMaterial& m, const char* tex_key, const char* key, bool scalar, const std::string& dir, uint32_t* out) {
    auto tex = m.unknown_param.find(tex_key);
    if (tex != m.unknown_param.end()) { *out = push_file(s, join_path(dir, tex->second)); return 1; }
    auto it = m.unknown_param.find(key);
    if (it == m.unknown_param.end()) return 0;
    if (scalar) {
        float v; if (!scalar_prop(m, key, &v)) return 0;
        return known_scalar(s, true, v, out) ? 1 : 0;
    }
    float v[3]; int r = vec_prop(m, key, v);
    if (r < 0) { set_error(std::string("material '") + m.name + "': '" + key + "' is not three space-separated numbers (the reference panics at scene.rs:305)"); return PTC_ERR_PARSE; }
    if (std::strcmp(key, "Ke") == 0 && v[0] == 0.0f && v[1] == 0.0f && v[2] == 0.0f) return 0;
    return known_vector(s, false, "", true, v, dir, out) ? 1 : 0;
}

// ---- tobj 4.0.3: MTL ----------------------------------------------------------------------
int parse_mtl(const std::string& text, std::vector<TobjMaterial>& mats, std::unordered_map<std::string, size_t>& mat_map) {
    std::istringstream is(text); std::string raw;
    TobjMaterial cur; bool have = false;
    auto flush = [&] { if (have && !cur.name.empty()) { mat_map[cur.name] = mats.size(); mats.push_back(cur); } };
    auto f3 = [&](const std::vector<std::string>& w, float* o) {
        if (w.size() < 4) return false;
        return parse_f32(w[1], &o[0]) && parse_f32(w[2], &o[1]) && parse_f32(w[3], &o[2]);
    };
    while (std::getline(is, raw)) {
        std::string line = trim(raw);
        std::vector<std::string> w = split_ws(line);
        if (w.empty() || w[0] == "#" || w[0][0] == '#') continue;
        const std::string& k = w[0];
        if (k == "newmtl") {
            flush(); cur = TobjMaterial(); have = true;
            cur.name = trim(line.substr(6));
            if (cur.name.empty()) { set_error("MTL: newmtl without a name"); return PTC_ERR_PARSE; }
        } else if (k == "Ka") { float t[3]; if (!f3(w, t)) { set_error("MTL: bad Ka"); return PTC_ERR_PARSE; } }
        else if (k == "Kd") { if (!f3(w, cur.diffuse)) { set_error("MTL: bad Kd"); return PTC_ERR_PARSE; } cur.has_diffuse = true; }
        else if (k == "Ks") { if (!f3(w, cur.specular)) { set_error("MTL: bad Ks"); return PTC_ERR_PARSE; } cur.has_specular = true; }
        else if (k == "Ns" || k == "d") { float t; if (w.size() < 2 || !parse_f32(w[1], &t)) { set_error("MTL: bad " + k); return PTC_ERR_PARSE; } }
        else if (k == "Ni") { if (w.size() < 2 || !parse_f32(w[1], &cur.optical_density)) { set_error("MTL: bad Ni"); return PTC_ERR_PARSE; } cur.has_ni = true; }
        else if (k == "illum") { /* typed field, unused by scene.rs */ }
        else if (k == "map_Ka" || k == "map_Ns" || k == "map_ns" || k == "map_NS" || k == "map_d") { /* typed fields, unused by scene.rs */ }
        else if (k == "map_Kd") { cur.diffuse_texture = trim(line.substr(6)); cur.has_diffuse_texture = true; }
        else if (k == "map_Ks") { cur.specular_texture = trim(line.substr(6)); cur.has_specular_texture = true; }
        else if (k == "map_Bump" || k == "map_bump") { cur.normal_texture = trim(line.substr(8)); cur.has_normal_texture = true; }
        else if (k == "bump") { cur.normal_texture = trim(line.substr(4)); cur.has_normal_texture = true; }
        else cur.unknown_param[k] = trim(line.substr(k.size()));
    }
    flush();
    return PTC_OK;
}

// ---- tobj 4.0.3: OBJ ----------------------------------------------------------------------
// ---- OBJ parsing, in three passes (scene ingest at scale — the reference's tobj parse is serial) ----------
//   1. the text is cut at line ends into one chunk per host thread; every chunk is scanned in place (no per-line strings):
//      v / vt / vn values, face elements as written (1-based, or negative = relative to the count at that line), and the
//      o / g / usemtl / mtllib lines as ordered events;
//   2. one serial walk over the chunks' events turns them into models (tobj's split rules) — cheap, it touches no geometry;
//   3. the models are exported in parallel: single_index de-duplication of (v, vt, vn) triples + triangulation per model.
// The result is the same std::vector<TobjModel> the serial parser produced, in the same order.
struct RawElem { long long v, vt, vn; };                 // 0 = absent
struct RawFace { uint32_t first, count; uint32_t np, nt, nn; };   // elements [first, first+count) of the chunk; v/vt/vn counts of the chunk before this line
struct ObjEvent { int kind; uint32_t face_pos; std::string arg; };  // kind 0 = o/g, 1 = usemtl, 2 = mtllib; face_pos = faces of the chunk before it
struct ObjChunk {
    const char *begin = nullptr, *end = nullptr;
    std::vector<float> pos, tex, nrm;
    std::vector<RawElem> elems;
    std::vector<RawFace> faces;
    std::vector<ObjEvent> events;
    int err = PTC_OK; std::string msg;
};
struct ObjGlobal { std::vector<float> pos, tex, nrm; std::vector<size_t> pos_base, tex_base, nrm_base; };   // *_base[c] = elements before chunk c

inline bool is_ws(char c) { return std::isspace((unsigned char)c) != 0; }

// f32 token: the fast path (std::from_chars, correctly rounded like strtof) covers plain decimals; anything else goes through parse_f32
bool parse_f32_tok(const char* a, const char* b, float* out) {
    const char* p = a;
    if (p < b && *p == '+') ++p;
    bool plain = p < b;
    for (const char* q = p; q < b && plain; ++q) plain = (*q >= '0' && *q <= '9') || *q == '.' || *q == '-' || *q == '+' || *q == 'e' || *q == 'E';
    if (plain && !(p < b && (*p == '+'))) {
        auto r = std::from_chars(p, b, *out);
        if (r.ec == std::errc() && r.ptr == b) return true;
    }
    return parse_f32(std::string(a, b), out);
}
bool parse_index_tok(const char* a, const char* b, long long* out) {
    if (a == b) { *out = 0; return true; }
    const char* p = a;
    if (*p == '+') ++p;                                   // strtoll accepts a leading '+'
    long long v = 0;
    auto r = std::from_chars(p, b, v);
    if (r.ec != std::errc() || r.ptr != b || p == b) return false;
    *out = v; return true;                                // "0" reads as absent, as an empty field does
}

void scan_chunk(ObjChunk& c) {
    const char* p = c.begin;
    std::vector<std::pair<const char*, const char*>> w;
    while (p < c.end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(c.end - p));
        if (!eol) eol = c.end;
        const char *a = p, *b = eol;
        p = eol < c.end ? eol + 1 : c.end;
        while (a < b && is_ws(*a)) ++a;
        while (b > a && is_ws(b[-1])) --b;
        if (a == b || *a == '#') continue;
        w.clear();
        for (const char* q = a; q < b;) {
            while (q < b && is_ws(*q)) ++q;
            const char* t = q;
            while (q < b && !is_ws(*q)) ++q;
            if (q > t) w.emplace_back(t, q);
        }
        const size_t klen = (size_t)(w[0].second - w[0].first);
        const char* k = w[0].first;
        auto is = [&](const char* s) { return strlen(s) == klen && memcmp(k, s, klen) == 0; };
        auto floats = [&](size_t n, std::vector<float>& dst) {
            if (w.size() < 1 + n) return false;
            float v[4];
            for (size_t i = 0; i < n; ++i) if (!parse_f32_tok(w[1 + i].first, w[1 + i].second, &v[i])) return false;
            dst.insert(dst.end(), v, v + n); return true;
        };
        auto fail = [&](const std::string& m) { c.err = PTC_ERR_PARSE; c.msg = m; };
        if (is("v")) { if (!floats(3, c.pos)) { fail("OBJ: bad v"); return; } }
        else if (is("vt")) { if (!floats(2, c.tex)) { fail("OBJ: bad vt"); return; } }
        else if (is("vn")) { if (!floats(3, c.nrm)) { fail("OBJ: bad vn"); return; } }
        else if (is("f") || is("l")) {
            RawFace f; f.first = (uint32_t)c.elems.size(); f.count = 0;
            f.np = (uint32_t)(c.pos.size() / 3); f.nt = (uint32_t)(c.tex.size() / 2); f.nn = (uint32_t)(c.nrm.size() / 3);
            for (size_t i = 1; i < w.size(); ++i) {
                const char *t = w[i].first, *te = w[i].second;
                const char* part[3][2]; int np = 0;
                for (const char* q = t;;) {
                    const char* sl = (const char*)memchr(q, '/', (size_t)(te - q));
                    part[np][0] = q; part[np][1] = sl ? sl : te; ++np;
                    if (!sl || np == 3) break;
                    q = sl + 1;
                }
                RawElem e{0, 0, 0};
                bool ok = parse_index_tok(part[0][0], part[0][1], &e.v);
                if (np > 1) ok = ok && parse_index_tok(part[1][0], part[1][1], &e.vt);
                if (np > 2) ok = ok && parse_index_tok(part[2][0], part[2][1], &e.vn);
                if (!ok || e.v == 0) { fail("OBJ: bad face element '" + std::string(t, te) + "'"); return; }
                c.elems.push_back(e); f.count++;
            }
            if (f.count == 0) { fail("OBJ: empty face"); return; }
            c.faces.push_back(f);
        } else if (is("o") || is("g") || is("usemtl") || is("mtllib")) {
            ObjEvent ev; ev.kind = (is("o") || is("g")) ? 0 : is("usemtl") ? 1 : 2;
            ev.face_pos = (uint32_t)c.faces.size();
            const char* r = k + klen;                      // rest of the (trimmed) line after the keyword, trimmed
            while (r < b && is_ws(*r)) ++r;
            ev.arg.assign(r, b);
            c.events.push_back(std::move(ev));
        }
    }
}

struct FaceSpan { uint32_t chunk, first, last; };        // faces [first, last) of a chunk
struct ModelJob { std::string name; bool has_mat; size_t mat_id; std::vector<FaceSpan> spans; };

int export_model(const std::vector<ObjChunk>& chunks, const ObjGlobal& g, const ModelJob& job, TobjModel& m, std::string& msg) {
    m.name = job.name; m.has_material = job.has_mat; m.material_id = job.mat_id;
    size_t nelem = 0;
    for (const FaceSpan& sp : job.spans) for (uint32_t f = sp.first; f < sp.last; ++f) nelem += chunks[sp.chunk].faces[f].count;
    std::unordered_map<VertKey, uint32_t, VertKeyHash> index_map;
    index_map.reserve(nelem / 2 + 16);
    m.indices.reserve(nelem);
    int err = PTC_OK;
    auto add = [&](const VertKey& k) {
        auto it = index_map.find(k);
        if (it != index_map.end()) { m.indices.push_back(it->second); return; }
        if (k.v == kMissing || k.v * 3 + 2 >= g.pos.size()) { err = PTC_ERR_PARSE; return; }
        m.positions.insert(m.positions.end(), {g.pos[k.v * 3], g.pos[k.v * 3 + 1], g.pos[k.v * 3 + 2]});
        if (!g.tex.empty() && k.vt != kMissing) {
            if (k.vt * 2 + 1 >= g.tex.size()) { err = PTC_ERR_PARSE; return; }
            m.texcoords.insert(m.texcoords.end(), {g.tex[k.vt * 2], g.tex[k.vt * 2 + 1]});
        }
        if (!g.nrm.empty() && k.vn != kMissing) {
            if (k.vn * 3 + 2 >= g.nrm.size()) { err = PTC_ERR_PARSE; return; }
            m.normals.insert(m.normals.end(), {g.nrm[k.vn * 3], g.nrm[k.vn * 3 + 1], g.nrm[k.vn * 3 + 2]});
        }
        uint32_t next = (uint32_t)index_map.size();
        m.indices.push_back(next); index_map.emplace(k, next);
    };
    std::vector<VertKey> face;
    for (const FaceSpan& sp : job.spans) {
        const ObjChunk& c = chunks[sp.chunk];
        for (uint32_t fi = sp.first; fi < sp.last; ++fi) {
            const RawFace& rf = c.faces[fi];
            if (rf.count < 3) continue;                               // ignore_points, ignore_lines
            face.clear();
            auto resolve = [](long long raw, size_t count) -> size_t {  // 1-based, or relative to the elements seen so far
                if (raw == 0) return kMissing;
                return raw < 0 ? (size_t)((long long)count + raw) : (size_t)(raw - 1);
            };
            for (uint32_t e = 0; e < rf.count; ++e) {
                const RawElem& re = c.elems[rf.first + e];
                face.push_back(VertKey{resolve(re.v, g.pos_base[sp.chunk] + rf.np), resolve(re.vt, g.tex_base[sp.chunk] + rf.nt),
                                       resolve(re.vn, g.nrm_base[sp.chunk] + rf.nn)});
            }
            if (face.size() == 3) { add(face[0]); add(face[1]); add(face[2]); }
            else if (face.size() == 4) { add(face[0]); add(face[1]); add(face[2]); add(face[0]); add(face[2]); add(face[3]); }
            else { const VertKey a = face[0]; VertKey b = face[1]; for (size_t i = 2; i < face.size(); ++i) { add(a); add(b); add(face[i]); b = face[i]; } }
            if (err != PTC_OK) { msg = "OBJ: face references a vertex out of bounds"; return err; }
        }
    }
    return PTC_OK;
}

template <class F>
void run_parallel(size_t n, F f) {                        // f(i) for i in [0, n), work-stealing over the host threads
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<size_t>(std::min<size_t>(nt, 32), n);
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&] { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& x : th) x.join();
}

int parse_obj(const std::string& text, const std::string& mtl_override, bool use_override, const std::string& obj_dir,
              std::vector<TobjModel>& models, std::vector<TobjMaterial>& mats) {
    // pass 1: chunks cut at line ends, scanned in parallel
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    size_t chunk_bytes = 1u << 20;
    if (const char* e = getenv("PTC_OBJ_CHUNK_BYTES")) { long v = atol(e); if (v >= 1) chunk_bytes = (size_t)v; }   // tests cut small files into many chunks
    size_t nchunks = std::max<size_t>(1, std::min<size_t>(chunk_bytes < (1u << 20) ? 4096 : std::min<unsigned>(nt, 32) * 4, text.size() / chunk_bytes));
    std::vector<ObjChunk> chunks(nchunks);
    {
        const char *base = text.data(), *endp = text.data() + text.size();
        const char* cur = base;
        for (size_t c = 0; c < nchunks; ++c) {
            const char* want = c + 1 == nchunks ? endp : base + (text.size() / nchunks) * (c + 1);
            if (want < cur) want = cur;
            if (c + 1 < nchunks) { const char* nl = (const char*)memchr(want, '\n', (size_t)(endp - want)); want = nl ? nl + 1 : endp; }
            chunks[c].begin = cur; chunks[c].end = want; cur = want;
        }
    }
    run_parallel(nchunks, [&](size_t c) { scan_chunk(chunks[c]); });
    for (const ObjChunk& c : chunks) if (c.err != PTC_OK) { set_error(c.msg); return c.err; }
    ObjGlobal g;
    g.pos_base.resize(nchunks); g.tex_base.resize(nchunks); g.nrm_base.resize(nchunks);
    size_t np = 0, ntx = 0, nn = 0;
    for (size_t c = 0; c < nchunks; ++c) {
        g.pos_base[c] = np / 3; g.tex_base[c] = ntx / 2; g.nrm_base[c] = nn / 3;
        np += chunks[c].pos.size(); ntx += chunks[c].tex.size(); nn += chunks[c].nrm.size();
    }
    g.pos.resize(np); g.tex.resize(ntx); g.nrm.resize(nn);
    run_parallel(nchunks, [&](size_t c) {
        if (!chunks[c].pos.empty()) std::memcpy(&g.pos[g.pos_base[c] * 3], chunks[c].pos.data(), chunks[c].pos.size() * 4);
        if (!chunks[c].tex.empty()) std::memcpy(&g.tex[g.tex_base[c] * 2], chunks[c].tex.data(), chunks[c].tex.size() * 4);
        if (!chunks[c].nrm.empty()) std::memcpy(&g.nrm[g.nrm_base[c] * 3], chunks[c].nrm.data(), chunks[c].nrm.size() * 4);
    });

    // pass 2: events -> models (tobj: a new model on o / g with pending faces and on a usemtl that changes the material with
    // pending faces; the last model is pushed unconditionally)
    std::unordered_map<std::string, size_t> mat_map;
    std::vector<ModelJob> jobs;
    ModelJob cur; cur.name = "unnamed_object"; cur.has_mat = false; cur.mat_id = 0;
    size_t pending = 0;
    auto flush = [&](bool always) {
        if (!always && pending == 0) return;
        jobs.push_back(cur);
        cur.spans.clear(); pending = 0;
    };
    for (size_t c = 0; c < nchunks; ++c) {
        uint32_t at = 0;
        auto take = [&](uint32_t upto) { if (upto > at) { cur.spans.push_back(FaceSpan{(uint32_t)c, at, upto}); pending += upto - at; at = upto; } };
        for (const ObjEvent& ev : chunks[c].events) {
            take(ev.face_pos);
            if (ev.kind == 0) {
                flush(false);
                cur.name = ev.arg.empty() ? std::string("unnamed_object") : ev.arg;
            } else if (ev.kind == 2) {
                std::string mtl_text;
                if (use_override) mtl_text = mtl_override;
                else {
                    std::string p = join_path(obj_dir, ev.arg);
                    std::ifstream f(p, std::ios::binary);
                    if (!f) { set_error("Failed to load MTL " + p); return PTC_ERR_IO; }
                    std::stringstream ss; ss << f.rdbuf(); mtl_text = ss.str();
                }
                std::vector<TobjMaterial> loaded; std::unordered_map<std::string, size_t> lm;
                int rc = parse_mtl(mtl_text, loaded, lm); if (rc) return rc;
                size_t base = mats.size();
                for (auto& kv : lm) mat_map[kv.first] = kv.second + base;
                mats.insert(mats.end(), loaded.begin(), loaded.end());
            } else {
                if (ev.arg.empty()) { set_error("OBJ: usemtl without a name"); return PTC_ERR_PARSE; }
                auto it = mat_map.find(ev.arg);
                bool new_has = it != mat_map.end(); size_t new_id = new_has ? it->second : 0;
                bool differs = (new_has != cur.has_mat) || (new_has && new_id != cur.mat_id);
                if (differs) flush(false);
                cur.has_mat = new_has; cur.mat_id = new_id;
            }
        }
        take((uint32_t)chunks[c].faces.size());
    }
    flush(true);

    // pass 3: models exported in parallel, in order
    const size_t m0 = models.size();
    models.resize(m0 + jobs.size());
    std::vector<int> errs(jobs.size(), PTC_OK); std::vector<std::string> msgs(jobs.size());
    run_parallel(jobs.size(), [&](size_t j) { errs[j] = export_model(chunks, g, jobs[j], models[m0 + j], msgs[j]); });
    for (size_t j = 0; j < jobs.size(); ++j) if (errs[j] != PTC_OK) { set_error(msgs[j]); return errs[j]; }
    return PTC_OK;
}

int ingest_unchecked(ptc_scene* s, const std::vector<TobjModel>& models, const std::vector<TobjMaterial>& mats, const std::string& dir);
// The reference panics on a bad OBJ/MTL, so no half-loaded Scene ever exists there; here a failing load rolls the scene
// back to the sizes it had before the call (the arrays only ever grow inside one load), so the caller keeps a usable scene.
int ingest(ptc_scene* s, const std::vector<TobjModel>& models, const std::vector<TobjMaterial>& mats, const std::string& dir) {
    const size_t n_mat = s->materials.size(), n_tex = s->textures.size(), n_vtx = s->vertices.size(), n_idx = s->indices.size(), n_mesh = s->meshes.size();
    const int rc = ingest_unchecked(s, models, mats, dir);
    if (rc != PTC_OK) {
        s->materials.resize(n_mat); s->textures.resize(n_tex); s->vertices.resize(n_vtx); s->indices.resize(n_idx); s->meshes.resize(n_mesh);
    }
    return rc;
}
int ingest_unchecked(ptc_scene* s, const std::vector<TobjModel>& models, const std::vector<TobjMaterial>& mats, const std::string& dir) {
    const size_t prev_num_materials = s->materials.size();
    for (const TobjMaterial& m : mats) {                                                  // scene.rs:134-167
        ptc_material out; std::memset(&out, 0, sizeof(out));
        uint32_t idx; int r;
        out.albedo_texture_idx = known_vector(s, m.has_diffuse_texture, m.diffuse_texture, m.has_diffuse, m.diffuse, dir, &idx) ? idx : 0;
        r = texture_prop(s, m, "map_Pr", "Pr", true, dir, &idx); if (r < 0) return r; out.roughness_texture_idx = r ? idx : 1;
        out.specular_texture_idx = known_vector(s, m.has_specular_texture, m.specular_texture, m.has_specular, m.specular, dir, &idx) ? idx : 2;
        r = texture_prop(s, m, "map_Pm", "Pm", true, dir, &idx); if (r < 0) return r; out.metallic_texture_idx = r ? idx : 3;
        r = texture_prop(s, m, "map_Ke", "Ke", false, dir, &idx); if (r < 0) return r; out.emissive_texture_idx = r ? idx : 4;
        r = texture_prop(s, m, "map_Ps", "Ps", true, dir, &idx); if (r < 0) return r; out.sheen_texture_idx = r ? idx : 5;
        out.normal_texture_idx = known_vector(s, m.has_normal_texture, m.normal_texture, false, nullptr, dir, &idx) ? idx : 6;
        r = texture_prop(s, m, "map_Tf", "Tf", false, dir, &idx); if (r < 0) return r; out.transmission_texture_idx = r ? idx : 7;
        out.ior_texture_idx = known_scalar(s, m.has_ni, m.optical_density, &idx) ? idx : 8;
        float f;
        out.clearcoat_thickness = scalar_prop(m, "Pc", &f) ? f : 0.0f;
        out.clearcoat_roughness = scalar_prop(m, "Pcr", &f) ? f : 0.0f;
        out.anisotropy = scalar_prop(m, "Pn", &f) ? f : 0.0f;
        out.anisotropy_rotation = scalar_prop(m, "Pnr", &f) ? f : 0.0f;
        s->materials.push_back(out);
    }
    for (const TobjModel& mesh : models) {                                                // scene.rs:169-234
        const size_t prev_num_vertices = s->vertices.size();
        const size_t nv = mesh.positions.size() / 3;
        if (!mesh.normals.empty() && mesh.normals.size() < nv * 3) { set_error("OBJ: '" + mesh.name + "' mixes vertices with and without normals (the reference panics)"); return PTC_ERR_PARSE; }
        if (!mesh.texcoords.empty() && mesh.texcoords.size() < nv * 2) { set_error("OBJ: '" + mesh.name + "' mixes vertices with and without texcoords (the reference panics)"); return PTC_ERR_PARSE; }
        for (size_t i = 0; i < nv; ++i) {
            ptc_vertex v; std::memset(&v, 0, sizeof(v));
            std::memcpy(v.position, &mesh.positions[i * 3], 12);
            if (!mesh.normals.empty()) std::memcpy(v.normal, &mesh.normals[i * 3], 12);
            if (!mesh.texcoords.empty()) { v.uv[0] = mesh.texcoords[i * 2]; v.uv[1] = 1.0f - mesh.texcoords[i * 2 + 1]; }
            s->vertices.push_back(v);
        }
        const size_t prev_num_indices = s->indices.size();
        for (uint32_t ix : mesh.indices) s->indices.push_back(ix + (uint32_t)prev_num_vertices);
        if (mesh.normals.empty()) {
            // scene.rs:202-215.  The reference indexes the GLOBAL vertex array with MESH-LOCAL indices (no
            // `+ prev_num_vertices`), so for every mesh after the first the face normals are computed from, and
            // written to, vertices of earlier meshes.  Reproduced by default; fix_normals applies the offset.
            const size_t off = s->fix_normals ? prev_num_vertices : 0;
            for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
                size_t i0 = mesh.indices[t] + off, i1 = mesh.indices[t + 1] + off, i2 = mesh.indices[t + 2] + off;
                V3 v0 = ld(s->vertices[i0].position), v1 = ld(s->vertices[i1].position), v2 = ld(s->vertices[i2].position);
                V3 n = normv(crossv(sub(v1, v0), sub(v2, v0)));
                st(s->vertices[i0].normal, n); st(s->vertices[i1].normal, n); st(s->vertices[i2].normal, n);
            }
        }
        ptc_mesh out;
        out.vertex_offset = (uint32_t)prev_num_vertices; out.vertex_count = (uint32_t)nv;
        out.index_offset = (uint32_t)prev_num_indices; out.index_count = (uint32_t)mesh.indices.size();
        out.material_index = (uint32_t)((mesh.has_material ? mesh.material_id : 0) + prev_num_materials);   // scene.rs:224
        if (out.material_index >= s->materials.size()) { set_error("mesh '" + mesh.name + "' has no material and the OBJ loaded none (the reference indexes out of bounds)"); return PTC_ERR_PARSE; }
        s->meshes.push_back(out);
    }
    return PTC_OK;
}

}  // namespace

extern "C" {

PTC_API int ptc_scene_new(const char* envmap_path, ptc_scene** out) {                     // scene.rs:77-122
    return ptc::guarded("ptc_scene_new", [&]() -> int {
    if (!out) { set_error("null argument"); return PTC_ERR_INVALID; }
    std::unique_ptr<ptc_scene> owner(new ptc_scene());
    ptc_scene* s = owner.get();
    push_color(s, 1.0f, 1.0f, 1.0f, 1.0f);      // 0 albedo default
    push_scalar(s, 0.99f);                      // 1 roughness default
    push_color(s, 0.5f, 0.5f, 0.5f, 1.0f);      // 2 specular default
    push_scalar(s, 0.0f);                       // 3 metallic default
    push_color(s, 0.0f, 0.0f, 0.0f, 1.0f);      // 4 emissive default (index hard-coded elsewhere, pt.rs:217)
    push_scalar(s, 0.0f);                       // 5 sheen default
    push_color(s, 0.5f, 0.5f, 1.0f, 1.0f);      // 6 normal default
    push_scalar(s, 0.0f);                       // 7 transmission weight default
    push_scalar(s, 1.0f);                       // 8 IOR default
    push_file(s, envmap_path ? envmap_path : "res/knob/envmap.exr");   // 9 envmap (sky.wgsl:10)
    ptc_material m; std::memset(&m, 0, sizeof(m));
    m.albedo_texture_idx = 0; m.roughness_texture_idx = 1; m.specular_texture_idx = 2; m.metallic_texture_idx = 3;
    m.emissive_texture_idx = 4; m.sheen_texture_idx = 5; m.normal_texture_idx = 6; m.transmission_texture_idx = 7; m.ior_texture_idx = 8;
    s->materials.push_back(m);
    *out = owner.release();
    return PTC_OK;
    });
}

PTC_API int ptc_scene_set_fix_normals(ptc_scene* s, int enable) {
    if (!s) { set_error("null scene"); return PTC_ERR_INVALID; }
    s->fix_normals = enable != 0; return PTC_OK;
}

PTC_API int ptc_scene_load_obj_text(ptc_scene* s, const char* obj_text, const char* mtl_text, const char* obj_dir) {
    return ptc::guarded("ptc_scene_load_obj_text", [&]() -> int {
    if (!s || !obj_text) { set_error("null argument"); return PTC_ERR_INVALID; }
    std::vector<TobjModel> models; std::vector<TobjMaterial> mats;
    std::string dir = obj_dir ? obj_dir : "";
    int rc = parse_obj(obj_text, mtl_text ? mtl_text : "", mtl_text != nullptr, dir, models, mats);
    if (rc) return rc;
    return ingest(s, models, mats, dir);
    });
}

PTC_API int ptc_scene_load_obj(ptc_scene* s, const char* obj_path) {                      // scene.rs:124-128
    return ptc::guarded("ptc_scene_load_obj", [&]() -> int {
    if (!s || !obj_path) { set_error("null argument"); return PTC_ERR_INVALID; }
    std::ifstream f(obj_path, std::ios::binary);
    if (!f) { set_error(std::string("Failed to load OBJ ") + obj_path); return PTC_ERR_IO; }
    std::stringstream ss; ss << f.rdbuf();
    std::string text = ss.str();
    if (text.compare(0, 40, "version https://git-lfs.github.com/spec/") == 0) { set_error(std::string(obj_path) + " is a Git-LFS pointer, not geometry"); return PTC_ERR_IO; }
    std::vector<TobjModel> models; std::vector<TobjMaterial> mats;
    std::string dir = parent_dir(obj_path);
    int rc = parse_obj(text, "", false, dir, models, mats);
    if (rc) return rc;
    return ingest(s, models, mats, dir);
    });
}

PTC_API uint64_t ptc_scene_num_textures(const ptc_scene* s) { return s ? s->textures.size() : 0; }

PTC_API int ptc_scene_texture_info_at(const ptc_scene* s, uint64_t i, ptc_scene_texture_info* out) {
    if (!s || !out || i >= s->textures.size()) { set_error("bad texture index"); return PTC_ERR_INVALID; }
    const SceneTexture& t = s->textures[i];
    std::memset(out, 0, sizeof(*out));
    out->source = t.source; out->width = t.width; out->height = t.height; out->format = t.format;
    std::memcpy(out->color, t.color, 16);
    out->path = t.source == PTC_TEX_FILE ? t.path.c_str() : nullptr;
    return PTC_OK;
}

PTC_API int ptc_scene_set_texels(ptc_scene* s, uint64_t i, uint32_t w, uint32_t h, uint32_t format, const void* texels) {
    return ptc::guarded("ptc_scene_set_texels", [&]() -> int {
    if (!s || !texels || i >= s->textures.size() || w == 0 || h == 0 || format > PTC_TEXEL_RGBA32F) { set_error("bad texel upload"); return PTC_ERR_INVALID; }
    if ((uint64_t)w * h > (1ull << 30)) { set_error("texture larger than 2^30 texels"); return PTC_ERR_INVALID; }      // the decoders' bound; keeps the byte count inside 64 bits
    SceneTexture& t = s->textures[i];
    size_t bytes = (size_t)w * h * (format == PTC_TEXEL_RGBA8 ? 4 : 16);
    t.texels.assign((const uint8_t*)texels, (const uint8_t*)texels + bytes);
    t.width = w; t.height = h; t.format = format;
    return PTC_OK;
    });
}

// Texture::load for every file-backed texture that has no texels yet (texture.rs:232-243), on n_threads host threads as
// Scene::load_obj does (scene.rs:239-254, chunks of 32 threads).  PNG and PFM are decoded on the host, JPEG by nvJPEG on
// `cuda_device` (each worker owns its decoder: handle, state, stream, device buffer).
static int read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return 1;
    std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return 1; }
    out.resize((size_t)n);
    const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n ? 0 : 1;
}
static int decode_bytes(const uint8_t* data, size_t n, int cuda_device, std::unique_ptr<ptc::JpegDecoder>& jpeg, ptc::DecodedImage& img, std::string& err) {
    switch (ptc::sniff_image(data, n)) {
        case ptc::IMG_PNG: return ptc::decode_png(data, n, img, err);
        case ptc::IMG_PFM: return ptc::decode_pfm(data, n, img, err);
        case ptc::IMG_EXR: return ptc::decode_exr(data, n, img, err);
        case ptc::IMG_HDR: return ptc::decode_hdr(data, n, img, err);
        case ptc::IMG_JPEG:
            if (!jpeg) jpeg.reset(new ptc::JpegDecoder(cuda_device));
            return jpeg->decode(data, n, img, err);
        default: err = "unknown image format (this library decodes PNG, JPEG, OpenEXR, Radiance HDR and PFM)"; return 1;
    }
}
PTC_API int ptc_scene_decode_textures(ptc_scene* s, int cuda_device, int n_threads) {
    return ptc::guarded("ptc_scene_decode_textures", [&]() -> int {
    if (!s) { set_error("null scene"); return PTC_ERR_INVALID; }
    std::vector<size_t> todo;
    for (size_t i = 0; i < s->textures.size(); ++i) if (s->textures[i].source == PTC_TEX_FILE && s->textures[i].texels.empty()) todo.push_back(i);
    if (todo.empty()) return PTC_OK;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    n_threads = std::max(1, std::min(n_threads, (int)todo.size()));
    std::atomic<size_t> next{0};
    std::mutex mu; std::string first_error;
    auto worker = [&] {
        std::unique_ptr<ptc::JpegDecoder> jpeg;
        std::vector<uint8_t> bytes;
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= todo.size()) break;
            SceneTexture& t = s->textures[todo[k]];
            ptc::DecodedImage img; std::string err;
            int rc = 1;
            try {                                                 // a worker thread must not let an exception escape (std::terminate)
                rc = read_file(t.path, bytes);
                if (rc) err = "cannot read the file";
                else rc = decode_bytes(bytes.data(), bytes.size(), cuda_device, jpeg, img, err);
            } catch (const std::bad_alloc&) { rc = 1; err = "out of host memory"; }
            catch (const std::exception& e) { rc = 1; err = e.what(); }
            if (rc) { std::lock_guard<std::mutex> l(mu); if (first_error.empty()) first_error = "texture " + t.path + ": " + err; continue; }
            t.texels.swap(img.texels); t.width = img.width; t.height = img.height; t.format = img.format;
        }
    };
    std::vector<std::thread> th;
    try { for (int i = 1; i < n_threads; ++i) th.emplace_back(worker); }
    catch (const std::exception&) {}                              // fewer threads than asked for: the started ones and this one finish the list
    worker();
    for (auto& x : th) x.join();
    if (!first_error.empty()) { set_error(first_error); return PTC_ERR_INVALID; }
    return PTC_OK;
    });
}

PTC_API int ptc_decode_image(const void* data, size_t bytes, int cuda_device, uint32_t* width, uint32_t* height, uint32_t* format, void** texels) {
    return ptc::guarded("ptc_decode_image", [&]() -> int {
    if (!data || !width || !height || !format || !texels) { set_error("null argument"); return PTC_ERR_INVALID; }
    std::unique_ptr<ptc::JpegDecoder> jpeg; ptc::DecodedImage img; std::string err;
    if (decode_bytes((const uint8_t*)data, bytes, cuda_device, jpeg, img, err)) { set_error(err); return PTC_ERR_INVALID; }
    void* p = std::malloc(img.texels.size());
    if (!p) { set_error("out of memory"); return PTC_ERR_OOM; }
    std::memcpy(p, img.texels.data(), img.texels.size());
    *width = img.width; *height = img.height; *format = img.format; *texels = p;
    return PTC_OK;
    });
}
PTC_API void ptc_free(void* p) { std::free(p); }

PTC_API int ptc_scene_get_view(const ptc_scene* cs, ptc_scene_view* out) {
    return ptc::guarded("ptc_scene_get_view", [&]() -> int {
    ptc_scene* s = const_cast<ptc_scene*>(cs);
    if (!s || !out) { set_error("null argument"); return PTC_ERR_INVALID; }
    s->descs.resize(s->textures.size());
    for (size_t i = 0; i < s->textures.size(); ++i) {
        const SceneTexture& t = s->textures[i];
        if (t.texels.empty()) { set_error("texture " + std::to_string(i) + " (" + t.path + ") has not been decoded; call ptc_scene_set_texels"); return PTC_ERR_INVALID; }
        s->descs[i].width = t.width; s->descs[i].height = t.height; s->descs[i].format = t.format; s->descs[i]._pad = 0;
        s->descs[i].texels = t.texels.data();
    }
    std::memset(out, 0, sizeof(*out));
    out->vertices = s->vertices.data(); out->num_vertices = s->vertices.size();
    out->indices = s->indices.data(); out->num_indices = s->indices.size();
    out->meshes = s->meshes.data(); out->num_meshes = s->meshes.size();
    out->materials = s->materials.data(); out->num_materials = s->materials.size();
    out->textures = s->descs.data(); out->num_textures = s->descs.size();
    return PTC_OK;
    });
}

PTC_API void ptc_scene_free(ptc_scene* s) { delete s; }

}  // extern "C"
