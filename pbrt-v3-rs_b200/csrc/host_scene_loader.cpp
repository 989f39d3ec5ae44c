// Scene ingestion: a reader for the subset of the pbrt-v3 scene format that reaches this
// path, the PLY reader behind Shape "plymesh", and PFM image IO.  It does on the host what the reference's `api`
// crate does above the accelerator / integrator boundary and hands the result over as a b200pt_scene_desc:
//   directives / graphics state   api/src/lib.rs:240-1000 (pbrt_* calls), api/src/graphics_state.rs
//   Transform algebra             core/src/geometry/transform.rs:60-260, 640-660, matrix4x4.rs:55-123, 181-200
//   trianglemesh / plymesh        shapes/src/triangle.rs:60-130, 184-330, shapes/src/plymesh.rs:25-255
//   camera / film / filter        cameras/src/perspective_camera.rs:35-75, 357-421, core/src/camera.rs:276-306,
//                                 core/src/film/mod.rs:89-146, 420-485, filters/src/{boxf,gaussian}.rs
//   lights / materials params     lights/src/{point,diffuse,infinite}.rs (From impls), materials/src/*.rs (From impls)
// Everything is f32 with the reference's operation order (host code is built with -ffp-contract=off).
// Outside this path (reported as B200PT_ERR_UNSUPPORTED, never silently dropped): textures, media, other shapes,
// cameras, samplers, integrators, lights and materials, animated transforms, spectrum files.
#include <cctype>
#include <cmath>
#include <iterator>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <zlib.h>

#include "../../include/b200pt.h"

extern "C" int b200pt_set_error(const char* msg);

namespace b2load {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
struct Invalid : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------------------------------------------
// Matrix4x4 / Transform
struct M4 {
    float m[4][4];
};
static M4 m4_identity() {
    M4 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0f : 0.0f;
    return r;
}
static M4 m4_mul(const M4& a, const M4& b) {  // matrix4x4.rs:181-200
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
static M4 m4_transpose(const M4& a) {
    M4 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
static M4 m4_inverse(const M4& src) {  // Gauss-Jordan with full pivoting, matrix4x4.rs:55-123
    int indxc[4], indxr[4], ipiv[4] = {0, 0, 0, 0};
    M4 minv = src;
    for (int i = 0; i < 4; ++i) {
        int irow = 0, icol = 0;
        float big = 0.0f;
        for (int j = 0; j < 4; ++j) {
            if (ipiv[j] != 1) {
                for (int k = 0; k < 4; ++k) {
                    if (ipiv[k] == 0) {
                        if (std::fabs(minv.m[j][k]) >= big) { big = std::fabs(minv.m[j][k]); irow = j; icol = k; }
                    } else if (ipiv[k] > 1) throw Invalid("singular matrix in a transform");
                }
            }
        }
        ipiv[icol] += 1;
        if (irow != icol) for (int k = 0; k < 4; ++k) std::swap(minv.m[irow][k], minv.m[icol][k]);
        indxr[i] = irow; indxc[i] = icol;
        if (minv.m[icol][icol] == 0.0f) throw Invalid("singular matrix in a transform");
        float pivinv = 1.0f / minv.m[icol][icol];
        minv.m[icol][icol] = 1.0f;
        for (int j = 0; j < 4; ++j) minv.m[icol][j] *= pivinv;
        for (int j = 0; j < 4; ++j) {
            if (j != icol) {
                float save = minv.m[j][icol];
                minv.m[j][icol] = 0.0f;
                for (int k = 0; k < 4; ++k) minv.m[j][k] -= minv.m[icol][k] * save;
            }
        }
    }
    for (int j = 3; j >= 0; --j)
        if (indxr[j] != indxc[j]) for (int k = 0; k < 4; ++k) std::swap(minv.m[k][indxr[j]], minv.m[k][indxc[j]]);
    return minv;
}

struct Xf {  // Transform {m, m_inv}
    M4 m, inv;
};
static Xf xf_identity() { return Xf{m4_identity(), m4_identity()}; }
static Xf xf_mul(const Xf& a, const Xf& b) { return Xf{m4_mul(a.m, b.m), m4_mul(b.inv, a.inv)}; }  // transform.rs:640-660
static Xf xf_inverse(const Xf& a) { return Xf{a.inv, a.m}; }
static Xf xf_from(const M4& m) { return Xf{m, m4_inverse(m)}; }
static float radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }  // f32::to_radians
static Xf xf_translate(float x, float y, float z) {
    Xf t = xf_identity();
    t.m.m[0][3] = x; t.m.m[1][3] = y; t.m.m[2][3] = z;
    t.inv.m[0][3] = -x; t.inv.m[1][3] = -y; t.inv.m[2][3] = -z;
    return t;
}
static Xf xf_scale(float x, float y, float z) {
    Xf t = xf_identity();
    t.m.m[0][0] = x; t.m.m[1][1] = y; t.m.m[2][2] = z;
    t.inv.m[0][0] = 1.0f / x; t.inv.m[1][1] = 1.0f / y; t.inv.m[2][2] = 1.0f / z;
    return t;
}
struct V3 { float x, y, z; };
static V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static V3 cross(V3 a, V3 b) { return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
static float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static V3 normalize(V3 a) { float inv = 1.0f / std::sqrt(len2(a)); return v3(inv * a.x, inv * a.y, inv * a.z); }  // v / length: multiply by 1/len
static Xf xf_rotate(float theta, V3 axis) {  // transform.rs rotate_axis
    V3 a = normalize(axis);
    float r = radians(theta), s = std::sin(r), c = std::cos(r);
    M4 m = m4_identity();
    m.m[0][0] = a.x * a.x + (1.0f - a.x * a.x) * c;
    m.m[0][1] = a.x * a.y * (1.0f - c) - a.z * s;
    m.m[0][2] = a.x * a.z * (1.0f - c) + a.y * s;
    m.m[1][0] = a.x * a.y * (1.0f - c) + a.z * s;
    m.m[1][1] = a.y * a.y + (1.0f - a.y * a.y) * c;
    m.m[1][2] = a.y * a.z * (1.0f - c) - a.x * s;
    m.m[2][0] = a.x * a.z * (1.0f - c) - a.y * s;
    m.m[2][1] = a.y * a.z * (1.0f - c) + a.x * s;
    m.m[2][2] = a.z * a.z + (1.0f - a.z * a.z) * c;
    return Xf{m, m4_transpose(m)};
}
static Xf xf_look_at(V3 pos, V3 look, V3 up) {  // transform.rs look_at: m = inverse(camera_to_world)
    V3 dir = normalize(sub(look, pos));
    V3 right = cross(normalize(up), dir);
    if (std::sqrt(len2(right)) == 0.0f) throw Invalid("LookAt: up vector and viewing direction point the same way");
    right = normalize(right);
    V3 nu = cross(dir, right);
    M4 c2w = m4_identity();
    c2w.m[0][0] = right.x; c2w.m[0][1] = nu.x; c2w.m[0][2] = dir.x; c2w.m[0][3] = pos.x;
    c2w.m[1][0] = right.y; c2w.m[1][1] = nu.y; c2w.m[1][2] = dir.y; c2w.m[1][3] = pos.y;
    c2w.m[2][0] = right.z; c2w.m[2][1] = nu.z; c2w.m[2][2] = dir.z; c2w.m[2][3] = pos.z;
    return Xf{m4_inverse(c2w), c2w};
}
static Xf xf_perspective(float fov, float n, float f) {
    M4 p = m4_identity();
    p.m[2][2] = f / (f - n);
    p.m[2][3] = -f * n / (f - n);
    p.m[3][2] = 1.0f;
    p.m[3][3] = 0.0f;
    float inv_tan = 1.0f / std::tan(radians(fov) / 2.0f);
    return xf_mul(xf_scale(inv_tan, inv_tan, 1.0f), xf_from(p));
}
static V3 xf_point(const M4& m, V3 p) {  // transform_point, transform.rs:288-302
    float xp = m.m[0][0] * p.x + m.m[0][1] * p.y + m.m[0][2] * p.z + m.m[0][3];
    float yp = m.m[1][0] * p.x + m.m[1][1] * p.y + m.m[1][2] * p.z + m.m[1][3];
    float zp = m.m[2][0] * p.x + m.m[2][1] * p.y + m.m[2][2] * p.z + m.m[2][3];
    float wp = m.m[3][0] * p.x + m.m[3][1] * p.y + m.m[3][2] * p.z + m.m[3][3];
    if (wp == 1.0f) return v3(xp, yp, zp);
    float inv = 1.0f / wp;
    return v3(inv * xp, inv * yp, inv * zp);
}
static V3 xf_vector(const M4& m, V3 v) {
    return v3(m.m[0][0] * v.x + m.m[0][1] * v.y + m.m[0][2] * v.z, m.m[1][0] * v.x + m.m[1][1] * v.y + m.m[1][2] * v.z,
              m.m[2][0] * v.x + m.m[2][1] * v.y + m.m[2][2] * v.z);
}
static V3 xf_normal(const M4& inv, V3 n) {  // transform_normal: inverse transpose
    return v3(inv.m[0][0] * n.x + inv.m[1][0] * n.y + inv.m[2][0] * n.z, inv.m[0][1] * n.x + inv.m[1][1] * n.y + inv.m[2][1] * n.z,
              inv.m[0][2] * n.x + inv.m[1][2] * n.y + inv.m[2][2] * n.z);
}
static bool swaps_handedness(const M4& m) {  // transform.rs:593-599
    float det = m.m[0][0] * (m.m[1][1] * m.m[2][2] - m.m[1][2] * m.m[2][1]) - m.m[0][1] * (m.m[1][0] * m.m[2][2] - m.m[1][2] * m.m[2][0]) +
                m.m[0][2] * (m.m[1][0] * m.m[2][1] - m.m[1][1] * m.m[2][0]);
    return det < 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// Tokens and parameter lists
struct Token {
    enum Kind { Ident, String, Number, LBracket, RBracket, End } kind;
    std::string text;
    double num = 0.0;
};
struct Lexer {
    std::string src, file;
    size_t pos = 0;
    int line = 1;
    Token peeked;
    bool has_peek = false;
    Token next() {
        if (has_peek) { has_peek = false; return peeked; }
        for (;;) {
            while (pos < src.size() && std::isspace((unsigned char)src[pos])) { if (src[pos] == '\n') ++line; ++pos; }
            if (pos < src.size() && src[pos] == '#') { while (pos < src.size() && src[pos] != '\n') ++pos; continue; }
            break;
        }
        Token t;
        if (pos >= src.size()) { t.kind = Token::End; return t; }
        char c = src[pos];
        if (c == '[') { ++pos; t.kind = Token::LBracket; return t; }
        if (c == ']') { ++pos; t.kind = Token::RBracket; return t; }
        if (c == '"') {
            size_t e = src.find('"', pos + 1);
            if (e == std::string::npos) fail("unterminated string");
            t.kind = Token::String; t.text = src.substr(pos + 1, e - pos - 1); pos = e + 1;
            return t;
        }
        if (std::isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.') {
            char* endp = nullptr;
            t.num = std::strtod(src.c_str() + pos, &endp);
            if (endp == src.c_str() + pos) fail("bad number");
            t.kind = Token::Number; t.text = src.substr(pos, (size_t)(endp - (src.c_str() + pos)));
            pos = (size_t)(endp - src.c_str());
            return t;
        }
        size_t b = pos;
        while (pos < src.size() && (std::isalnum((unsigned char)src[pos]) || src[pos] == '_')) ++pos;
        if (pos == b) fail(std::string("unexpected character '") + c + "'");
        t.kind = Token::Ident; t.text = src.substr(b, pos - b);
        return t;
    }
    const Token& peek() { if (!has_peek) { peeked = next(); has_peek = true; } return peeked; }
    [[noreturn]] void fail(const std::string& m) const { throw Invalid(file + ":" + std::to_string(line) + ": " + m); }
};

struct Param {
    std::string type, name;
    std::vector<float> nums;          // f32 like the reference's parser (strtod then `as f32`)
    std::vector<int32_t> ints;        // "integer" parameters, parsed as i32 like the reference (api/src/parser/mod.rs:638): indices above 2^24 survive
    std::vector<std::string> strs;    // string / bool / texture / spectrum-file values
    mutable bool used = false;
};
// <dir of libb200pt.so>/data or $B200PT_DATA_DIR: the published constant tables the reference links in (Sobol matrices, Perlin's
// permutation, CIE colour-matching functions).
static std::string lib_data_dir() {
    if (const char* e = std::getenv("B200PT_DATA_DIR")) return e;
    Dl_info info;
    if (dladdr((void*)&b200pt_load_pbrt, &info) && info.dli_fname) {
        std::string lib(info.dli_fname);
        size_t sl = lib.find_last_of('/');
        return (sl == std::string::npos ? std::string(".") : lib.substr(0, sl)) + "/data";
    }
    return "data";
}
// "blackbody L" [T scale] -> RGB exactly as the reference forms it: ParamSet::add_blackbody_spectrum (paramset/mod.rs:236-249)
// samples blackbody_normalized (spectrum/common.rs:361-398, f32 arithmetic, expf) at the 471 CIE wavelengths, RGBSpectrum::from
// (spectrum/rgb_spectrum.rs:76-103) sums value x CIE_X / Y / Z in order (the samples sit exactly on the CIE wavelengths, so
// interpolate_spectrum_samples returns them unchanged), scales by (830 - 360) / (CIE_Y_INTEGRAL * 471) and converts XYZ -> RGB;
// then `scale * spectrum`.
static float planck(float lambda_nm, float t) {
    const float C = 299792458.0f, H = 6.62606957e-34f, KB = 1.3806488e-23f;
    const float l = lambda_nm * 1e-9f;
    const float lambda5 = (l * l) * (l * l) * l;
    return (2.0f * H * C * C) / (lambda5 * (std::exp((H * C) / (l * KB * t)) - 1.0f));
}
static void blackbody_rgb(float t, float scale, float out[3]) {
    static std::vector<float> cie;
    if (cie.empty()) {
        std::ifstream f(lib_data_dir() + "/cie_xyz.bin", std::ios::binary);
        cie.resize(3 * 471);
        if (!f || !f.read((char*)cie.data(), (std::streamsize)(cie.size() * 4))) { cie.clear(); throw Invalid("\"blackbody\" parameter: cannot read " + lib_data_dir() + "/cie_xyz.bin (set B200PT_DATA_DIR)"); }
    }
    float xyz[3] = {0.0f, 0.0f, 0.0f};
    if (t > 0.0f) {
        const float lambda_max = 2.8977721e-3f / t * 1e9f;
        const float max_l = planck(lambda_max, t);
        for (int i = 0; i < 471; ++i) {
            const float val = planck((float)(360 + i), t) / max_l;
            xyz[0] += val * cie[(size_t)i]; xyz[1] += val * cie[471 + (size_t)i]; xyz[2] += val * cie[942 + (size_t)i];
        }
    }
    const float sc = (float)(830 - 360) / (106.856895f * (float)471);
    for (int c = 0; c < 3; ++c) xyz[c] *= sc;
    const float rgbv[3] = {3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2], -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2],
                           0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2]};
    for (int c = 0; c < 3; ++c) out[c] = scale * rgbv[c];
}

struct ParamSet {
    std::vector<Param> ps;
    const Param* find(const std::string& name, const char* t1, const char* t2 = nullptr, const char* t3 = nullptr) const {
        for (const Param& p : ps)
            if (p.name == name && (p.type == t1 || (t2 && p.type == t2) || (t3 && p.type == t3))) { p.used = true; return &p; }
        return nullptr;
    }
    float one_float(const std::string& n, float d) const { const Param* p = find(n, "float"); return p && !p->nums.empty() ? p->nums[0] : d; }
    int one_int(const std::string& n, int d) const { const Param* p = find(n, "integer"); return p && !p->ints.empty() ? (int)p->ints[0] : d; }
    bool one_bool(const std::string& n, bool d) const { const Param* p = find(n, "bool"); return p && !p->strs.empty() ? p->strs[0] == "true" : d; }
    std::string one_string(const std::string& n, const std::string& d) const { const Param* p = find(n, "string"); return p && !p->strs.empty() ? p->strs[0] : d; }
    std::vector<float> floats(const std::string& n) const { const Param* p = find(n, "float"); return p ? p->nums : std::vector<float>(); }
    std::string one_texture(const std::string& n) const { const Param* p = find(n, "texture"); return p && !p->strs.empty() ? p->strs[0] : std::string(); }
    bool has_texture(const std::string& n) const { for (const Param& p : ps) if (p.name == n && p.type == "texture") return true; return false; }
    // find_one_spectrum with RGB values ("rgb" / "color"); "spectrum" files and "blackbody" are outside this path
    void one_rgb(const std::string& n, const float d[3], float out[3]) const {
        for (const Param& p : ps)
            if (p.name == n && (p.type == "spectrum" || p.type == "xyz")) throw Unsupported("parameter \"" + p.type + " " + n + "\": only rgb / color / blackbody spectra are on this path");
        if (const Param* b = find(n, "blackbody")) {
            if (b->nums.size() < 2) throw Invalid("parameter \"blackbody " + n + "\" needs a temperature and a scale");
            blackbody_rgb(b->nums[0], b->nums[1], out);  // find_one_spectrum takes the first of the (T, scale) pairs
            return;
        }
        const Param* p = find(n, "rgb", "color");
        if (p && p->nums.size() >= 3) { out[0] = p->nums[0]; out[1] = p->nums[1]; out[2] = p->nums[2]; }
        else { out[0] = d[0]; out[1] = d[1]; out[2] = d[2]; }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Loaded scene: owns every array the b200pt_scene_desc points into.
struct ObjectDef {
    std::vector<float> verts, uvs, normals, tangents;
    std::vector<uint32_t> flags;
    std::vector<int32_t> material;
    std::vector<int32_t> alpha_tex;     // 2 per triangle: alpha / shadowalpha float-texture index or -1
    std::vector<b200pt_bvh_node> nodes;
    std::vector<uint32_t> ordered;
    bool any_uv = false, any_n = false, any_s = false;
};
struct Loaded {
    b200pt_scene_desc desc;
    std::string output;                 // Film "filename"
    std::vector<float> verts, uvs, normals, tangents;
    std::vector<uint32_t> flags;
    std::vector<int32_t> material, light;
    std::vector<b200pt_material> materials;
    std::vector<b200pt_light> lights;
    std::vector<std::unique_ptr<std::vector<float>>> images;
    std::vector<b200pt_bvh_node> nodes;
    std::vector<uint32_t> ordered;
    std::vector<uint32_t> sobol;        // SOBOL_MATRICES_32 when the scene asks for the sobol sampler
    std::vector<b200pt_float_texture> float_textures;  // Texture "name" "float" ... used as alpha masks
    std::vector<int32_t> alpha_tex;     // 2 per triangle (top-level triangles first, then the objects')
    std::vector<b200pt_spectrum_texture> spectrum_textures;  // Texture "name" "spectrum" | "color" ... used as a material's Kd
    std::vector<int32_t> material_kd_tex;  // per material: index into spectrum_textures or -1
    std::vector<uint8_t> noise_perm;    // NOISE_PERM[0..256) when a "dots" texture is used
    std::vector<ObjectDef> objects;
    std::vector<b200pt_object> object_descs;
    std::vector<b200pt_instance> instances;
    int max_node_prims = 4;
};

// ---------------------------------------------------------------------------------------------------------------
// PFM (core/src/image_io.rs:227-375: the reference reads/writes .pfm next to .exr/.png/.tga)
static void read_pfm(const std::string& path, std::vector<float>* rgb, int* w, int* h) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Invalid("cannot open image '" + path + "'");
    std::string magic;
    float scale = 0;
    f >> magic >> *w >> *h >> scale;
    if (!f || (magic != "PF" && magic != "Pf") || *w <= 0 || *h <= 0) throw Invalid("'" + path + "' is not a PFM image");
    f.get();
    const int ch = magic == "PF" ? 3 : 1;
    std::vector<float> raw((size_t)*w * *h * ch);
    f.read((char*)raw.data(), (std::streamsize)(raw.size() * 4));
    if (!f) throw Invalid("'" + path + "': truncated PFM data");
    if (scale > 0) {  // big-endian file
        for (float& v : raw) { unsigned char* b = (unsigned char*)&v; std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
    }
    const float s = std::fabs(scale);
    rgb->resize((size_t)*w * *h * 3);
    for (int y = 0; y < *h; ++y)  // PFM rows run bottom to top
        for (int x = 0; x < *w; ++x)
            for (int c = 0; c < 3; ++c) (*rgb)[((size_t)y * *w + x) * 3 + c] = s * raw[((size_t)(*h - 1 - y) * *w + x) * ch + (ch == 3 ? c : 0)];
}
// 8-bit PNG as the reference's read_8_bit takes it (core/src/image_io.rs:192-218: image::open(..).into_rgb8(), value / 255, no
// gamma): non-interlaced, 8 bits per channel, grey / RGB / palette with or without alpha (alpha is dropped).  Rows top to bottom.
static void read_png(const std::string& path, std::vector<float>* rgb, int* w, int* h) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Invalid("cannot open image '" + path + "'");
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) throw Invalid("'" + path + "' is not a PNG image");
    auto be32 = [&](size_t o) { return ((uint32_t)file[o] << 24) | ((uint32_t)file[o + 1] << 16) | ((uint32_t)file[o + 2] << 8) | (uint32_t)file[o + 3]; };
    uint32_t width = 0, height = 0;
    int depth = 0, color = 0, interlace = 0;
    std::vector<unsigned char> idat, plte;
    for (size_t o = 8; o + 12 <= file.size();) {
        const uint32_t len = be32(o);
        if (o + 12 + (size_t)len > file.size()) throw Invalid("'" + path + "': truncated PNG chunk");
        const std::string type((const char*)&file[o + 4], 4);
        const unsigned char* data = &file[o + 8];
        if (type == "IHDR" && len >= 13) { width = be32(o + 8); height = be32(o + 12); depth = data[8]; color = data[9]; interlace = data[12]; }
        else if (type == "PLTE") plte.assign(data, data + len);
        else if (type == "IDAT") idat.insert(idat.end(), data, data + len);
        else if (type == "IEND") break;
        o += 12 + (size_t)len;
    }
    if (width == 0 || height == 0 || width > (1u << 16) || height > (1u << 16)) throw Invalid("'" + path + "': bad PNG header");
    if (depth != 8 || interlace != 0) throw Unsupported("'" + path + "': only non-interlaced PNGs with 8 bits per channel are decoded here");
    const int ch = color == 0 ? 1 : (color == 2 ? 3 : (color == 3 ? 1 : (color == 4 ? 2 : (color == 6 ? 4 : 0))));
    if (ch == 0 || (color == 3 && plte.size() < 3)) throw Invalid("'" + path + "': bad PNG colour type");
    const size_t stride = (size_t)width * ch;
    std::vector<unsigned char> raw((stride + 1) * height);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) throw Invalid("'" + path + "': cannot inflate the PNG data");
    std::vector<unsigned char> px(stride * height);
    for (uint32_t y = 0; y < height; ++y) {  // undo the scanline filters (PNG specification, section 9)
        const unsigned char* in = &raw[(stride + 1) * y];
        unsigned char* cur = &px[stride * y];
        const unsigned char* up = y ? &px[stride * (y - 1)] : nullptr;
        const int ft = in[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = b;
            else if (ft == 3) pred = (a + b) / 2;
            else if (ft == 4) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
            else if (ft != 0) throw Invalid("'" + path + "': bad PNG filter type");
            cur[i] = (unsigned char)(in[1 + i] + pred);
        }
    }
    *w = (int)width; *h = (int)height;
    rgb->resize((size_t)width * height * 3);
    for (size_t k = 0; k < (size_t)width * height; ++k) {
        unsigned char r, g, bl;
        const unsigned char* q = &px[k * ch];
        if (color == 0 || color == 4) r = g = bl = q[0];
        else if (color == 3) { const size_t e = (size_t)q[0] * 3; if (e + 3 > plte.size()) throw Invalid("'" + path + "': PNG palette index out of range"); r = plte[e]; g = plte[e + 1]; bl = plte[e + 2]; }
        else { r = q[0]; g = q[1]; bl = q[2]; }
        (*rgb)[3 * k] = (float)r / 255.0f; (*rgb)[3 * k + 1] = (float)g / 255.0f; (*rgb)[3 * k + 2] = (float)bl / 255.0f;
    }
}
static bool has_ext(const std::string& path, const char* ext) {
    const size_t n = std::strlen(ext);
    if (path.size() < n) return false;
    for (size_t i = 0; i < n; ++i) if (std::tolower((unsigned char)path[path.size() - n + i]) != ext[i]) return false;
    return true;
}
// read_image (core/src/image_io.rs:227-240) for the formats decoded here: .pfm and 8-bit .png
static void read_image(const std::string& path, std::vector<float>* rgb, int* w, int* h) {
    if (has_ext(path, ".pfm")) read_pfm(path, rgb, w, h);
    else if (has_ext(path, ".png")) read_png(path, rgb, w, h);
    else throw Unsupported("image '" + path + "': only .pfm and 8-bit .png images are decoded here (convert .exr / .tga)");
}
static void write_pfm(const std::string& path, const float* rgb, int w, int h) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Invalid("cannot create '" + path + "'");
    f << "PF\n" << w << " " << h << "\n-1.0\n";  // little-endian
    for (int y = h - 1; y >= 0; --y) f.write((const char*)(rgb + (size_t)y * w * 3), (std::streamsize)((size_t)w * 12));
    if (!f) throw Invalid("error writing '" + path + "'");
}

// ---------------------------------------------------------------------------------------------------------------
// PLY (shapes/src/plymesh.rs:25-255): vertex x y z [nx ny nz] [u v | s t | texture_u texture_v | texture_s texture_t]
// (float properties only), face vertex_indices / vertex_index lists of 3 or 4; ascii and binary little / big endian.
struct PlyMesh {
    std::vector<float> p, n, uv;
    std::vector<int> idx;
};
static void read_ply(const std::string& path, PlyMesh* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Invalid("Unable to open PLY file '" + path + "'");
    struct Prop { std::string name, type, count_type, item_type; bool list = false; };
    struct Elem { std::string name; long count = 0; std::vector<Prop> props; };
    std::vector<Elem> elems;
    std::string line, format;
    std::getline(f, line);
    if (line.substr(0, 3) != "ply") throw Invalid("'" + path + "' is not a PLY file");
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string kw;
        ls >> kw;
        if (kw == "format") ls >> format;
        else if (kw == "element") { Elem e; ls >> e.name >> e.count; elems.push_back(e); }
        else if (kw == "property") {
            if (elems.empty()) throw Invalid("PLY: property before element");
            Prop p; std::string t; ls >> t;
            if (t == "list") { p.list = true; ls >> p.count_type >> p.item_type >> p.name; }
            else { p.type = t; ls >> p.name; }
            elems.back().props.push_back(p);
        } else if (kw == "end_header") break;
    }
    const bool ascii = format == "ascii", big = format == "binary_big_endian";
    if (!ascii && !big && format != "binary_little_endian") throw Invalid("PLY: unknown format '" + format + "'");
    auto size_of = [](const std::string& t) -> int {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        throw Invalid("PLY: unknown property type '" + t + "'");
    };
    auto read_num = [&](const std::string& t) -> double {
        if (ascii) { double v; f >> v; if (!f) throw Invalid("PLY: truncated ascii data"); return v; }
        unsigned char b[8];
        int n = size_of(t);
        f.read((char*)b, n);
        if (!f) throw Invalid("PLY: truncated binary data");
        if (big) for (int i = 0; i < n / 2; ++i) std::swap(b[i], b[n - 1 - i]);
        if (t == "char" || t == "int8") return (double)*(signed char*)b;
        if (t == "uchar" || t == "uint8") return (double)*(unsigned char*)b;
        if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
        if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, b, 4); return v; }
        if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
        double v; std::memcpy(&v, b, 8); return v;
    };
    bool has_n = true, has_uv = true;
    long faces = 0;
    for (const Elem& e : elems) {
        for (long k = 0; k < e.count; ++k) {
            float px = 0, py = 0, pz = 0, nx = 0, ny = 0, nz = 0, u = 0, v = 0;
            int nc = 0, uvc = 0;
            for (const Prop& p : e.props) {
                if (p.list) {
                    long cnt = (long)read_num(p.count_type);
                    std::vector<long> items((size_t)cnt);
                    for (long i = 0; i < cnt; ++i) items[(size_t)i] = (long)read_num(p.item_type);
                    const bool integral = p.item_type != "float" && p.item_type != "float32" && p.item_type != "double" && p.item_type != "float64";
                    if (e.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index") && integral) {
                        if (cnt != 3 && cnt != 4) throw Unsupported("PLY: only triangles and quads are supported");
                        out->idx.push_back((int)items[0]); out->idx.push_back((int)items[1]); out->idx.push_back((int)items[2]);
                        if (cnt == 4) { out->idx.push_back((int)items[3]); out->idx.push_back((int)items[0]); out->idx.push_back((int)items[2]); }
                    }
                } else {
                    double val = read_num(p.type);
                    if (e.name != "vertex" || (p.type != "float" && p.type != "float32")) continue;  // Property::Float only
                    float fv = (float)val;
                    if (p.name == "x") px = fv; else if (p.name == "y") py = fv; else if (p.name == "z") pz = fv;
                    else if (p.name == "nx") { nx = fv; ++nc; } else if (p.name == "ny") { ny = fv; ++nc; } else if (p.name == "nz") { nz = fv; ++nc; }
                    else if (p.name == "u" || p.name == "s" || p.name == "texture_u" || p.name == "texture_s") { u = fv; ++uvc; }
                    else if (p.name == "v" || p.name == "t" || p.name == "texture_v" || p.name == "texture_t") { v = fv; ++uvc; }
                }
            }
            if (e.name == "vertex") {
                out->p.push_back(px); out->p.push_back(py); out->p.push_back(pz);
                has_n = has_n && nc == 3;
                if (has_n) { out->n.push_back(nx); out->n.push_back(ny); out->n.push_back(nz); }
                has_uv = has_uv && uvc == 2;
                if (has_uv) { out->uv.push_back(u); out->uv.push_back(v); }
            } else if (e.name == "face") ++faces;
        }
    }
    if (out->p.empty() || faces == 0) throw Invalid("PLY file '" + path + "' is invalid: no face / vertex elements");
    if (!has_n) out->n.clear();
    if (!has_uv) out->uv.clear();
}

// ---------------------------------------------------------------------------------------------------------------
// The API state machine (api/src/lib.rs)
struct GState {
    Xf ctm = xf_identity();
    bool reverse = false;
    int material = -1;  // index into Loaded::materials; -1 = the default matte, created on first use
    bool has_area = false;
    float area_L[3] = {1, 1, 1};
    bool area_two_sided = false;
    std::map<std::string, int> float_textures;  // GraphicsState::float_textures: name -> index into Loaded::float_textures (scoped by Attribute blocks)
    std::map<std::string, int> spectrum_textures;  // GraphicsState::spectrum_textures, likewise
};

struct Builder {
    Loaded* L;
    std::string dir;
    std::vector<GState> stack;
    std::vector<Xf> xf_stack;
    GState gs;
    std::map<std::string, Xf> named_cs;
    std::map<std::string, int> named_materials;
    std::map<std::string, int> object_ids;
    int cur_object = -1;
    bool in_world = false;
    // options
    std::string camera_name = "perspective", sampler_name = "halton", filter_name = "box", integrator_name = "path", film_name = "image", accel_name = "bvh";
    ParamSet camera_p, sampler_p, filter_p, integrator_p, film_p, accel_p;
    Xf camera_to_world = xf_identity();
    int default_matte = -1;
    int include_depth = 0;

    std::string resolve(const std::string& p) const { return (!p.empty() && p[0] == '/') || dir.empty() ? p : dir + "/" + p; }

    static void fill_rgb(float* dst, const float* src) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }

    static const int kNoMaterial = -2;  // gs.material: -1 = the default matte (not created yet), kNoMaterial = Material "none"
    int make_material(const std::string& type, const ParamSet& p) {
        for (const char* t : {"Ks", "Kr", "Kt", "eta", "k", "sigma", "roughness", "uroughness", "vroughness", "index", "bumpmap"})
            if (p.has_texture(t)) throw Unsupported(std::string("material parameter \"texture ") + t + "\": textures are outside this path (constant values, and a textured Kd)");
        // "texture Kd": TextureParams::get_spectrum_texture_or_else looks the name up among the spectrum textures in scope
        // (an unknown name falls back to the constant parameter / default, texture_params.rs)
        int kd_tex = -1;
        if (p.has_texture("Kd")) {
            if (type != "matte" && type != "plastic") throw Unsupported("\"texture Kd\" on Material \"" + type + "\" is outside this path");
            auto it = gs.spectrum_textures.find(p.one_texture("Kd"));
            if (it != gs.spectrum_textures.end()) kd_tex = it->second;
        }
        b200pt_material m;
        std::memset(&m, 0, sizeof(m));
        m.remap_roughness = p.one_bool("remaproughness", true) ? 1 : 0;
        const float one[3] = {1, 1, 1}, half[3] = {0.5f, 0.5f, 0.5f}, quarter[3] = {0.25f, 0.25f, 0.25f};
        // GraphicsState::make_material: "none" | "" => None (api/src/graphics_state.rs:336): the primitive gets no material and
        // the path integrator passes through it (integrators/src/path.rs:141-150)
        if (type == "none" || type == "") return kNoMaterial;
        if (type == "matte") {
            m.type = B200PT_MAT_MATTE;
            p.one_rgb("Kd", half, m.kd);
            m.sigma = p.one_float("sigma", 0.0f);
        } else if (type == "plastic") {
            m.type = B200PT_MAT_PLASTIC;
            p.one_rgb("Kd", quarter, m.kd);
            p.one_rgb("Ks", quarter, m.ks);
            m.urough = m.vrough = p.one_float("roughness", 0.1f);
        } else if (type == "glass") {
            m.type = B200PT_MAT_GLASS;
            p.one_rgb("Kr", one, m.ks);
            p.one_rgb("Kt", one, m.kt);
            m.eta[0] = p.find("eta", "float") ? p.one_float("eta", 1.5f) : p.one_float("index", 1.5f);
            m.urough = p.one_float("uroughness", 0.0f);
            m.vrough = p.one_float("vroughness", 0.0f);
        } else if (type == "mirror") {  // mirror.rs:62-70
            m.type = B200PT_MAT_MIRROR;
            const float kr[3] = {0.9f, 0.9f, 0.9f};
            p.one_rgb("Kr", kr, m.ks);
        } else if (type == "metal") {
            m.type = B200PT_MAT_METAL;
            // copper SPD -> RGB as the reference's Spectrum::from(&samples).to_rgb() evaluates it (materials/src/metal.rs:109-133)
            const float cu_eta[3] = {0.19999069f, 0.92208463f, 1.09987593f}, cu_k[3] = {3.90463543f, 2.44763327f, 2.13765264f};
            p.one_rgb("eta", cu_eta, m.eta);
            p.one_rgb("k", cu_k, m.k);
            float r = p.one_float("roughness", 0.01f);
            m.urough = p.find("uroughness", "float") ? p.one_float("uroughness", r) : r;
            m.vrough = p.find("vroughness", "float") ? p.one_float("vroughness", r) : r;
        } else throw Unsupported("Material \"" + type + "\" is outside this path (matte, plastic, glass, metal, mirror)");
        if (kd_tex >= 0 && L->spectrum_textures[(size_t)kd_tex].type == B200PT_STEX_CONSTANT) {  // a constant texture is just the value
            fill_rgb(m.kd, L->spectrum_textures[(size_t)kd_tex].tex1);
            kd_tex = -1;
        }
        L->materials.push_back(m);
        L->material_kd_tex.push_back(kd_tex);
        return (int)L->materials.size() - 1;
    }
    int current_material() {  // index into materials, or -1 for a primitive without a material
        if (gs.material == kNoMaterial) return -1;
        if (gs.material >= 0) return gs.material;
        if (default_matte < 0) default_matte = make_material("matte", ParamSet());  // GraphicsState's default material
        return default_matte;
    }

    std::string data_dir() const { return lib_data_dir(); }
    // Texture "name" "float" "class": the float textures a mesh can use as alpha / shadowalpha (api/src/lib.rs pbrt_texture ->
    // make_float_texture; textures/src/{constant,checkerboard_2d,dots,imagemap}.rs from-params).  Spectrum textures feed
    // materials, which take constants only on this path.
    // Texture "name" "spectrum" | "color" "class": what a matte / plastic material can name as "texture Kd"
    // (make_spectrum_texture; textures/src/{constant,checkerboard_2d}.rs from-params with constant tex1 / tex2).
    void spectrum_texture(const std::string& name, const std::string& cls, const ParamSet& p) {
        b200pt_spectrum_texture t;
        std::memset(&t, 0, sizeof(t));
        auto const_or_named = [&](const char* pn, float dflt, float out[3]) {  // get_spectrum_texture_or_else with constant sub-textures
            const std::string tn = p.one_texture(pn);
            if (!tn.empty()) {
                auto it = gs.spectrum_textures.find(tn);
                if (it != gs.spectrum_textures.end()) {
                    const b200pt_spectrum_texture& s = L->spectrum_textures[(size_t)it->second];
                    if (s.type != B200PT_STEX_CONSTANT) throw Unsupported("Texture \"" + name + "\": nested non-constant textures are outside this path");
                    fill_rgb(out, s.tex1);
                    return;
                }
            }
            const float d[3] = {dflt, dflt, dflt};
            p.one_rgb(pn, d, out);
        };
        t.su = t.sv = 1.0f;
        if (cls == "constant") { t.type = B200PT_STEX_CONSTANT; const_or_named("value", 1.0f, t.tex1); }
        else if (cls == "checkerboard") {
            if (p.one_int("dimension", 2) != 2) throw Unsupported("Texture \"" + name + "\": 3-D checkerboards are outside this path");
            const std::string mapping = p.one_string("mapping", "uv");
            if (mapping != "uv") throw Unsupported("Texture \"" + name + "\": mapping \"" + mapping + "\" is outside this path (uv)");
            t.type = B200PT_STEX_CHECKERBOARD;
            t.su = p.one_float("uscale", 1.0f); t.sv = p.one_float("vscale", 1.0f); t.du = p.one_float("udelta", 0.0f); t.dv = p.one_float("vdelta", 0.0f);
            const_or_named("tex1", 1.0f, t.tex1); const_or_named("tex2", 0.0f, t.tex2);
            t.aa_closedform = p.one_string("aamode", "closedform") == "none" ? 0 : 1;  // anything else warns and means closedform (checkerboard_2d.rs:134-145)
        } else throw Unsupported("Texture \"" + name + "\" \"spectrum\" \"" + cls + "\" is outside this path (constant, checkerboard)");
        L->spectrum_textures.push_back(t);
        gs.spectrum_textures[name] = (int)L->spectrum_textures.size() - 1;
    }

    void texture(const std::string& name, const std::string& type, const std::string& cls, const ParamSet& p) {
        if (type == "spectrum" || type == "color") { spectrum_texture(name, cls, p); return; }
        if (type != "float") throw Unsupported("Texture \"" + name + "\" \"" + type + "\": only float, spectrum and color textures exist");
        b200pt_float_texture t;
        std::memset(&t, 0, sizeof(t));
        if (cls != "constant") {
            const std::string mapping = p.one_string("mapping", "uv");
            if (mapping != "uv") throw Unsupported("Texture \"" + name + "\": mapping \"" + mapping + "\" is outside this path (uv)");
        }
        t.su = p.one_float("uscale", 1.0f); t.sv = p.one_float("vscale", 1.0f); t.du = p.one_float("udelta", 0.0f); t.dv = p.one_float("vdelta", 0.0f);
        auto const_or_named = [&](const char* pn, float dflt) {  // get_float_texture_or_else with constant sub-textures
            const std::string tn = p.one_texture(pn);
            if (!tn.empty()) {
                auto it = gs.float_textures.find(tn);
                if (it != gs.float_textures.end()) {
                    const b200pt_float_texture& s = L->float_textures[(size_t)it->second];
                    if (s.type != B200PT_TEX_CONSTANT) throw Unsupported("Texture \"" + name + "\": nested non-constant textures are outside this path");
                    return s.value[0];
                }
            }
            return p.one_float(pn, dflt);
        };
        if (cls == "constant") { t.type = B200PT_TEX_CONSTANT; t.value[0] = const_or_named("value", 1.0f); }
        else if (cls == "checkerboard") {
            if (p.one_int("dimension", 2) != 2) throw Unsupported("Texture \"" + name + "\": 3-D checkerboards are outside this path");
            t.type = B200PT_TEX_CHECKERBOARD;
            t.value[0] = const_or_named("tex1", 1.0f); t.value[1] = const_or_named("tex2", 0.0f);
        } else if (cls == "dots") {
            // dots.rs:82-86 hands (inside, outside) to new(outside_dot, inside_dot): value[0] is what the texture returns OUTSIDE the dots
            t.type = B200PT_TEX_DOTS;
            t.value[0] = const_or_named("inside", 1.0f); t.value[1] = const_or_named("outside", 0.0f);
            if (L->noise_perm.empty()) {
                std::ifstream f(data_dir() + "/noise_perm.bin", std::ios::binary);
                L->noise_perm.resize(256);
                if (!f || !f.read((char*)L->noise_perm.data(), 256)) throw Invalid("Texture \"dots\": cannot read " + data_dir() + "/noise_perm.bin (set B200PT_DATA_DIR)");
            }
        } else if (cls == "imagemap") {
            // imagemap.rs:104-141; images: PFM only on this path, power-of-two sides (MIPMap::new would resample others)
            std::string fn = p.one_string("filename", "");
            if (fn.empty()) throw Invalid("Texture \"" + name + "\": imagemap without \"filename\"");
            const std::string path = resolve(fn);
            std::vector<float> rgb;
            int w = 0, h = 0;
            read_image(path, &rgb, &w, &h);
            if ((w & (w - 1)) || (h & (h - 1))) throw Unsupported("Texture \"" + name + "\": image sides must be powers of two on this path (no resampling)");
            const float scale = p.one_float("scale", 1.0f);
            const bool gamma = p.one_bool("gamma", has_ext(path, ".png") || has_ext(path, ".tga"));  // imagemap.rs:129
            const std::string wrap = p.one_string("wrap", "repeat");
            t.type = B200PT_TEX_IMAGEMAP;
            t.wrap = wrap == "black" ? 1 : (wrap == "clamp" ? 2 : 0);
            t.width = w; t.height = h;
            auto tex = std::make_unique<std::vector<float>>((size_t)w * h);
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    // generate_mipmap (mipmap/cache.rs): texel (x, y) of the image lands in row h - 1 - y (texture space has
                    // t = 0 at the bottom); convert_in::<Float>: scale * (gamma ? inv_gamma(y()) : y())
                    const float* c = &rgb[((size_t)y * w + x) * 3];
                    float lum = 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2];
                    if (gamma) lum = lum <= 0.04045f ? lum * 1.0f / 12.92f : std::pow((lum + 0.055f) * 1.0f / 1.055f, 2.4f);  // inv_gamma_correct, pbrt/common.rs:152-158
                    (*tex)[(size_t)(h - 1 - y) * w + x] = scale * lum;
                }
            t.texels = tex->data();
            L->images.push_back(std::move(tex));
        } else throw Unsupported("Texture \"" + name + "\" \"float\" \"" + cls + "\" is outside this path (constant, checkerboard, dots, imagemap)");
        L->float_textures.push_back(t);
        gs.float_textures[name] = (int)L->float_textures.size() - 1;
    }

    void add_triangles(const std::vector<int>& idx, const std::vector<float>& P, const std::vector<float>& N, const std::vector<float>& S,
                       const std::vector<float>& UV, const ParamSet& params) {
        // triangle.rs:278-312: "texture alpha" names a float texture; an unknown name falls back to the float parameter
        float alpha = params.one_float("alpha", 1.0f), shadow_alpha = params.one_float("shadowalpha", 1.0f);
        int alpha_tex[2] = {-1, -1};
        const char* names[2] = {"alpha", "shadowalpha"};
        for (int c = 0; c < 2; ++c) {
            const std::string tn = params.one_texture(names[c]);
            if (tn.empty()) continue;
            auto it = gs.float_textures.find(tn);
            if (it == gs.float_textures.end()) continue;
            const b200pt_float_texture& ft = L->float_textures[(size_t)it->second];
            if (ft.type == B200PT_TEX_CONSTANT) (c == 0 ? alpha : shadow_alpha) = ft.value[0];
            else alpha_tex[c] = it->second;
        }
        const size_t np = P.size() / 3;
        for (int i : idx) if (i < 0 || (size_t)i >= np) throw Invalid("triangle mesh has an out-of-bounds vertex index");
        const Xf& o2w = gs.ctm;
        const bool flip = gs.reverse != swaps_handedness(o2w.m);
        uint32_t fl = (flip ? B200PT_PRIM_FLIP_NORMAL : 0u) | (alpha == 0.0f ? B200PT_PRIM_ALPHA_ZERO : 0u) | (shadow_alpha == 0.0f ? B200PT_PRIM_SHADOW_ALPHA_ZERO : 0u) |
                      (gs.reverse ? B200PT_PRIM_REVERSE_ORIENTATION : 0u) | (!UV.empty() ? B200PT_PRIM_HAS_UV : 0u) | (!N.empty() ? B200PT_PRIM_HAS_NORMALS : 0u) |
                      (!S.empty() ? B200PT_PRIM_HAS_TANGENTS : 0u) | ((alpha_tex[0] >= 0 || alpha_tex[1] >= 0) ? B200PT_PRIM_ALPHA_TEXTURE : 0u);
        // TriangleMesh::new: everything to world space once (triangle.rs:92-99)
        std::vector<V3> wp(np), wn(N.empty() ? 0 : np), ws(S.empty() ? 0 : np);
        for (size_t i = 0; i < np; ++i) wp[i] = xf_point(o2w.m, v3(P[3 * i], P[3 * i + 1], P[3 * i + 2]));
        for (size_t i = 0; i < wn.size(); ++i) wn[i] = xf_normal(o2w.inv, v3(N[3 * i], N[3 * i + 1], N[3 * i + 2]));
        for (size_t i = 0; i < ws.size(); ++i) ws[i] = xf_vector(o2w.m, v3(S[3 * i], S[3 * i + 1], S[3 * i + 2]));
        const int mat = current_material();
        const bool into_object = cur_object >= 0;
        if (into_object && gs.has_area) throw Unsupported("area lights inside ObjectBegin/ObjectEnd are not supported (as in pbrt)");
        std::vector<float>& V = into_object ? L->objects[(size_t)cur_object].verts : L->verts;
        std::vector<float>& U = into_object ? L->objects[(size_t)cur_object].uvs : L->uvs;
        std::vector<float>& Nn = into_object ? L->objects[(size_t)cur_object].normals : L->normals;
        std::vector<float>& Ss = into_object ? L->objects[(size_t)cur_object].tangents : L->tangents;
        std::vector<uint32_t>& F = into_object ? L->objects[(size_t)cur_object].flags : L->flags;
        std::vector<int32_t>& M = into_object ? L->objects[(size_t)cur_object].material : L->material;
        std::vector<int32_t>& AT = into_object ? L->objects[(size_t)cur_object].alpha_tex : L->alpha_tex;
        if (into_object) {
            ObjectDef& o = L->objects[(size_t)cur_object];
            o.any_uv |= !UV.empty(); o.any_n |= !N.empty(); o.any_s |= !S.empty();
        }
        for (size_t t = 0; t + 2 < idx.size(); t += 3) {
            for (int k = 0; k < 3; ++k) {
                const size_t vi = (size_t)idx[t + k];
                V.push_back(wp[vi].x); V.push_back(wp[vi].y); V.push_back(wp[vi].z);
                U.push_back(UV.empty() ? 0.0f : UV[2 * vi]); U.push_back(UV.empty() ? 0.0f : UV[2 * vi + 1]);
                Nn.push_back(wn.empty() ? 0.0f : wn[vi].x); Nn.push_back(wn.empty() ? 0.0f : wn[vi].y); Nn.push_back(wn.empty() ? 0.0f : wn[vi].z);
                Ss.push_back(ws.empty() ? 0.0f : ws[vi].x); Ss.push_back(ws.empty() ? 0.0f : ws[vi].y); Ss.push_back(ws.empty() ? 0.0f : ws[vi].z);
            }
            F.push_back(fl);
            M.push_back(mat);
            AT.push_back(alpha_tex[0]); AT.push_back(alpha_tex[1]);
            if (!into_object) {
                int light = -1;
                if (gs.has_area) {  // one DiffuseAreaLight per triangle, api/src/lib.rs:783-803
                    b200pt_light l;
                    std::memset(&l, 0, sizeof(l));
                    l.type = B200PT_LIGHT_AREA;
                    fill_rgb(l.L, gs.area_L);
                    l.prim = (int32_t)(L->flags.size() - 1);
                    l.two_sided = gs.area_two_sided ? 1 : 0;
                    M4 id = m4_identity();
                    std::memcpy(l.light_to_world, id.m, 64); std::memcpy(l.world_to_light, id.m, 64);
                    L->lights.push_back(l);
                    light = (int)L->lights.size() - 1;
                }
                L->light.push_back(light);
            }
        }
    }

    void shape(const std::string& name, const ParamSet& p) {
        if (name == "trianglemesh") {
            std::vector<int> idx;
            if (const Param* q = p.find("indices", "integer")) for (int32_t v : q->ints) idx.push_back((int)v);
            std::vector<float> P, N, S, UV;
            if (const Param* q = p.find("P", "point", "point3")) P = q->nums;
            if (const Param* q = p.find("N", "normal", "normal3")) N = q->nums;
            if (const Param* q = p.find("S", "vector", "vector3")) S = q->nums;
            const Param* q = p.find("uv", "point2", "float");
            if (!q) q = p.find("st", "point2", "float");
            if (q) UV = q->nums;
            if (idx.empty()) throw Invalid("Vertex indices 'indices' not provided with triangle mesh shape");
            if (P.empty()) throw Invalid("Vertex positions 'P' not provided with triangle mesh shape");
            const size_t np = P.size() / 3;
            if (!UV.empty() && UV.size() / 2 < np) UV.clear();  // "Not enough of 'uv' ... Discarding" (triangle.rs:217-224)
            if (!S.empty() && S.size() / 3 != np) S.clear();
            if (!N.empty() && N.size() / 3 != np) N.clear();
            add_triangles(idx, P, N, S, UV, p);
        } else if (name == "plymesh") {
            std::string fn = p.one_string("filename", "");
            if (fn.empty()) throw Invalid("PLY filename not provided");
            PlyMesh m;
            read_ply(resolve(fn), &m);
            add_triangles(m.idx, m.p, m.n, std::vector<float>(), m.uv, p);
        } else throw Unsupported("Shape \"" + name + "\" is outside this path (trianglemesh, plymesh)");
    }

    void light_source(const std::string& name, const ParamSet& p) {
        b200pt_light l;
        std::memset(&l, 0, sizeof(l));
        l.prim = -1;
        const float one[3] = {1, 1, 1};
        float sc[3];
        p.one_rgb("scale", one, sc);
        if (name == "point") {
            l.type = B200PT_LIGHT_POINT;
            float I[3];
            p.one_rgb("I", one, I);
            for (int c = 0; c < 3; ++c) l.L[c] = I[c] * sc[c];
            V3 from = v3(0, 0, 0);
            if (const Param* q = p.find("from", "point", "point3")) if (q->nums.size() >= 3) from = v3(q->nums[0], q->nums[1], q->nums[2]);
            Xf l2w = xf_mul(xf_translate(from.x, from.y, from.z), gs.ctm);  // point.rs:157
            V3 pl = xf_point(l2w.m, v3(0, 0, 0));
            l.pos[0] = pl.x; l.pos[1] = pl.y; l.pos[2] = pl.z;
            std::memcpy(l.light_to_world, l2w.m.m, 64); std::memcpy(l.world_to_light, l2w.inv.m, 64);
        } else if (name == "infinite" || name == "exinfinite") {
            l.type = B200PT_LIGHT_INFINITE;
            float Lv[3];
            p.one_rgb("L", one, Lv);
            for (int c = 0; c < 3; ++c) l.L[c] = Lv[c] * sc[c];
            std::memcpy(l.light_to_world, gs.ctm.m.m, 64); std::memcpy(l.world_to_light, gs.ctm.inv.m, 64);
            std::string map = p.one_string("mapname", "");
            if (!map.empty()) {
                auto img = std::make_unique<std::vector<float>>();
                int w = 0, h = 0;
                read_image(resolve(map), img.get(), &w, &h);
                l.map_rgb = img->data(); l.map_width = w; l.map_height = h;
                L->images.push_back(std::move(img));
            }
        } else if (name == "distant") {  // distant.rs:131-143 + DistantLight::new :44-52
            l.type = B200PT_LIGHT_DISTANT;
            float Lv[3];
            p.one_rgb("L", one, Lv);
            for (int c = 0; c < 3; ++c) l.L[c] = Lv[c] * sc[c];
            V3 from = v3(0, 0, 0), to = v3(0, 0, 1);
            if (const Param* q = p.find("from", "point", "point3")) if (q->nums.size() >= 3) from = v3(q->nums[0], q->nums[1], q->nums[2]);
            if (const Param* q = p.find("to", "point", "point3")) if (q->nums.size() >= 3) to = v3(q->nums[0], q->nums[1], q->nums[2]);
            V3 w = xf_vector(gs.ctm.m, v3(from.x - to.x, from.y - to.y, from.z - to.z));
            const float len = std::sqrt(w.x * w.x + w.y * w.y + w.z * w.z);
            const float inv = 1.0f / len;  // Vector3::normalize = self / length = self * (1 / length)
            l.pos[0] = inv * w.x; l.pos[1] = inv * w.y; l.pos[2] = inv * w.z;
            std::memcpy(l.light_to_world, gs.ctm.m.m, 64); std::memcpy(l.world_to_light, gs.ctm.inv.m, 64);
        } else if (name == "goniometric" || name == "projection") {  // goniometric.rs:219-238 + GonioPhotometricLight::new :59-100; projection.rs:266-288 + ::new :50-110
            l.type = name == "projection" ? B200PT_LIGHT_PROJECTION : B200PT_LIGHT_GONIOMETRIC;
            float I[3];
            p.one_rgb("I", one, I);
            for (int c = 0; c < 3; ++c) l.L[c] = I[c] * sc[c];
            const V3 pl = xf_point(gs.ctm.m, v3(0, 0, 0));
            l.pos[0] = pl.x; l.pos[1] = pl.y; l.pos[2] = pl.z;
            std::memcpy(l.light_to_world, gs.ctm.m.m, 64); std::memcpy(l.world_to_light, gs.ctm.inv.m, 64);
            std::string map = p.one_string("mapname", "");
            if (!map.empty()) {
                auto img = std::make_unique<std::vector<float>>();
                int w = 0, h = 0;
                read_image(resolve(map), img.get(), &w, &h);
                l.map_rgb = img->data(); l.map_width = w; l.map_height = h;
                L->images.push_back(std::move(img));
            }
            if (name == "projection") {
                l.fov = p.one_float("fov", 45.0f);
                const float aspect = l.map_rgb ? (float)l.map_width / (float)l.map_height : 1.0f;
                const float cx = aspect > 1.0f ? aspect : 1.0f, cy = aspect > 1.0f ? 1.0f : 1.0f / aspect;  // screen_bounds.p_max
                const Xf s2l = xf_inverse(xf_perspective(l.fov, 1e-3f, 1e30f));
                const V3 wc = xf_point(s2l.m, v3(cx, cy, 0.0f));
                const float inv = 1.0f / std::sqrt(wc.x * wc.x + wc.y * wc.y + wc.z * wc.z);
                l.cos_total_width = wc.z * inv;  // Vector3::normalize = self * (1 / length)
            }
        } else if (name == "spot") {  // spot.rs:200-237 + SpotLight::new :44-60
            l.type = B200PT_LIGHT_SPOT;
            float I[3];
            p.one_rgb("I", one, I);
            for (int c = 0; c < 3; ++c) l.L[c] = I[c] * sc[c];
            const float cone_angle = p.one_float("coneangle", 30.0f);
            const float cone_delta = p.one_float("conedeltaangle", 5.0f);  // the reference reads "conedeltaangle" (pbrt-v3 scenes' "conedelta" is ignored, spot.rs:208)
            V3 from = v3(0, 0, 0), to = v3(0, 0, 1);
            if (const Param* q = p.find("from", "point", "point3")) if (q->nums.size() >= 3) from = v3(q->nums[0], q->nums[1], q->nums[2]);
            if (const Param* q = p.find("to", "point", "point3")) if (q->nums.size() >= 3) to = v3(q->nums[0], q->nums[1], q->nums[2]);
            V3 dir = v3(to.x - from.x, to.y - from.y, to.z - from.z);
            const float dinv = 1.0f / std::sqrt(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);  // Vector3::normalize = self * (1 / length)
            dir = v3(dir.x * dinv, dir.y * dinv, dir.z * dinv);
            V3 du, dv;  // coordinate_system(&dir), coordinate_system.rs:12-20
            if (std::fabs(dir.x) > std::fabs(dir.y)) { const float k = 1.0f / std::sqrt(dir.x * dir.x + dir.z * dir.z); du = v3(-dir.z * k, 0.0f * k, dir.x * k); }
            else { const float k = 1.0f / std::sqrt(dir.y * dir.y + dir.z * dir.z); du = v3(0.0f * k, dir.z * k, -dir.y * k); }
            dv = v3((dir.y * du.z) - (dir.z * du.y), (dir.z * du.x) - (dir.x * du.z), (dir.x * du.y) - (dir.y * du.x));
            M4 dz = m4_identity();
            dz.m[0][0] = du.x; dz.m[0][1] = du.y; dz.m[0][2] = du.z;
            dz.m[1][0] = dv.x; dz.m[1][1] = dv.y; dz.m[1][2] = dv.z;
            dz.m[2][0] = dir.x; dz.m[2][1] = dir.y; dz.m[2][2] = dir.z;
            const Xf l2w = xf_mul(xf_mul(gs.ctm, xf_translate(from.x, from.y, from.z)), xf_inverse(xf_from(dz)));
            const V3 pl = xf_point(l2w.m, v3(0, 0, 0));
            l.pos[0] = pl.x; l.pos[1] = pl.y; l.pos[2] = pl.z;
            std::memcpy(l.light_to_world, l2w.m.m, 64); std::memcpy(l.world_to_light, l2w.inv.m, 64);
            const float rad = 3.14159265358979323846f / 180.0f;  // f32::to_radians
            l.cos_total_width = std::cos(cone_angle * rad);
            l.cos_falloff_start = std::cos((cone_angle - cone_delta) * rad);
        } else throw Unsupported("LightSource \"" + name + "\" is not one of the reference's lights (point, spot, projection, goniometric, distant, infinite)");
        L->lights.push_back(l);
    }

    void finish();
};

static float gaussian_1d(float d, float expv, float alpha) { float g = std::exp(-alpha * d * d) - expv; return g > 0.0f ? g : 0.0f; }

void Builder::finish() {
    b200pt_scene_desc& d = L->desc;
    std::memset(&d, 0, sizeof(d));
    // --- Film + filter (film/mod.rs:89-146, 420-485) ---
    if (film_name != "image") throw Unsupported("Film \"" + film_name + "\"");
    const int xres = film_p.one_int("xresolution", 1280), yres = film_p.one_int("yresolution", 720);
    float crop[4] = {0, 1, 0, 1};
    std::vector<float> cw = film_p.floats("cropwindow");
    auto clamp01 = [](float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); };
    if (cw.size() == 4) {
        crop[0] = clamp01(cw[0] < cw[1] ? cw[0] : cw[1]); crop[1] = clamp01(cw[0] > cw[1] ? cw[0] : cw[1]);
        crop[2] = clamp01(cw[2] < cw[3] ? cw[2] : cw[3]); crop[3] = clamp01(cw[2] > cw[3] ? cw[2] : cw[3]);
    } else if (!cw.empty()) throw Invalid("'cropwindow' expects 4 values");
    d.film.xres = xres; d.film.yres = yres;
    d.film.crop[0] = (int)std::ceil((float)xres * crop[0]); d.film.crop[1] = (int)std::ceil((float)yres * crop[2]);
    d.film.crop[2] = (int)std::ceil((float)xres * crop[1]); d.film.crop[3] = (int)std::ceil((float)yres * crop[3]);
    d.film.scale = film_p.one_float("scale", 1.0f);
    d.film.max_sample_luminance = film_p.one_float("maxsampleluminance", INFINITY);
    L->output = film_p.one_string("filename", "pbrt.exr");
    float rx, ry;
    if (filter_name == "box") {
        rx = filter_p.one_float("xwidth", 0.5f); ry = filter_p.one_float("ywidth", 0.5f);
        for (int i = 0; i < 256; ++i) d.film.filter_table[i] = 1.0f;
    } else if (filter_name == "gaussian") {
        rx = filter_p.one_float("xwidth", 2.0f); ry = filter_p.one_float("ywidth", 2.0f);
        const float alpha = filter_p.one_float("alpha", 2.0f);
        const float ex = std::exp(-alpha * rx * rx), ey = std::exp(-alpha * ry * ry);
        int k = 0;
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                float px = ((float)x + 0.5f) * rx * (1.0f / 16.0f), py = ((float)y + 0.5f) * ry * (1.0f / 16.0f);
                d.film.filter_table[k++] = gaussian_1d(px, ex, alpha) * gaussian_1d(py, ey, alpha);
            }
    } else if (filter_name == "triangle" || filter_name == "mitchell" || filter_name == "sinc") {
        // filters/src/{triangle,mitchell,sinc}.rs evaluated at the table points of Film::new (film/mod.rs:113-125)
        const float dflt = filter_name == "sinc" ? 4.0f : 2.0f;
        rx = filter_p.one_float("xwidth", dflt); ry = filter_p.one_float("ywidth", dflt);
        const float B = filter_p.one_float("B", 1.0f / 3.0f), Cc = filter_p.one_float("C", 1.0f / 3.0f), tau = filter_p.one_float("tau", 3.0f);
        const float irx = 1.0f / rx, iry = 1.0f / ry;
        auto mitchell_1d = [&](float x) {  // mitchell.rs:41-57 (incl. its 8 C + 24 C constant term)
            x = std::fabs(2.0f * x);
            if (x > 1.0f) return ((-B - 6.0f * Cc) * x * x * x + (6.0f * B + 30.0f * Cc) * x * x + (-12.0f * B - 48.0f * Cc) * x + (8.0f * Cc + 24.0f * Cc)) * (1.0f / 6.0f);
            return ((12.0f - 9.0f * B - 6.0f * Cc) * x * x * x + (-18.0f + 12.0f * B + 6.0f * Cc) * x * x + (6.0f - 2.0f * B)) * (1.0f / 6.0f);
        };
        auto sinc = [](float x) { x = std::fabs(x); return x < 1e-5f ? 1.0f : std::sin(3.14159265358979323846f * x) / (3.14159265358979323846f * x); };  // sinc.rs:72-79
        auto windowed_sinc = [&](float x, float radius) { x = std::fabs(x); if (x > radius) return 0.0f; const float lanczos = sinc(x / tau); return sinc(x) * lanczos; };
        int k = 0;
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) {
                const float px = ((float)x + 0.5f) * rx * (1.0f / 16.0f), py = ((float)y + 0.5f) * ry * (1.0f / 16.0f);
                float v;
                if (filter_name == "triangle") { const float a = rx - std::fabs(px), b = ry - std::fabs(py); v = (0.0f > a ? 0.0f : a) * (0.0f > b ? 0.0f : b); }
                else if (filter_name == "mitchell") v = mitchell_1d(px * irx) * mitchell_1d(py * iry);
                else v = windowed_sinc(px, rx) * windowed_sinc(py, ry);
                d.film.filter_table[k++] = v;
            }
    } else throw Unsupported("PixelFilter \"" + filter_name + "\" is not one of the reference's filters (box, gaussian, triangle, mitchell, sinc)");
    d.film.filter_radius[0] = rx; d.film.filter_radius[1] = ry;

    // --- Camera (perspective_camera.rs:35-75, 357-421; camera.rs:276-306) ---
    if (camera_name != "perspective" && camera_name != "orthographic" && camera_name != "environment")
        throw Unsupported("Camera \"" + camera_name + "\" is outside this path (perspective, orthographic, environment)");
    d.camera.type = camera_name == "perspective" ? B200PT_CAMERA_PERSPECTIVE : (camera_name == "orthographic" ? B200PT_CAMERA_ORTHOGRAPHIC : B200PT_CAMERA_ENVIRONMENT);
    float so = camera_p.one_float("shutteropen", 0.0f), sc = camera_p.one_float("shutterclose", 1.0f);
    if (sc < so) std::swap(so, sc);
    d.camera.shutter_open = so; d.camera.shutter_close = sc;
    // EnvironmentCamera::from reads the shutter only (environment_camera.rs:62-80)
    d.camera.lens_radius = camera_name == "environment" ? 0.0f : camera_p.one_float("lensradius", 0.0f);
    d.camera.focal_distance = camera_p.one_float("focaldistance", 1e6f);
    const float frame = camera_p.one_float("frameaspectratio", (float)xres / (float)yres);
    float sw[4];  // x0 x1 y0 y1
    if (frame > 1.0f) { sw[0] = -frame; sw[1] = frame; sw[2] = -1.0f; sw[3] = 1.0f; }
    else { sw[0] = -1.0f; sw[1] = 1.0f; sw[2] = -1.0f / frame; sw[3] = 1.0f / frame; }
    std::vector<float> swp = camera_p.floats("screenwindow");
    if (swp.size() == 4) for (int i = 0; i < 4; ++i) sw[i] = swp[(size_t)i];
    float fov = camera_p.one_float("fov", 90.0f);
    const float half_fov = camera_p.one_float("halffov", -1.0f);
    if (half_fov > 0.0f) fov = 2.0f * half_fov;
    // PerspectiveCamera: Transform::perspective(fov, 1e-2, 1000); OrthographicCamera: Transform::orthographic(0, 1) =
    // scale(1, 1, 1 / (far - near)) * translate(0, 0, -near) (transform.rs:222-225)
    Xf c2s = camera_name == "orthographic" ? xf_mul(xf_scale(1.0f, 1.0f, 1.0f / (1.0f - 0.0f)), xf_translate(0.0f, 0.0f, -0.0f)) : xf_perspective(fov, 1e-2f, 1000.0f);
    Xf s2r = xf_mul(xf_mul(xf_scale((float)xres, (float)yres, 1.0f), xf_scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f)), xf_translate(-sw[0], -sw[3], 0.0f));
    Xf r2c = xf_mul(xf_inverse(c2s), xf_inverse(s2r));
    std::memcpy(d.camera.raster_to_camera, r2c.m.m, 64);
    std::memcpy(d.camera.camera_to_world, camera_to_world.m.m, 64);

    // --- Sampler ---
    if (sampler_name == "halton") {
        d.sampler.type = B200PT_SAMPLER_HALTON;
        d.sampler.spp = sampler_p.one_int("pixelsamples", 16);
        d.sampler.sample_at_center = sampler_p.one_bool("samplepixelcenter", false) ? 1 : 0;
        d.sampler.dimensions = 4;
    } else if (sampler_name == "02sequence" || sampler_name == "lowdiscrepancy") {
        d.sampler.type = B200PT_SAMPLER_ZEROTWO;
        d.sampler.spp = sampler_p.one_int("pixelsamples", 16);
        d.sampler.dimensions = sampler_p.one_int("dimensions", 4);
    } else if (sampler_name == "sobol") {
        d.sampler.type = B200PT_SAMPLER_SOBOL;
        d.sampler.spp = sampler_p.one_int("pixelsamples", 16);
        d.sampler.dimensions = 4;
        // the generator matrices live next to the library: <dir of libb200pt.so>/data/sobol_matrices_32.bin
        // (or $B200PT_DATA_DIR); the reference links them in as SOBOL_MATRICES_32
        std::string dir;
        if (const char* e = std::getenv("B200PT_DATA_DIR")) dir = e;
        else {
            Dl_info info;
            if (dladdr((void*)&b200pt_load_pbrt, &info) && info.dli_fname) {
                std::string lib(info.dli_fname);
                size_t sl = lib.find_last_of('/');
                dir = (sl == std::string::npos ? std::string(".") : lib.substr(0, sl)) + "/data";
            }
        }
        std::ifstream f(dir + "/sobol_matrices_32.bin", std::ios::binary);
        L->sobol.resize(1024 * 52);
        if (!f || !f.read((char*)L->sobol.data(), (std::streamsize)(L->sobol.size() * 4)))
            throw Invalid("Sampler \"sobol\": cannot read " + dir + "/sobol_matrices_32.bin (set B200PT_DATA_DIR)");
        d.sobol_matrices_32 = L->sobol.data();
    } else throw Unsupported("Sampler \"" + sampler_name + "\" is outside this path (halton, 02sequence, sobol)");

    // --- Integrator (path.rs:287-326) ---
    if (integrator_name != "path" && integrator_name != "whitted" && integrator_name != "directlighting")
        throw Unsupported("Integrator \"" + integrator_name + "\" is outside this path (path, whitted, directlighting)");
    // whitted.rs:133-158 reads maxdepth and pixelbounds only; direct_lighting.rs:155-196 also "strategy" (unknown -> all)
    d.integrator.type = integrator_name == "whitted" ? B200PT_INTEGRATOR_WHITTED : integrator_name == "directlighting" ? B200PT_INTEGRATOR_DIRECT : B200PT_INTEGRATOR_PATH;
    d.integrator.direct_strategy = integrator_p.one_string("strategy", "all") == "one" ? B200PT_DIRECT_ONE : B200PT_DIRECT_ALL;
    d.integrator.max_depth = integrator_p.one_int("maxdepth", 5);
    d.integrator.rr_threshold = integrator_p.one_float("rrthreshold", 1.0f);
    std::string strat = integrator_p.one_string("lightsamplestrategy", integrator_name != "path" ? "uniform" : "spatial");
    if (strat == "uniform") d.integrator.light_strategy = B200PT_LIGHTS_UNIFORM;
    else if (strat == "power") d.integrator.light_strategy = B200PT_LIGHTS_POWER;
    else d.integrator.light_strategy = B200PT_LIGHTS_SPATIAL;  // "spatial", and any unknown name (path.rs:314-324 warns and uses spatial)
    int sb[4] = {(int)std::floor((float)d.film.crop[0] + 0.5f - rx), (int)std::floor((float)d.film.crop[1] + 0.5f - ry),
                 (int)std::ceil((float)d.film.crop[2] - 0.5f + rx), (int)std::ceil((float)d.film.crop[3] - 0.5f + ry)};
    if (const Param* pb = integrator_p.find("pixelbounds", "integer")) {
        if (pb->nums.size() != 4) throw Invalid("'pixelbounds' expects 4 values");
        // Bounds2i::new(Point2i(pb[0], pb[1]), Point2i(pb[2], pb[3])) intersected with the sample bounds (path.rs:296-311;
        // note the order x0 y0 x1 y1, unlike pbrt-v3's C++ reader)
        int x0 = pb->ints[0], y0 = pb->ints[1], x1 = pb->ints[2], y1 = pb->ints[3];
        sb[0] = sb[0] > x0 ? sb[0] : x0; sb[1] = sb[1] > y0 ? sb[1] : y0; sb[2] = sb[2] < x1 ? sb[2] : x1; sb[3] = sb[3] < y1 ? sb[3] : y1;
    }
    for (int i = 0; i < 4; ++i) d.integrator.pixel_bounds[i] = sb[i];

    // --- Accelerator (accelerators/src/bvh/mod.rs:339-360) ---
    if (accel_name != "bvh") throw Unsupported("Accelerator \"" + accel_name + "\" is outside this path (bvh)");
    const std::string split_method = accel_p.one_string("splitmethod", "sah");
    if (split_method != "sah" && split_method != "hlbvh") throw Unsupported("Accelerator \"bvh\": splitmethod \"sah\" and \"hlbvh\" are on this path, got \"" + split_method + "\"");
    L->max_node_prims = accel_p.one_int("maxnodeprims", 4) & 0xff;

    // --- geometry: top-level triangles, then each object's block; BVHs with the product's host SAH builder ---
    const int64_t n_top = (int64_t)L->flags.size();
    auto build = [&](const std::vector<float>& bounds, std::vector<b200pt_bvh_node>* nodes, std::vector<uint32_t>* ordered) {
        const int64_t n = (int64_t)bounds.size() / 6;
        nodes->resize((size_t)(n > 0 ? 2 * n : 1));
        ordered->resize((size_t)n);
        int64_t n_nodes = 0;
        // same bytes either way; the GPU builder (csrc/bvh_build.cu) is used once a device is bound and the input is large enough to pay for its launches
        auto fn = split_method == "hlbvh" ? b200pt_bvh_build_hlbvh : (b200pt_device_sm_count() > 0 && n >= 4096) ? b200pt_bvh_build_sah_gpu : b200pt_bvh_build_sah;
        if (n > 0 && fn(bounds.data(), n, L->max_node_prims, nodes->data(), &n_nodes, ordered->data()) != B200PT_OK)
            throw Invalid(std::string("BVH build failed: ") + b200pt_last_error());
        nodes->resize((size_t)n_nodes);
    };
    auto tri_bounds = [&](const std::vector<float>& v) {
        std::vector<float> b(v.size() / 9 * 6);
        if (!v.empty()) b200pt_triangle_bounds(v.data(), (int64_t)v.size() / 9, b.data());
        return b;
    };
    bool any_uv = false, any_n = false, any_s = false;
    for (uint32_t f : L->flags) { any_uv |= (f & B200PT_PRIM_HAS_UV) != 0; any_n |= (f & B200PT_PRIM_HAS_NORMALS) != 0; any_s |= (f & B200PT_PRIM_HAS_TANGENTS) != 0; }
    std::vector<float> top_bounds = tri_bounds(L->verts);
    // object ids referenced by instances keep their slot; unreferenced objects are still uploaded (harmless)
    for (ObjectDef& o : L->objects) {
        if (o.flags.empty()) throw Invalid("ObjectBegin/ObjectEnd without shapes");
        build(tri_bounds(o.verts), &o.nodes, &o.ordered);
        any_uv |= o.any_uv; any_n |= o.any_n; any_s |= o.any_s;
    }
    for (const b200pt_instance& in : L->instances) {
        // TransformedPrimitive::world_bound = primitive_to_world.transform_bounds(object root bounds), transform.rs:552-561
        const float* b = L->objects[(size_t)in.object].nodes[0].bounds;
        M4 m; std::memcpy(m.m, in.instance_to_world, 64);
        const float cs[8][3] = {{b[0], b[1], b[2]}, {b[3], b[1], b[2]}, {b[0], b[4], b[2]}, {b[0], b[1], b[5]}, {b[0], b[4], b[5]}, {b[3], b[4], b[2]}, {b[3], b[1], b[5]}, {b[3], b[4], b[5]}};
        float lo[3], hi[3];
        for (int k = 0; k < 8; ++k) {
            V3 q = xf_point(m, v3(cs[k][0], cs[k][1], cs[k][2]));
            const float c[3] = {q.x, q.y, q.z};
            for (int a = 0; a < 3; ++a) { lo[a] = k == 0 ? c[a] : (c[a] < lo[a] ? c[a] : lo[a]); hi[a] = k == 0 ? c[a] : (c[a] > hi[a] ? c[a] : hi[a]); }
        }
        for (int a = 0; a < 3; ++a) top_bounds.push_back(lo[a]);
        for (int a = 0; a < 3; ++a) top_bounds.push_back(hi[a]);
    }
    build(top_bounds, &L->nodes, &L->ordered);
    int64_t first = n_top;
    for (ObjectDef& o : L->objects) {
        b200pt_object od;
        od.nodes = o.nodes.data(); od.n_nodes = (int64_t)o.nodes.size(); od.ordered_prims = o.ordered.data();
        od.first_prim = first; od.n_prims = (int64_t)o.flags.size();
        L->object_descs.push_back(od);
        first += od.n_prims;
        L->verts.insert(L->verts.end(), o.verts.begin(), o.verts.end());
        L->uvs.insert(L->uvs.end(), o.uvs.begin(), o.uvs.end());
        L->normals.insert(L->normals.end(), o.normals.begin(), o.normals.end());
        L->tangents.insert(L->tangents.end(), o.tangents.begin(), o.tangents.end());
        L->flags.insert(L->flags.end(), o.flags.begin(), o.flags.end());
        L->material.insert(L->material.end(), o.material.begin(), o.material.end());
        L->light.insert(L->light.end(), o.flags.size(), -1);
        L->alpha_tex.insert(L->alpha_tex.end(), o.alpha_tex.begin(), o.alpha_tex.end());
    }
    {
        bool any_kd_tex = false;
        for (int32_t k : L->material_kd_tex) any_kd_tex |= k >= 0;
        if (any_kd_tex) {
            d.spectrum_textures = L->spectrum_textures.data(); d.n_spectrum_textures = (int32_t)L->spectrum_textures.size();
            d.material_kd_tex = L->material_kd_tex.data();
        }
    }
    if (!L->float_textures.empty()) {
        d.float_textures = L->float_textures.data(); d.n_float_textures = (int32_t)L->float_textures.size();
        d.prim_alpha_tex = L->alpha_tex.data();
        d.noise_perm = L->noise_perm.empty() ? nullptr : L->noise_perm.data();
    }
    d.nodes = L->nodes.data(); d.n_nodes = (int64_t)L->nodes.size(); d.ordered_prims = L->ordered.data();
    d.tri_verts = L->verts.data(); d.prim_flags = L->flags.data(); d.prim_material = L->material.data(); d.prim_light = L->light.data();
    d.n_prims = (int64_t)L->flags.size();
    d.tri_uvs = any_uv ? L->uvs.data() : nullptr;
    d.tri_normals = any_n ? L->normals.data() : nullptr;
    d.tri_tangents = any_s ? L->tangents.data() : nullptr;
    d.materials = L->materials.data(); d.n_materials = (int32_t)L->materials.size();
    d.lights = L->lights.data(); d.n_lights = (int32_t)L->lights.size();
    d.n_top_tris = n_top;
    if (!L->objects.empty()) {
        d.objects = L->object_descs.data(); d.n_objects = (int32_t)L->object_descs.size();
        d.instances = L->instances.data(); d.n_instances = (int32_t)L->instances.size();
    }
}

static ParamSet parse_params(Lexer& lx) {
    ParamSet ps;
    while (lx.peek().kind == Token::String) {
        Token decl = lx.next();
        std::istringstream ds(decl.text);
        Param p;
        ds >> p.type >> p.name;
        if (p.type.empty() || p.name.empty()) lx.fail("bad parameter declaration \"" + decl.text + "\"");
        auto take = [&](const Token& t) {
            if (t.kind == Token::Number) {
                p.nums.push_back((float)t.num);
                if (p.type == "integer") {
                    if (!(t.num >= -2147483648.0 && t.num <= 2147483647.0) || t.num != std::floor(t.num)) lx.fail("integer parameter \"" + decl.text + "\": value is not an i32");
                    p.ints.push_back((int32_t)t.num);
                }
            }
            else if (t.kind == Token::String) p.strs.push_back(t.text);
            else if (t.kind == Token::Ident && (t.text == "true" || t.text == "false")) p.strs.push_back(t.text);
            else lx.fail("bad value for parameter \"" + decl.text + "\"");
        };
        Token t = lx.next();
        if (t.kind == Token::LBracket) {
            for (;;) {
                Token v = lx.next();
                if (v.kind == Token::RBracket) break;
                if (v.kind == Token::End) lx.fail("unterminated parameter list");
                take(v);
            }
        } else take(t);
        ps.ps.push_back(p);
    }
    return ps;
}

static void parse_file(const std::string& path, Builder& B);

static void parse_stream(Lexer& lx, Builder& B) {
    auto num = [&]() -> float {
        Token t = lx.next();
        if (t.kind != Token::Number) lx.fail("number expected");
        return (float)t.num;
    };
    auto str = [&]() -> std::string {
        Token t = lx.next();
        if (t.kind != Token::String) lx.fail("quoted string expected");
        return t.text;
    };
    auto nums = [&](int n, float* out) {
        bool br = lx.peek().kind == Token::LBracket;
        if (br) lx.next();
        for (int i = 0; i < n; ++i) out[i] = num();
        if (br && lx.next().kind != Token::RBracket) lx.fail("']' expected");
    };
    for (;;) {
        Token t = lx.next();
        if (t.kind == Token::End) return;
        if (t.kind != Token::Ident) lx.fail("directive expected, got '" + t.text + "'");
        const std::string& d = t.text;
        GState& gs = B.gs;
        if (d == "Identity") gs.ctm = xf_identity();
        else if (d == "Translate") { float v[3]; nums(3, v); gs.ctm = xf_mul(gs.ctm, xf_translate(v[0], v[1], v[2])); }
        else if (d == "Scale") { float v[3]; nums(3, v); gs.ctm = xf_mul(gs.ctm, xf_scale(v[0], v[1], v[2])); }
        else if (d == "Rotate") { float v[4]; nums(4, v); gs.ctm = xf_mul(gs.ctm, xf_rotate(v[0], v3(v[1], v[2], v[3]))); }
        else if (d == "LookAt") { float v[9]; nums(9, v); gs.ctm = xf_mul(gs.ctm, xf_look_at(v3(v[0], v[1], v[2]), v3(v[3], v[4], v[5]), v3(v[6], v[7], v[8]))); }
        else if (d == "Transform" || d == "ConcatTransform") {
            float v[16]; nums(16, v);
            M4 m;  // the file stores the matrix column-major: pbrt transposes it (api/src/lib.rs pbrt_transform)
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.m[i][j] = v[4 * j + i];
            gs.ctm = d == "Transform" ? xf_from(m) : xf_mul(gs.ctm, xf_from(m));
        }
        else if (d == "CoordinateSystem") B.named_cs[str()] = gs.ctm;
        else if (d == "CoordSysTransform") { std::string n = str(); if (B.named_cs.count(n)) gs.ctm = B.named_cs[n]; }
        else if (d == "ReverseOrientation") gs.reverse = !gs.reverse;
        else if (d == "AttributeBegin") { B.stack.push_back(gs); }
        else if (d == "AttributeEnd") { if (B.stack.empty()) lx.fail("unmatched AttributeEnd"); gs = B.stack.back(); B.stack.pop_back(); }
        else if (d == "TransformBegin") B.xf_stack.push_back(gs.ctm);
        else if (d == "TransformEnd") { if (B.xf_stack.empty()) lx.fail("unmatched TransformEnd"); gs.ctm = B.xf_stack.back(); B.xf_stack.pop_back(); }
        else if (d == "Camera") { B.camera_name = str(); B.camera_p = parse_params(lx); B.camera_to_world = xf_inverse(gs.ctm); B.named_cs["camera"] = B.camera_to_world; }
        else if (d == "Film") { B.film_name = str(); B.film_p = parse_params(lx); }
        else if (d == "Sampler") { B.sampler_name = str(); B.sampler_p = parse_params(lx); }
        else if (d == "PixelFilter") { B.filter_name = str(); B.filter_p = parse_params(lx); }
        else if (d == "Integrator") { B.integrator_name = str(); B.integrator_p = parse_params(lx); }
        else if (d == "Accelerator") { B.accel_name = str(); B.accel_p = parse_params(lx); }
        else if (d == "WorldBegin") { B.in_world = true; gs.ctm = xf_identity(); B.named_cs["world"] = gs.ctm; }
        else if (d == "WorldEnd") { B.in_world = false; }
        else if (d == "Material") { std::string n = str(); ParamSet p = parse_params(lx); gs.material = B.make_material(n, p); }
        else if (d == "MakeNamedMaterial") {
            std::string n = str(); ParamSet p = parse_params(lx);
            B.named_materials[n] = B.make_material(p.one_string("type", ""), p);
        }
        else if (d == "NamedMaterial") { std::string n = str(); if (!B.named_materials.count(n)) lx.fail("NamedMaterial \"" + n + "\" is not defined"); gs.material = B.named_materials[n]; }
        else if (d == "LightSource") { std::string n = str(); ParamSet p = parse_params(lx); B.light_source(n, p); }
        else if (d == "AreaLightSource") {
            std::string n = str(); ParamSet p = parse_params(lx);
            if (n != "diffuse" && n != "area") throw Unsupported("AreaLightSource \"" + n + "\" is outside this path (diffuse)");
            const float one[3] = {1, 1, 1};
            float Lv[3], sc[3];
            p.one_rgb("L", one, Lv); p.one_rgb("scale", one, sc);
            for (int c = 0; c < 3; ++c) gs.area_L[c] = Lv[c] * sc[c];
            gs.area_two_sided = p.one_bool("twosided", false);
            gs.has_area = true;
        }
        else if (d == "Shape") { std::string n = str(); ParamSet p = parse_params(lx); B.shape(n, p); }
        else if (d == "ObjectBegin") {
            std::string n = str();
            if (B.cur_object >= 0) lx.fail("ObjectBegin called inside of instance definition");
            B.stack.push_back(gs);  // pbrt_attribute_begin
            B.L->objects.emplace_back();
            B.cur_object = (int)B.L->objects.size() - 1;
            B.object_ids[n] = B.cur_object;
        }
        else if (d == "ObjectEnd") {
            if (B.cur_object < 0) lx.fail("ObjectEnd called outside of instance definition");
            B.cur_object = -1;
            gs = B.stack.back(); B.stack.pop_back();
        }
        else if (d == "ObjectInstance") {
            std::string n = str();
            if (B.cur_object >= 0) lx.fail("ObjectInstance can't be called inside of instance definition");
            if (!B.object_ids.count(n)) lx.fail("Unable to find object instance named '" + n + "'");
            b200pt_instance in;
            in.object = B.object_ids[n];
            std::memcpy(in.instance_to_world, gs.ctm.m.m, 64);
            std::memcpy(in.world_to_instance, gs.ctm.inv.m, 64);
            B.L->instances.push_back(in);
        }
        else if (d == "Include") { parse_file(B.resolve(str()), B); }
        else if (d == "Texture") { std::string n = str(); std::string ty = str(); std::string cls = str(); ParamSet p = parse_params(lx); B.texture(n, ty, cls, p); }
        else if (d == "MakeNamedMedium" || d == "MediumInterface" || d == "ActiveTransform" || d == "TransformTimes")
            throw Unsupported("directive " + d + " is outside this path");
        else lx.fail("unknown directive '" + d + "'");
    }
}

static void parse_file(const std::string& path, Builder& B) {
    struct Depth { int& d; explicit Depth(int& x) : d(x) { ++d; } ~Depth() { --d; } } depth(B.include_depth);
    if (B.include_depth > 32) throw Invalid("Include nested deeper than 32 files (a file that includes itself?): '" + path + "'");
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Invalid("cannot open scene file '" + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    Lexer lx;
    lx.src = ss.str();
    lx.file = path;
    parse_stream(lx, B);
}

}  // namespace b2load

struct b200pt_loaded_scene {
    b2load::Loaded L;
};

extern "C" {

int b200pt_load_pbrt(const char* path, b200pt_loaded_scene** out) {
    if (!path || !out) { b200pt_set_error("b200pt_load_pbrt: null argument"); return B200PT_ERR_INVALID; }
    *out = nullptr;
    std::unique_ptr<b200pt_loaded_scene> s(new b200pt_loaded_scene());
    try {
        b2load::Builder B;
        B.L = &s->L;
        std::string p(path);
        size_t slash = p.find_last_of('/');
        B.dir = slash == std::string::npos ? "" : p.substr(0, slash);
        b2load::parse_file(p, B);
        B.finish();
    } catch (const b2load::Unsupported& e) {
        b200pt_set_error((std::string("b200pt_load_pbrt: ") + e.what()).c_str());
        return B200PT_ERR_UNSUPPORTED;
    } catch (const std::exception& e) {
        b200pt_set_error((std::string("b200pt_load_pbrt: ") + e.what()).c_str());
        return B200PT_ERR_INVALID;
    }
    *out = s.release();
    return B200PT_OK;
}
const b200pt_scene_desc* b200pt_loaded_scene_desc(const b200pt_loaded_scene* s) { return s ? &s->L.desc : nullptr; }
const char* b200pt_loaded_scene_output(const b200pt_loaded_scene* s) { return s ? s->L.output.c_str() : ""; }
void b200pt_loaded_scene_free(b200pt_loaded_scene* s) { delete s; }

// 8-bit PNG as core/src/image_io.rs writes it (write_8_bit, :291-390): every channel clamp(255 * gamma_correct(v) + 0.5, 0, 255)
// as u8, RGB, 8 bits, no interlace.  The deflate stream uses stored blocks (no compressor in this image; any PNG reader
// accepts them), CRC-32 / Adler-32 computed here.
namespace b2load {
static uint32_t crc32_update(uint32_t c, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t v = i; for (int k = 0; k < 8; ++k) v = (v & 1u) ? 0xedb88320u ^ (v >> 1) : v >> 1; table[i] = v; }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c;
}
static void png_chunk(std::ofstream& f, const char type[4], const std::vector<uint8_t>& data) {
    auto be32 = [&](uint32_t v) { const uint8_t b[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v}; f.write((const char*)b, 4); };
    be32((uint32_t)data.size());
    f.write(type, 4);
    if (!data.empty()) f.write((const char*)data.data(), (std::streamsize)data.size());
    uint32_t c = crc32_update(0xffffffffu, (const uint8_t*)type, 4);
    c = crc32_update(c, data.data(), data.size()) ^ 0xffffffffu;
    be32(c);
}
static float gamma_correct(float v) { return v <= 0.0031308f ? 12.92f * v : 1.055f * std::pow(v, 1.0f / 2.4f) - 0.055f; }  // pbrt/common.rs:140-146
static void write_png(const std::string& path, const float* rgb, int w, int h) {
    std::vector<uint8_t> raw((size_t)h * (1 + 3 * (size_t)w));
    for (int y = 0; y < h; ++y) {
        uint8_t* row = &raw[(size_t)y * (1 + 3 * (size_t)w)];
        row[0] = 0;  // filter type None
        for (int x = 0; x < 3 * w; ++x) {
            float v = 255.0f * gamma_correct(rgb[(size_t)y * 3 * w + x]) + 0.5f;
            v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);  // clamp(); NaN -> 0 like Rust's `as u8`
            row[1 + x] = v == v ? (uint8_t)v : 0;
        }
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
        for (size_t i = 0; i < n; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        z.insert(z.end(), raw.begin() + (std::ptrdiff_t)pos, raw.begin() + (std::ptrdiff_t)(pos + n));
        pos += n;
        if (last) break;
    }
    const uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Invalid("cannot write image '" + path + "'");
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write((const char*)sig, 8);
    std::vector<uint8_t> ihdr = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    png_chunk(f, "IHDR", ihdr);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
    if (!f) throw Invalid("error writing image '" + path + "'");
}
}  // namespace b2load

int b200pt_write_png(const char* path, const float* rgb, int32_t width, int32_t height) {
    if (!path || !rgb || width <= 0 || height <= 0) { b200pt_set_error("b200pt_write_png: invalid argument"); return B200PT_ERR_INVALID; }
    try { b2load::write_png(path, rgb, width, height); }
    catch (const std::exception& e) { b200pt_set_error(e.what()); return B200PT_ERR_INVALID; }
    return B200PT_OK;
}
// write_image (core/src/image_io.rs:227-289): dispatch on the extension; .png and .pfm are on this path.
int b200pt_write_image(const char* path, const float* rgb, int32_t width, int32_t height) {
    if (!path) { b200pt_set_error("b200pt_write_image: null path"); return B200PT_ERR_INVALID; }
    const std::string p(path);
    const size_t dot = p.find_last_of('.');
    const std::string ext = dot == std::string::npos ? std::string() : p.substr(dot);
    if (ext == ".png") return b200pt_write_png(path, rgb, width, height);
    if (ext == ".pfm") return b200pt_write_pfm(path, rgb, width, height);
    b200pt_set_error("b200pt_write_image: only .png and .pfm are written on this path (the reference also writes .exr / .tga)");
    return B200PT_ERR_UNSUPPORTED;
}

int b200pt_write_pfm(const char* path, const float* rgb, int32_t width, int32_t height) {
    if (!path || !rgb || width <= 0 || height <= 0) { b200pt_set_error("b200pt_write_pfm: invalid argument"); return B200PT_ERR_INVALID; }
    try { b2load::write_pfm(path, rgb, width, height); }
    catch (const std::exception& e) { b200pt_set_error(e.what()); return B200PT_ERR_INVALID; }
    return B200PT_OK;
}
int b200pt_read_pfm(const char* path, float* rgb_out, int32_t size2[2]) {
    if (!path || !size2) { b200pt_set_error("b200pt_read_pfm: invalid argument"); return B200PT_ERR_INVALID; }
    try {
        std::vector<float> px;
        int w = 0, h = 0;
        b2load::read_pfm(path, &px, &w, &h);
        size2[0] = w; size2[1] = h;
        if (rgb_out) std::memcpy(rgb_out, px.data(), px.size() * 4);
    } catch (const std::exception& e) { b200pt_set_error(e.what()); return B200PT_ERR_INVALID; }
    return B200PT_OK;
}

}  // extern "C"
