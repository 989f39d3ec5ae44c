// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY: pinned by execution for Whitted-class renders, see oracle_math.h.
//
// extern "C" surface of the CPU oracle, loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs.  Nothing in the product path links or loads this library.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "oracle_bvh.h"
#include "oracle_hlbvh.h"
#include "oracle_math.h"
#include "oracle_rng.h"
#include "oracle_render.h"

using namespace orc;

namespace {
template <class F> void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 1024) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    size_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min(n, chunk * t), e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { f(b, e, t); });
    }
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

struct orc_ray { float o[3]; float tmax; float d[3]; float time; };      // 32 B
struct orc_hit { float t; uint32_t prim; float b0, b1; };                // 16 B
struct orc_hit_diag { float b2, det, min_e_abs, second_t; };             // exemption-band diagnostics

// ---- known-answer helpers --------------------------------------------------
void orc_pcg32_stream(uint64_t seq, uint64_t seed, int use_default, uint32_t* out, int n) {
    RNG r;
    if (!use_default) r.set_sequence(seq, seed);
    for (int i = 0; i < n; ++i) out[i] = r.uniform_u32();
}
void orc_pcg32_floats(uint64_t seq, float* out, int n) {
    RNG r(seq);
    for (int i = 0; i < n; ++i) out[i] = r.uniform_float();
}
uint32_t orc_pcg32_bounded(uint64_t seq, uint32_t lo, uint32_t hi, int skip) {
    RNG r(seq);
    uint32_t v = 0;
    for (int i = 0; i <= skip; ++i) v = r.bounded_u32(lo, hi);
    return v;
}
float orc_radical_inverse(int base_index, uint64_t a) { return radical_inverse(base_index, a); }
float orc_scrambled_radical_inverse(int base_index, uint64_t a) {
    return scrambled_radical_inverse(base_index, a, &halton_permutations()[prime_tables().sums[base_index]]);
}
int orc_prime(int i) { return prime_tables().primes[i]; }
int orc_prime_sum(int i) { return prime_tables().sums[i]; }
int orc_prime_total() { return prime_tables().total; }
// Copies the permutation table (u16 x total) into out.
void orc_halton_permutations(uint16_t* out) {
    const std::vector<uint16_t>& p = halton_permutations();
    std::copy(p.begin(), p.end(), out);
}
// Halton sampler: fills out[n_samples][n_dims] for one pixel.
void orc_halton_pixel(int spp, int res_x, int res_y, int px, int py, int n_dims, float* out) {
    HaltonSampler s(spp, res_x, res_y, false);
    s.start_pixel(px, py);
    int k = 0;
    do {
        for (int d = 0; d < n_dims; ++d) out[k * n_dims + d] = s.get_1d();
        ++k;
    } while (s.start_next_sample());
}
// Sobol sampler: the caller supplies the 1024 x 52 generator matrices (SOBOL_MATRICES_32).
void orc_set_sobol_matrices(const uint32_t* m, int64_t n) { sobol_matrices_32().assign(m, m + n); }
// VD_C_SOBOL_MATRICES[m - 1] / VD_C_SOBOL_MATRICES_INV[m - 1] as derived from dimensions 0 and 1 (52 + 52 u64)
void orc_sobol_interval_tables(int m, uint64_t* vdc, uint64_t* vdc_inv) {
    SobolIntervalTables t = sobol_interval_tables(m);
    for (int c = 0; c < kSobolMatrixSize; ++c) { vdc[c] = t.vdc[c]; vdc_inv[c] = t.vdc_inv[c]; }
}
// fills out[n_samples][n_dims] for one pixel; sample_bounds = x0 y0 x1 y1; returns the sample count actually taken
int orc_sobol_pixel(int spp, const int* sample_bounds, int px, int py, int n_dims, float* out, uint64_t* index_out) {
    SobolSampler s(spp, sample_bounds);
    s.start_pixel(px, py);
    int k = 0;
    do {
        if (index_out) index_out[k] = s.interval_sample_index;
        for (int d = 0; d < n_dims; ++d) out[k * n_dims + d] = s.get_1d();
        ++k;
    } while (s.start_next_sample());
    return k;
}
uint64_t orc_halton_index(int spp, int res_x, int res_y, int px, int py, int sample) {
    HaltonSampler s(spp, res_x, res_y, false);
    s.start_pixel(px, py);
    return s.get_index_for_sample((uint64_t)sample);
}
float orc_gamma(int n) { return gamma(n); }
float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
void orc_coordinate_system(const float* v1, float* v2, float* v3) {
    V3 a, b;
    coordinate_system(V3(v1[0], v1[1], v1[2]), &a, &b);
    v2[0] = a.x; v2[1] = a.y; v2[2] = a.z;
    v3[0] = b.x; v3[1] = b.y; v3[2] = b.z;
}
void orc_cross(const float* a, const float* b, float* o) {
    V3 r = cross(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void orc_normalize(const float* a, float* o) {
    V3 r = normalize(V3(a[0], a[1], a[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void orc_offset_ray_origin(const float* p, const float* perr, const float* n, const float* w, float* o) {
    V3 r = offset_ray_origin(V3(p[0], p[1], p[2]), V3(perr[0], perr[1], perr[2]), V3(n[0], n[1], n[2]), V3(w[0], w[1], w[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void orc_matrix_inverse(const float* m, float* o) {
    M4 a;
    std::memcpy(a.m, m, 64);
    M4 r = minverse(a);
    std::memcpy(o, r.m, 64);
}
// Camera matrices exactly as the reference's api builds them:
// LookAt (transform.rs:191) -> camera_to_world = inverse; PerspectiveCamera::new
// (perspective_camera.rs:35-75) + ProjectiveCameraData::new (camera.rs:276-306).
void orc_camera_matrices(const float* eye, const float* look, const float* up, float fov, int xres, int yres,
                         const float* screen_window /*4 or null*/, float* camera_to_world, float* raster_to_camera) {
    Transform w2c = tlook_at(V3(eye[0], eye[1], eye[2]), V3(look[0], look[1], look[2]), V3(up[0], up[1], up[2]));
    std::memcpy(camera_to_world, w2c.m_inv.m, 64);
    float frame = (float)xres / (float)yres;
    float sw[4];
    if (screen_window) { std::memcpy(sw, screen_window, 16); }
    else if (frame > 1.0f) { sw[0] = -frame; sw[1] = frame; sw[2] = -1.0f; sw[3] = 1.0f; }
    else { sw[0] = -1.0f; sw[1] = 1.0f; sw[2] = -1.0f / frame; sw[3] = 1.0f / frame; }
    Transform c2s = tperspective(fov, 1e-2f, 1000.0f);
    Transform s2r = tmul(tmul(tscale((float)xres, (float)yres, 1.0f), tscale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f)),
                         ttranslate(V3(-sw[0], -sw[3], 0.0f)));
    Transform r2s = tinverse(s2r);
    Transform r2c = tmul(tinverse(c2s), r2s);
    std::memcpy(raster_to_camera, r2c.m.m, 64);
}

// ---- BVH ------------------------------------------------------------------
// Returns number of nodes; nodes_out must hold 2n-1 nodes, ordered_out n u32.
int64_t orc_bvh_build_sah(const float* prim_bounds, int64_t n, int max_prims_in_node, void* nodes_out, uint32_t* ordered_out) {
    std::vector<LinearBVHNode> nodes;
    std::vector<uint32_t> ordered;
    bvh_build_sah(prim_bounds, (size_t)n, max_prims_in_node, nodes, ordered);
    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(LinearBVHNode));
    std::copy(ordered.begin(), ordered.end(), ordered_out);
    return (int64_t)nodes.size();
}
// BVHAccel::new(.., SplitMethod::HLBVH).  Returns the node count, or -1 where the reference panics.
// morton_sorted_out (optional): the n Morton codes after the radix sort.
int64_t orc_bvh_build_hlbvh(const float* prim_bounds, int64_t n, int max_prims_in_node, void* nodes_out, uint32_t* ordered_out,
                            uint32_t* morton_sorted_out) {
    std::vector<LinearBVHNode> nodes;
    std::vector<uint32_t> ordered, codes;
    if (!bvh_build_hlbvh(prim_bounds, (size_t)n, max_prims_in_node, nodes, ordered, morton_sorted_out ? &codes : nullptr)) return -1;
    std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(LinearBVHNode));
    std::copy(ordered.begin(), ordered.end(), ordered_out);
    if (morton_sorted_out) std::copy(codes.begin(), codes.end(), morton_sorted_out);
    return (int64_t)nodes.size();
}
void orc_triangle_bounds(const float* verts, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        const float* v = verts + 9 * i;
        triangle_world_bound(V3(v[0], v[1], v[2]), V3(v[3], v[4], v[5]), V3(v[6], v[7], v[8]), out + 6 * i);
    }
}

void* orc_accel_create(const void* nodes, int64_t n_nodes, const uint32_t* ordered, const float* verts, const uint32_t* flags,
                       int64_t n_prims) {
    Accel* a = new Accel();
    a->nodes.resize((size_t)n_nodes);
    std::memcpy(a->nodes.data(), nodes, (size_t)n_nodes * sizeof(LinearBVHNode));
    a->ordered.assign(ordered, ordered + n_prims);
    a->verts.assign(verts, verts + 9 * n_prims);
    if (flags) a->flags.assign(flags, flags + n_prims);
    return a;
}
// Optional "uv"/"st" of the meshes (6 floats per primitive; flags carry PRIM_HAS_UV): they enter the traversal through the
// degenerate-hit rejection of triangle.rs:551-572.
void orc_accel_set_uvs(void* accel, const float* uvs, int64_t n_prims) {
    Accel* a = (Accel*)accel;
    a->uvs.assign(uvs, uvs + 6 * n_prims);
}
// Alpha-mask textures of a stand-alone accelerator (same arrays as b200pt_accel_set_alpha_textures).
void orc_accel_set_alpha(void* accel, const b200pt_float_texture* tex, int n_tex, const int32_t* prim_alpha_tex, const uint8_t* noise_perm, int64_t n_prims) {
    Accel* a = (Accel*)accel;
    a->alpha_tex.assign(prim_alpha_tex, prim_alpha_tex + 2 * n_prims);
    a->textures.clear();
    for (int k = 0; k < n_tex; ++k) a->textures.push_back(float_texture_from(tex[k]));
    a->noise_perm.clear();
    if (noise_perm) a->noise_perm.assign(noise_perm, noise_perm + 256);
}
// Texture<Spectrum>::evaluate for uv and the four screen-space uv derivatives der = {dudx, dvdx, dudy, dvdy}
void orc_spectrum_texture_evaluate(const b200pt_spectrum_texture* t, float u, float v, const float* der, float* out3) {
    SpectrumTexture T;
    T.type = t->type; T.su = t->su; T.sv = t->sv; T.du = t->du; T.dv = t->dv; T.closedform = t->aa_closedform != 0;
    for (int c = 0; c < 3; ++c) { T.tex1[c] = t->tex1[c]; T.tex2[c] = t->tex2[c]; }
    UVDerivs d;
    d.dudx = der[0]; d.dvdx = der[1]; d.dudy = der[2]; d.dvdy = der[3];
    spectrum_texture_evaluate(T, u, v, d, out3);
}
float orc_float_texture_evaluate(const b200pt_float_texture* tex, const uint8_t* noise_perm, float u, float v) {
    return float_texture_evaluate(float_texture_from(*tex), noise_perm, u, v);
}
float orc_noise_3d(const uint8_t* noise_perm, float x, float y, float z) { return noise_3d(noise_perm, x, y, z); }
void orc_accel_destroy(void* a) { delete (Accel*)a; }
// Hit geometry of a triangle at barycentrics b (triangle.rs:547-725).  attrs: uv6 / n9 / s9 may be null.
// out: p(3) p_error(3) n(3) dpdu(3) dpdv(3) shading_n(3) shading_dpdu(3) = 21 floats.  Returns 0 for a degenerate hit.
int orc_triangle_geometry(const float* verts9, const float* b3, const float* uv6, const float* n9, const float* s9, int flip, int reverse, float* out21) {
    V3 p0(verts9[0], verts9[1], verts9[2]), p1(verts9[3], verts9[4], verts9[5]), p2(verts9[6], verts9[7], verts9[8]);
    TriAttr at;
    at.uv = uv6; at.n = n9; at.s = s9; at.flip = flip != 0; at.reverse = reverse != 0;
    TriGeom g;
    if (!triangle_geometry(p0, p1, p2, b3[0], b3[1], b3[2], at, &g)) return 0;
    const V3 v[7] = {g.p, g.p_error, g.n, g.dpdu, g.dpdv, g.shading_n, g.shading_dpdu};
    for (int i = 0; i < 7; ++i) { out21[3 * i] = v[i].x; out21[3 * i + 1] = v[i].y; out21[3 * i + 2] = v[i].z; }
    return 1;
}

// Closest hit over a batch.  counters (optional): 2 x u32 per ray = nodes
// tested, triangles tested (SURVEY §8d algorithmic-bytes accounting).
void orc_intersect_batch(const void* accel, const orc_ray* rays, int64_t n, orc_hit* hits, orc_hit_diag* diag, uint32_t* counters,
                         int nthreads) {
    const Accel& a = *(const Accel*)accel;
    parallel_for((size_t)n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            Ray r(V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmax, rays[i].time);
            HitRecord h;
            TraversalCounters c;
            bool ok = bvh_intersect(a, r, &h, counters ? &c : nullptr);
            hits[i].t = ok ? h.t : kInfinity;
            hits[i].prim = ok ? h.prim : 0xffffffffu;
            hits[i].b0 = ok ? h.b0 : 0.0f;
            hits[i].b1 = ok ? h.b1 : 0.0f;
            if (diag) {
                diag[i].b2 = ok ? h.b2 : 0.0f; diag[i].det = ok ? h.det : 0.0f;
                diag[i].min_e_abs = ok ? h.min_e_abs : 0.0f; diag[i].second_t = ok ? h.second_t : kInfinity;
            }
            if (counters) { counters[2 * i] = (uint32_t)c.nodes; counters[2 * i + 1] = (uint32_t)c.tris; }
        }
    });
}
void orc_occluded_batch(const void* accel, const orc_ray* rays, int64_t n, uint8_t* out, uint32_t* counters, int nthreads) {
    const Accel& a = *(const Accel*)accel;
    parallel_for((size_t)n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            Ray r(V3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), V3(rays[i].d[0], rays[i].d[1], rays[i].d[2]), rays[i].tmax, rays[i].time);
            TraversalCounters c;
            out[i] = bvh_intersect_p(a, r, counters ? &c : nullptr) ? 1 : 0;
            if (counters) { counters[2 * i] = (uint32_t)c.nodes; counters[2 * i + 1] = (uint32_t)c.tris; }
        }
    });
}
// Single triangle test (hand-checkable KATs).  Returns 1 on hit; out = t,b0,b1,b2.
int orc_triangle_intersect(const float* ray8, const float* verts9, float* out4) {
    Ray r(V3(ray8[0], ray8[1], ray8[2]), V3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7]);
    TriHit th;
    V3 p0(verts9[0], verts9[1], verts9[2]), p1(verts9[3], verts9[4], verts9[5]), p2(verts9[6], verts9[7], verts9[8]);
    if (!triangle_test(r, p0, p1, p2, &th)) return 0;
    TriGeom g;
    if (!triangle_geometry(p0, p1, p2, th.b0, th.b1, th.b2, TriAttr(), &g)) return 0;
    out4[0] = th.t; out4[1] = th.b0; out4[2] = th.b1; out4[3] = th.b2;
    return 1;
}
int orc_bounds_intersect(const float* bounds6, const float* ray8) {
    Ray r(V3(ray8[0], ray8[1], ray8[2]), V3(ray8[4], ray8[5], ray8[6]), ray8[3], ray8[7]);
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int neg[3] = {inv_dir.x < 0.0f ? 1 : 0, inv_dir.y < 0.0f ? 1 : 0, inv_dir.z < 0.0f ? 1 : 0};
    return bounds_intersect_p_inv(bounds6, r, inv_dir, neg) ? 1 : 0;
}

// Environment map preparation (oracle_envmap.h + compute_scalar_image): level 0 texels, importance image, power lookup.
// Call with null outputs to get size4 = {w0, h0, 2*w0, 2*h0}.
void orc_envmap_prepare(const float* rgb, int w, int h, const float* L, int32_t* size4, float* level0_out, float* importance_out, float* power_out) {
    std::vector<RGB> tex;
    int mw = 1, mh = 1;
    RGB l(L[0], L[1], L[2]);
    if (rgb && w > 0 && h > 0) {
        mw = w; mh = h;
        for (size_t k = 0; k < (size_t)w * h; ++k) tex.push_back(RGB(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]) * l);
    } else tex.push_back(l);
    MipMap m;
    m.build(mw, mh, tex);
    const int nu = 2 * m.width(), nv = 2 * m.height();
    size4[0] = m.width(); size4[1] = m.height(); size4[2] = nu; size4[3] = nv;
    if (level0_out)
        for (size_t k = 0; k < m.pyramid[0].px.size(); ++k)
            for (int c = 0; c < 3; ++c) level0_out[3 * k + c] = m.pyramid[0].px[k].c[c];
    if (importance_out) {
        const Float fwidth = 0.5f / (Float)(nu < nv ? nu : nv);
        for (int v = 0; v < nv; ++v) {
            Float vp = ((Float)v + 0.5f) / (Float)nv, sin_t = std::sin(kPi * ((Float)v + 0.5f) / (Float)nv);
            for (int u = 0; u < nu; ++u) importance_out[(size_t)v * nu + u] = lum_y(m.lookup_triangle(P2(((Float)u + 0.5f) / (Float)nu, vp), fwidth)) * sin_t;
        }
    }
    if (power_out) { RGB p = m.lookup_triangle(P2(0.5f, 0.5f), 0.5f); power_out[0] = p.c[0]; power_out[1] = p.c[1]; power_out[2] = p.c[2]; }
}
// One MIPMap::lookup_triangle on a freshly built map (KATs).
void orc_envmap_lookup(const float* rgb, int w, int h, const float* st2, float width, float* out3) {
    std::vector<RGB> tex;
    for (size_t k = 0; k < (size_t)w * h; ++k) tex.push_back(RGB(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]));
    MipMap m;
    m.build(w, h, tex);
    RGB v = m.lookup_triangle(P2(st2[0], st2[1]), width);
    out3[0] = v.c[0]; out3[1] = v.c[1]; out3[2] = v.c[2];
}

// ---- rendering (oracle_render.h) ------------------------------------------
void* orc_scene_create(const b200pt_scene_desc* d) { return scene_create(d); }
void orc_scene_destroy(void* s) { scene_destroy((RenderScene*)s); }
// Renders with `nthreads` workers pulling 16x16 tiles (sampler_integrator.rs:243-304).
// rgb_out: 3 floats per cropped pixel, row-major top-to-bottom (write_image order).
// stats_out (optional, 4 x u64): camera rays, closest-hit rays, shadow rays, reserved.
double orc_render(void* s, float* rgb_out, uint64_t* stats_out, int nthreads) { return render((RenderScene*)s, rgb_out, stats_out, nthreads); }
// Li for explicit (pixel, sample) pairs; out: 3 floats per entry.
void orc_li_batch(void* s, const int32_t* pixel_sample /*x,y,sample per entry*/, int64_t n, float* out, int nthreads) {
    li_batch((RenderScene*)s, pixel_sample, n, out, nthreads);
}
// Camera rays for explicit (pixel, sample) pairs.
void orc_camera_rays(void* s, const int32_t* pixel_sample, int64_t n, orc_ray* out) {
    camera_rays((RenderScene*)s, pixel_sample, n, (float*)out);
}

}  // extern "C"
