// CPU oracle (test infrastructure, never linked into the product): the float textures a triangle mesh can use as
// "alpha" / "shadowalpha", restated from
//   textures/src/{constant,checkerboard_2d,dots,imagemap}.rs, textures/src/lib.rs:43-66 (get_texture_mapping),
//   core/src/texture/mapping/uv_2d.rs, core/src/texture/common.rs:38-125 (Perlin noise), core/src/mipmap/mod.rs:212-311, 580-608.
// The interaction Triangle::intersect / intersect_p builds for the alpha test has zero derivatives
// (shapes/src/triangle.rs:587-607, 856-884), which is what makes these closed forms complete.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "oracle_math.h"

namespace orc {

struct FloatTexture {
    int type = 0;  // B200PT_TEX_*: 0 constant, 1 checkerboard, 2 dots, 3 imagemap
    Float su = 1.0f, sv = 1.0f, du = 0.0f, dv = 0.0f;
    Float a = 0.0f, b = 0.0f;  // constant: a; checkerboard: tex1, tex2; dots: outside_dot, inside_dot (as stored, see dots.rs:31,86)
    int wrap = 0, width = 0, height = 0;
    std::vector<Float> texels;  // level 0, texels[t * width + s]
};

// core/src/texture/common.rs:104-119.  `perm` holds NOISE_PERM[0..256); the reference table repeats it once.
inline Float noise_grad(const uint8_t* perm, long x, long y, long z, Float dx, Float dy, Float dz) {
    auto P = [&](long i) { return (long)perm[i & 255]; };
    long h = P(P(P(x) + y) + z) & 15;
    Float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
    Float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
    return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
inline Float noise_weight(Float t) {  // common.rs:121-125
    Float t3 = t * t * t;
    Float t4 = t3 * t;
    return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3;
}
inline Float lerp_f(Float t, Float a, Float b) { return (1.0f - t) * a + t * b; }
inline long floor_as_isize(Float v) {  // Rust `as isize`: saturating, NaN -> 0
    Float f = std::floor(v);
    if (std::isnan(f)) return 0;
    if (f >= 9.2233720e18f) return (long)0x7fffffffffffffffLL;
    if (f <= -9.2233720e18f) return (long)0x8000000000000000LL;
    return (long)f;
}
inline Float noise_3d(const uint8_t* perm, Float x, Float y, Float z) {  // common.rs:38-73
    long ix = floor_as_isize(x), iy = floor_as_isize(y), iz = floor_as_isize(z);
    Float dx = x - (Float)ix, dy = y - (Float)iy, dz = z - (Float)iz;
    ix &= 255; iy &= 255; iz &= 255;
    Float w000 = noise_grad(perm, ix, iy, iz, dx, dy, dz);
    Float w100 = noise_grad(perm, ix + 1, iy, iz, dx - 1.0f, dy, dz);
    Float w010 = noise_grad(perm, ix, iy + 1, iz, dx, dy - 1.0f, dz);
    Float w110 = noise_grad(perm, ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
    Float w001 = noise_grad(perm, ix, iy, iz + 1, dx, dy, dz - 1.0f);
    Float w101 = noise_grad(perm, ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
    Float w011 = noise_grad(perm, ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f);
    Float w111 = noise_grad(perm, ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
    Float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    Float x00 = lerp_f(wx, w000, w100), x10 = lerp_f(wx, w010, w110), x01 = lerp_f(wx, w001, w101), x11 = lerp_f(wx, w011, w111);
    Float y0 = lerp_f(wy, x00, x10), y1 = lerp_f(wy, x01, x11);
    return lerp_f(wz, y0, y1);
}

inline Float ftex_texel(const FloatTexture& T, long s, long t) {  // mipmap/mod.rs:580-608
    long w = T.width, h = T.height;
    if (T.wrap == 0) { s %= w; if (s < 0) s += w; t %= h; if (t < 0) t += h; }
    else if (T.wrap == 2) { s = s < 0 ? 0 : (s > w - 1 ? w - 1 : s); t = t < 0 ? 0 : (t > h - 1 ? h - 1 : t); }
    else if (s < 0 || s >= w || t < 0 || t >= h) return 0.0f;
    return T.texels[(size_t)(t * w + s)];
}
inline int as_i32(Float f) {  // Rust `as i32`
    if (std::isnan(f)) return 0;
    if (f >= 2147483648.0f) return 0x7fffffff;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}

// Texture<Float>::evaluate for an interaction with zero derivatives at parametric (u, v).
inline Float float_texture_evaluate(const FloatTexture& T, const uint8_t* perm, Float u, Float v) {
    if (T.type == 0) return T.a;                          // constant.rs
    Float s = T.su * u + T.du, t = T.sv * v + T.dv;       // uv_2d.rs:44-50
    if (T.type == 1) {                                    // checkerboard_2d.rs:62-84 (both aa modes point-sample here)
        int sum = (int)((unsigned)as_i32(std::floor(s)) + (unsigned)as_i32(std::floor(t)));
        return sum % 2 == 0 ? T.a : T.b;
    }
    if (T.type == 2) {                                    // dots.rs:45-68
        Float s_cell = std::floor(s + 0.5f), t_cell = std::floor(t + 0.5f);
        if (noise_3d(perm, s_cell + 0.5f, t_cell + 0.5f, 0.5f) > 0.0f) {
            Float radius = 0.35f;
            Float max_shift = 0.5f - radius;
            Float s_center = s_cell + max_shift * noise_3d(perm, s_cell + 1.5f, t_cell + 2.8f, 0.5f);
            Float t_center = t_cell + max_shift * noise_3d(perm, s_cell + 4.5f, t_cell + 9.8f, 0.5f);
            Float ds = s - s_center, dt = t - t_center;
            if (ds * ds + dt * dt < radius * radius) return T.b;
        }
        return T.a;
    }
    // imagemap.rs:86-93 -> MIPMap::lookup with zero width -> triangle(0, st) (mipmap/mod.rs:293-311)
    Float ps = s * (Float)T.width - 0.5f, pt = t * (Float)T.height - 0.5f;
    long s0 = floor_as_isize(ps), t0 = floor_as_isize(pt);
    Float ds = ps - (Float)s0, dt = pt - (Float)t0;
    return ftex_texel(T, s0, t0) * (1.0f - ds) * (1.0f - dt) + ftex_texel(T, s0, t0 + 1) * (1.0f - ds) * dt + ftex_texel(T, s0 + 1, t0) * ds * (1.0f - dt) +
           ftex_texel(T, s0 + 1, t0 + 1) * ds * dt;
}

// ---- spectrum textures a material's "Kd" can name (textures/src/{constant,checkerboard_2d}.rs over UVMapping2D) ----
struct SpectrumTexture {
    int type = 0;  // B200PT_STEX_*: 0 constant, 1 checkerboard (2-D)
    Float su = 1.0f, sv = 1.0f, du = 0.0f, dv = 0.0f;
    Float tex1[3] = {1.0f, 1.0f, 1.0f}, tex2[3] = {0.0f, 0.0f, 0.0f};
    bool closedform = true;
};
// si.der's screen-space uv derivatives (SurfaceInteraction::compute_differentials, surface_interaction.rs:203-277)
struct UVDerivs {
    Float dudx = 0.0f, dvdx = 0.0f, dudy = 0.0f, dvdy = 0.0f;
};
inline Float bump_int(Float x) {  // checkerboard_2d.rs:101-103
    Float h = std::floor(x / 2.0f);
    return h + 2.0f * pmax((x / 2.0f) - h - 0.5f, 0.0f);
}
// Texture<Spectrum>::evaluate; out = 3 floats (not clamped: the material clamps, matte.rs:63)
inline void spectrum_texture_evaluate(const SpectrumTexture& T, Float u, Float v, const UVDerivs& der, Float out[3]) {
    if (T.type == 0) { out[0] = T.tex1[0]; out[1] = T.tex1[1]; out[2] = T.tex1[2]; return; }
    // uv_2d.rs:44-50
    Float dsdx = T.su * der.dudx, dtdx = T.sv * der.dvdx, dsdy = T.su * der.dudy, dtdy = T.sv * der.dvdy;
    Float s = T.su * u + T.du, t = T.sv * v + T.dv;
    auto point_sample = [&]() {
        int sum = (int)((unsigned)as_i32(std::floor(s)) + (unsigned)as_i32(std::floor(t)));
        const Float* c = sum % 2 == 0 ? T.tex1 : T.tex2;
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
    };
    if (!T.closedform) { point_sample(); return; }
    // checkerboard_2d.rs:72-97
    Float ds = pmax(pabs(dsdx), pabs(dsdy)), dt = pmax(pabs(dtdx), pabs(dtdy));
    Float s0 = s - ds, s1 = s + ds, t0 = t - dt, t1 = t + dt;
    if (std::floor(s0) == std::floor(s1) && std::floor(t0) == std::floor(t1)) { point_sample(); return; }
    Float sint = (bump_int(s1) - bump_int(s0)) / (2.0f * ds);
    Float tint = (bump_int(t1) - bump_int(t0)) / (2.0f * dt);
    Float area2 = (ds > 1.0f || dt > 1.0f) ? 0.5f : sint + tint - 2.0f * sint * tint;
    for (int c = 0; c < 3; ++c) out[c] = T.tex1[c] * (1.0f - area2) + T.tex2[c] * area2;
}

}  // namespace orc
