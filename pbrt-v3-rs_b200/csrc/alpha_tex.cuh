// Alpha-masked triangles: the float textures a mesh names as "alpha" / "shadowalpha" (shapes/src/triangle.rs:278-312),
// evaluated in the accept path of Triangle::intersect (triangle.rs:587-607) and Triangle::intersect_p (triangle.rs:840-899).
//
// The SurfaceInteraction the reference builds for the test carries the interpolated uv and ZERO differentials, so
//   * UVMapping2D (core/src/texture/mapping/uv_2d.rs) gives st = (su * u + du, sv * v + dv), dstdx = dstdy = 0;
//   * a 2-D checkerboard point-samples in both "aamode"s (checkerboard_2d.rs:62-84: with ds = dt = 0 the closed form
//     takes its "filter entirely inside one check" branch);
//   * an image map resolves to MIPMap::triangle(0, st) whatever the filter: trilinear has width 0 -> level < 0
//     (mipmap/mod.rs:226-236), EWA has minor_length == 0 (mod.rs:272-274);
//   * dots (dots.rs:45-68) needs Perlin's gradient noise (core/src/texture/common.rs:38-125).
// Constant textures never get here: they are folded into the B200PT_PRIM_ALPHA_ZERO / _SHADOW_ALPHA_ZERO flag bits.
#pragma once
#include "pt_math.cuh"

namespace b2 {

struct DFloatTex {
    int type;            // B200PT_TEX_*
    float su, sv, du, dv;
    float v0, v1;        // constant: v0; checkerboard: tex1, tex2; dots: outside_dot, inside_dot
    int wrap;            // imagemap: 0 repeat, 1 black, 2 clamp
    int width, height;
    const float* texels; // imagemap level 0: texels[t * width + s]
};

struct DeviceAlpha {
    const float* uv;          // 6 floats per ORIGINAL primitive (defaults (0,0) (1,0) (1,1) filled in for meshes without uvs)
    const int* prim_tex;      // 2 per ORIGINAL primitive: alpha, shadowalpha texture index or -1
    const DFloatTex* tex;
    const unsigned char* perm;  // NOISE_PERM[0..256)
};

// core/src/texture/common.rs:104-125
B2_D float noise_grad(const unsigned char* perm, int x, int y, int z, float dx, float dy, float dz) {
    // NOISE_PERM is the 256-entry permutation stored twice: NOISE_PERM[i] = perm[i & 255]
    int h = perm[(perm[(perm[x & 255] + y) & 255] + z) & 255] & 15;
    float u = (h < 8 || h == 12 || h == 13) ? dx : dy;
    float v = (h < 4 || h == 12 || h == 13) ? dy : dz;
    float a = (h & 1) ? -u : u;
    float b = (h & 2) ? -v : v;
    return a + b;
}
B2_D float noise_weight(float t) {
    float t3 = t * t * t;
    float t4 = t3 * t;
    return 6.0f * t4 * t - 15.0f * t4 + 10.0f * t3;
}
B2_D float noise_lerp(float t, float a, float b) { return (1.0f - t) * a + t * b; }
B2_D int floor_to_isize(float v) {  // `x.floor() as isize`: saturating, NaN -> 0; only the low 8 bits are used afterwards
    float f = floorf(v);
    if (!(f == f)) return 0;
    if (f >= 9.2233720e18f) return -1;          // isize::MAX & 255 = 255
    if (f <= -9.2233720e18f) return 0;          // isize::MIN & 255 = 0
    return (int)((long long)f & 0xffffffffll);  // low bits survive the truncation to int
}
// noise_3d, common.rs:38-73
B2_D float noise_3d(const unsigned char* perm, float x, float y, float z) {
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    int ix = floor_to_isize(x), iy = floor_to_isize(y), iz = floor_to_isize(z);
    // `ix as Float`: the saturated integer converted back (equals floor(x) whenever |x| < 2^63)
    float dx = x - fx, dy = y - fy, dz = z - fz;
    ix &= 255; iy &= 255; iz &= 255;
    float w000 = noise_grad(perm, ix, iy, iz, dx, dy, dz);
    float w100 = noise_grad(perm, ix + 1, iy, iz, dx - 1.0f, dy, dz);
    float w010 = noise_grad(perm, ix, iy + 1, iz, dx, dy - 1.0f, dz);
    float w110 = noise_grad(perm, ix + 1, iy + 1, iz, dx - 1.0f, dy - 1.0f, dz);
    float w001 = noise_grad(perm, ix, iy, iz + 1, dx, dy, dz - 1.0f);
    float w101 = noise_grad(perm, ix + 1, iy, iz + 1, dx - 1.0f, dy, dz - 1.0f);
    float w011 = noise_grad(perm, ix, iy + 1, iz + 1, dx, dy - 1.0f, dz - 1.0f);
    float w111 = noise_grad(perm, ix + 1, iy + 1, iz + 1, dx - 1.0f, dy - 1.0f, dz - 1.0f);
    float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
    float x00 = noise_lerp(wx, w000, w100);
    float x10 = noise_lerp(wx, w010, w110);
    float x01 = noise_lerp(wx, w001, w101);
    float x11 = noise_lerp(wx, w011, w111);
    float y0 = noise_lerp(wy, x00, x10);
    float y1 = noise_lerp(wy, x01, x11);
    return noise_lerp(wz, y0, y1);
}

// mipmap/mod.rs:580-608 texel() at level 0 for a Float image
B2_D float ftex_texel(const DFloatTex& T, int s, int t) {
    if (T.wrap == 0) {  // rem(): non-negative remainder
        s %= T.width; if (s < 0) s += T.width;
        t %= T.height; if (t < 0) t += T.height;
    } else if (T.wrap == 2) {
        s = s < 0 ? 0 : (s > T.width - 1 ? T.width - 1 : s);
        t = t < 0 ? 0 : (t > T.height - 1 ? T.height - 1 : t);
    } else if (s < 0 || s >= T.width || t < 0 || t >= T.height) {
        return 0.0f;
    }
    return T.texels[(long long)t * T.width + s];
}
B2_D int floor_to_int_sat(float v) {  // `.floor() as isize` / `as Int`, clamped into int range (coordinates wrap / clamp right after)
    float f = floorf(v);
    if (!(f == f)) return 0;
    return f >= 2147483520.0f ? 0x7fffff80 : (f <= -2147483520.0f ? -0x7fffff80 : (int)f);
}

B2_D float float_tex_eval(const DeviceAlpha& D, const DFloatTex& T, float u, float v) {
    if (T.type == B200PT_TEX_CONSTANT) return T.v0;
    const float s = T.su * u + T.du, t = T.sv * v + T.dv;  // UVMapping2D::map
    if (T.type == B200PT_TEX_CHECKERBOARD) {
        // `st[0].floor() as Int + st[1].floor() as Int) % 2 == 0` (i32, truncating remainder)
        const float fs = floorf(s), ft = floorf(t);
        const int is = !(fs == fs) ? 0 : (fs >= 2147483648.0f ? 0x7fffffff : (fs <= -2147483648.0f ? (int)0x80000000 : (int)fs));
        const int it = !(ft == ft) ? 0 : (ft >= 2147483648.0f ? 0x7fffffff : (ft <= -2147483648.0f ? (int)0x80000000 : (int)ft));
        const int sum = (int)((unsigned)is + (unsigned)it);  // release-mode wrapping add
        return (sum % 2 == 0) ? T.v0 : T.v1;
    }
    if (T.type == B200PT_TEX_DOTS) {
        const float s_cell = floorf(s + 0.5f), t_cell = floorf(t + 0.5f);
        if (noise_3d(D.perm, s_cell + 0.5f, t_cell + 0.5f, 0.5f) > 0.0f) {
            const float radius = 0.35f;
            const float max_shift = 0.5f - radius;
            const float s_center = s_cell + max_shift * noise_3d(D.perm, s_cell + 1.5f, t_cell + 2.8f, 0.5f);
            const float t_center = t_cell + max_shift * noise_3d(D.perm, s_cell + 4.5f, t_cell + 9.8f, 0.5f);
            const float ds = s - s_center, dt = t - t_center;
            if (ds * ds + dt * dt < radius * radius) return T.v1;
        }
        return T.v0;
    }
    // image map: MIPMap::triangle(0, st), mipmap/mod.rs:293-311
    const float ps = s * (float)T.width - 0.5f, pt = t * (float)T.height - 0.5f;
    const int s0 = floor_to_int_sat(ps), t0 = floor_to_int_sat(pt);
    const float ds = ps - (float)s0, dt = pt - (float)t0;
    return ftex_texel(T, s0, t0) * (1.0f - ds) * (1.0f - dt) + ftex_texel(T, s0, t0 + 1) * (1.0f - ds) * dt + ftex_texel(T, s0 + 1, t0) * ds * (1.0f - dt) +
           ftex_texel(T, s0 + 1, t0 + 1) * ds * dt;
}

// true = the hit survives its mesh's alpha (closest hit) / alpha and shadowalpha (any hit) textures
B2_D bool alpha_tex_accepts_inl(const DeviceAlpha* __restrict__ Dp, uint32_t prim, float b0, float b1, float b2, bool shadow) {
    const DeviceAlpha D = *Dp;
    const float* uv = D.uv + 6ll * prim;
    const float u = b0 * uv[0] + b1 * uv[2] + b2 * uv[4];  // uv_hit = b0 * uv[0] + b1 * uv[1] + b2 * uv[2] (triangle.rs:585)
    const float v = b0 * uv[1] + b1 * uv[3] + b2 * uv[5];
    const int ta = D.prim_tex[2ll * prim], ts = D.prim_tex[2ll * prim + 1];
    if (ta >= 0 && float_tex_eval(D, D.tex[ta], u, v) == 0.0f) return false;
    if (shadow && ts >= 0 && float_tex_eval(D, D.tex[ts], u, v) == 0.0f) return false;
    return true;
}

static __device__ __noinline__ bool alpha_tex_accepts(const DeviceAlpha* __restrict__ Dp, uint32_t prim, float b0, float b1, float b2, bool shadow) {
    return alpha_tex_accepts_inl(Dp, prim, b0, b1, b2, shadow);
}

}  // namespace b2
