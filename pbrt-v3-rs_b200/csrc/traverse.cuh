// Software BVH traversal + watertight triangle test for sm_100a (B200 has no
// RT cores).  Device restatement of
//   BVHAccel::intersect / intersect_p      accelerators/src/bvh/mod.rs:173-283
//   Bounds3::intersect_p_inv               core/src/geometry/bounds3.rs:292-325
//   Triangle::intersect / intersect_p      shapes/src/triangle.rs:438-545, 731-903
// with the reference's visiting order, so closest-hit primitive ids and t are
// bit-identical to the CPU path (tests/test_traversal_gpu.py).
//
// Device data layout:
//   * "wide" nodes, 64 B = 4 x float4, one per INTERIOR reference node, holding
//     BOTH children's boxes, so one 64-byte fetch (half a 128-B line, always
//     line-aligned pairs) resolves two slab tests and the dependent-load chain
//     per ray is halved versus the 32-B LinearBVHNode walk:
//        q0 = c0.min.xyz, c0.max.x   q1 = c0.max.yz, c1.min.xy
//        q2 = c1.min.z, c1.max.xyz   q3 = code0, code1, axis, -
//     c0 = first child (reference index i+1), c1 = second child (node.offset).
//     code >= 0: wide-node index of an interior child; code < 0: leaf, first
//     ordered triangle = ~code.
//   * triangles in BVHAccel.primitives (leaf) order, 64 B = 2 x 256-bit loads:
//        p0.xyz p1.xyz p2.xy | p2.z, bits(original primitive index), bits(flags),
//        bits(leaf primitive count; valid in the first triangle of a leaf), uv0 - uv2, uv1 - uv2
//   Equivalence with the reference's "test the node when it is popped": the
//   slab test of the far child is evaluated early, its entry distance is kept
//   on the stack and re-compared with the (possibly shrunk) ray.t_max when the
//   entry is popped; every other term of the test does not depend on t_max.
#pragma once
#include "pt_math.cuh"
#include "../../include/b200pt.h"
#include "alpha_tex.cuh"

namespace b2 {

struct Ray32 {  // b200pt_ray
    float ox, oy, oz, tmax, dx, dy, dz, time;
};

struct DeviceAccel {
    const float4* wide;   // 4 float4 per interior node
    const float4* tris;   // 4 float4 (64 B) per ordered triangle
    const float4* ref_nodes;  // 2 float4 per reference LinearBVHNode (baseline variant)
    float root_bounds[6];
    int root_code;        // code of the root (>=0 interior 0, <0 leaf), INT_MIN/empty => no nodes
    int n_nodes;
    long long n_prims;
    int device;           // CUDA device the arrays live on (host side: which DevCtx launches use)
    const DeviceAlpha* alpha;  // alpha-mask textures: a DeviceAlpha in device memory (null when the scene has none)
};

// The alpha test of Triangle::intersect (closest hit: "alpha") / intersect_p (any hit: "alpha" and "shadowalpha").
// Constant textures are flag bits; B200PT_PRIM_ALPHA_TEXTURE sends the hit through the texture evaluation (alpha_tex.cuh).


#define B2_EMPTY_ROOT 0x7fffffff

// Ray count of a persistent launch: the host value, or - wavefront loop - the queue size a previous kernel left in device
// memory.  Read where lanes are refilled (a uniform load every few dozen rays) instead of being held in a register.
B2_D long long ray_count(long long n, const int* n_dev) { return n_dev ? (long long)*reinterpret_cast<const volatile int*>(n_dev) : n; }

// ---- slab test -----------------------------------------------------------
// Returns the geometric part of Bounds3::intersect_p_inv and the entry
// distance; the caller applies `t_min < ray.t_max`.
struct RayCtx {
    float ox, oy, oz;
    float ix, iy, iz;   // 1/d
    int nx, ny, nz;     // dir_is_neg
};

B2_D bool slab(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float* t_entry) {
    float t_min = ((r.nx ? hix : lox) - r.ox) * r.ix;
    float t_max = ((r.nx ? lox : hix) - r.ox) * r.ix;
    float t_y_min = ((r.ny ? hiy : loy) - r.oy) * r.iy;
    float t_y_max = ((r.ny ? loy : hiy) - r.oy) * r.iy;
    t_max *= kSlabInflate;
    t_y_max *= kSlabInflate;
    if (t_min > t_y_max || t_y_min > t_max) return false;
    if (t_y_min > t_min) t_min = t_y_min;
    if (t_y_max < t_max) t_max = t_y_max;
    float t_z_min = ((r.nz ? hiz : loz) - r.oz) * r.iz;
    float t_z_max = ((r.nz ? loz : hiz) - r.oz) * r.iz;  // QUIRK: not inflated (bounds3.rs:313-319)
    if (t_min > t_z_max || t_z_min > t_max) return false;
    if (t_z_min > t_min) t_min = t_z_min;
    if (t_z_max < t_max) t_max = t_z_max;
    *t_entry = t_min;
    return t_max > 0.0f;
}

// Same test without early-out branches: every comparison keeps the reference's direction and NaN behaviour
// (`a > b` false on NaN), the results of the later axes are simply ignored when an earlier one already failed.
B2_D bool slab_bf(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float* t_entry) {
    float t_min = ((r.nx ? hix : lox) - r.ox) * r.ix;
    float t_max = ((r.nx ? lox : hix) - r.ox) * r.ix;
    float t_y_min = ((r.ny ? hiy : loy) - r.oy) * r.iy;
    float t_y_max = ((r.ny ? loy : hiy) - r.oy) * r.iy;
    float t_z_min = ((r.nz ? hiz : loz) - r.oz) * r.iz;
    float t_z_max = ((r.nz ? loz : hiz) - r.oz) * r.iz;
    t_max *= kSlabInflate;
    t_y_max *= kSlabInflate;
    bool ok = !(t_min > t_y_max) & !(t_y_min > t_max);
    t_min = t_y_min > t_min ? t_y_min : t_min;
    t_max = t_y_max < t_max ? t_y_max : t_max;
    ok &= !(t_min > t_z_max) & !(t_z_min > t_max);
    t_min = t_z_min > t_min ? t_z_min : t_min;
    t_max = t_z_max < t_max ? t_z_max : t_max;
    *t_entry = t_min;
    return ok & (t_max > 0.0f);
}

// Min/max form of the same test for rays whose direction components are all finite and non-zero and whose origin is
// finite (slab_fast_ok): then no product below is NaN, `neg ? hi : lo` picks exactly min / max of the two plane
// distances (f32 subtraction and multiplication are monotonic), the running t_min / t_max of the reference are
// max3 / min3 of the per-axis values, and the reference's six cross-axis comparisons are `t_min <= t_max`: the three
// same-axis pairs that form adds can only fail when an inflated far distance is negative, where the reference
// rejects through `t_max > 0`.  Same boolean, same t_entry (up to the sign of zero, which no comparison sees);
// 25 instead of 37 instructions per box (FMNMX3 on sm_100).
B2_D bool slab_fast(const RayCtx& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float* t_entry) {
    float ax = (lox - r.ox) * r.ix, bx = (hix - r.ox) * r.ix;
    float ay = (loy - r.oy) * r.iy, by = (hiy - r.oy) * r.iy;
    float az = (loz - r.oz) * r.iz, bz = (hiz - r.oz) * r.iz;
    float t_min = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float t_max = fminf(fminf(fmaxf(ax, bx) * kSlabInflate, fmaxf(ay, by) * kSlabInflate), fmaxf(az, bz));
    *t_entry = t_min;
    return (t_min <= t_max) & (t_max > 0.0f);
}
B2_D bool slab_fast_ok(float ox, float oy, float oz, float ix, float iy, float iz) {
    const float big = __int_as_float(0x7f7fffff);
    return (pabs(ix) <= big) & (pabs(iy) <= big) & (pabs(iz) <= big) & (ix != 0.0f) & (iy != 0.0f) & (iz != 0.0f) & (pabs(ox) <= big) & (pabs(oy) <= big) &
           (pabs(oz) <= big);
}

// ---- triangle ---------------------------------------------------------------
struct TriCtx {  // per-ray constants of the watertight test
    int kx, ky, kz;
    float sx, sy, sz;
};
B2_D TriCtx make_tri_ctx(float dx, float dy, float dz) {
    TriCtx c;
    V3 d = mk(dx, dy, dz);
    c.kz = max_dimension(vabs(d));
    c.kx = c.kz + 1; if (c.kx == 3) c.kx = 0;
    c.ky = c.kx + 1; if (c.ky == 3) c.ky = 0;
    float pdx = comp(d, c.kx), pdy = comp(d, c.ky), pdz = comp(d, c.kz);
    c.sx = -pdx / pdz;
    c.sy = -pdy / pdz;
    c.sz = 1.0f / pdz;
    return c;
}

B2_D V3 permute(V3 v, int kx, int ky, int kz) { return mk(comp(v, kx), comp(v, ky), comp(v, kz)); }

// triangle.rs:438-545.  t_max is the ray's current t_max.  On acceptance
// writes t, b0, b1, b2.
B2_D bool triangle_test(V3 o, const TriCtx& c, float t_max, V3 p0, V3 p1, V3 p2, float* t_out, float* b0_out, float* b1_out,
                        float* b2_out) {
    V3 p0t = permute(p0 - o, c.kx, c.ky, c.kz);
    V3 p1t = permute(p1 - o, c.kx, c.ky, c.kz);
    V3 p2t = permute(p2 - o, c.kx, c.ky, c.kz);
    p0t.x += c.sx * p0t.z; p0t.y += c.sy * p0t.z;
    p1t.x += c.sx * p1t.z; p1t.y += c.sy * p1t.z;
    p2t.x += c.sx * p2t.z; p2t.y += c.sy * p2t.z;
    float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // f64 fallback at edges, triangle.rs:483-495
        double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= c.sz; p1t.z *= c.sz; p2t.z *= c.sz;
    float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
    float inv_det = 1.0f / det;
    float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    float t = t_scaled * inv_det;
    float max_z_t = max_component(vabs(mk(p0t.z, p1t.z, p2t.z)));
    float delta_z = kGamma3 * max_z_t;
    float max_x_t = max_component(vabs(mk(p0t.x, p1t.x, p2t.x)));
    float max_y_t = max_component(vabs(mk(p0t.y, p1t.y, p2t.y)));
    float delta_x = kGamma5 * (max_x_t + max_z_t);
    float delta_y = kGamma5 * (max_y_t + max_z_t);
    float delta_e = 2.0f * (kGamma2 * max_x_t * max_y_t + delta_y * max_x_t + delta_x * max_y_t);
    float max_e = max_component(vabs(mk(e0, e1, e2)));
    float delta_t = 3.0f * (kGamma3 * max_e * max_z_t + delta_e * max_z_t + delta_z * max_e) * pabs(inv_det);
    if (t <= delta_t) return false;
    *t_out = t; *b0_out = b0; *b1_out = b1; *b2_out = b2;
    return true;
}

// triangle.rs:547-572: an accepted candidate is still rejected when its partial derivatives AND its geometric normal
// are degenerate.  The uvs (get_uvs, triangle.rs:384-394) enter only through duv = {uv0 - uv2, uv1 - uv2}, carried in
// the fourth float4 of the triangle record ((-1,-1),(0,-1) for the default uvs).  Returns false for "bogus" hits.
B2_D bool triangle_nondegenerate(V3 p0, V3 p1, V3 p2, float4 duv) {
    const float duv02x = duv.x, duv02y = duv.y, duv12x = duv.z, duv12y = duv.w;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = duv02x * duv12y - duv02y * duv12x;
    bool degenerate_uv = pabs(determinant) < 1e-8f;
    bool bad = degenerate_uv;
    if (!degenerate_uv) {
        float invdet = 1.0f / determinant;
        V3 dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        V3 dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
        bad = length_squared(cross(dpdu, dpdv)) == 0.0f;
    }
    if (bad) {
        V3 ng = cross(p2 - p0, p1 - p0);
        if (length_squared(ng) == 0.0f) return false;
    }
    return true;
}

// The uv differences are fetched only for a candidate that passed the triangle test (rare): keeping them live across
// the test costs four registers, which pushed the default kernels over their occupancy boundary (-9 %).
// Out of line on purpose: it runs once per accepted candidate and must not shape the register allocation of the walk.
static __device__ __noinline__ bool triangle_nondegenerate(V3 p0, V3 p1, V3 p2, const float4* tris, long long i) {
    return triangle_nondegenerate(p0, p1, p2, __ldg(tris + 4 * i + 3));
}

B2_D void load_tri(const float4* tris, long long i, V3* p0, V3* p1, V3* p2, uint32_t* prim, uint32_t* flags, uint32_t* leaf_n) {
    float4 a, b, c, d;
    ldg8(tris + 4 * i, &a, &b);
    ldg8(tris + 4 * i + 2, &c, &d);
    *p0 = mk(a.x, a.y, a.z); *p1 = mk(a.w, b.x, b.y); *p2 = mk(b.z, b.w, c.x);
    *prim = __float_as_uint(c.y); *flags = __float_as_uint(c.z); *leaf_n = __float_as_uint(c.w);
}

// The alpha test of Triangle::intersect (closest hit: "alpha") / intersect_p (any hit: "alpha" and "shadowalpha").
// Constant textures are flag bits; B200PT_PRIM_ALPHA_TEXTURE sends the hit through the texture evaluation (alpha_tex.cuh).
// kTex = false: an instantiation for accelerators without alpha textures (no out-of-line call in the kernel at all; the
// default kernels are launched that way unless the scene has textures - the call costs the any-hit kernel two spilled
// registers in its inner loop, 6 % on C2).
template <bool ANY, bool kTex = true>
B2_D bool alpha_ok(const DeviceAccel& A, uint32_t flags, uint32_t prim, float b0, float b1, float b2) {
    if (flags & (ANY ? 6u : 2u)) return false;
    if (!kTex) return true;
    if (!(flags & B200PT_PRIM_ALPHA_TEXTURE)) return true;
    return alpha_tex_accepts(A.alpha, prim, b0, b1, b2, ANY);
}
// Any-hit form: the kernels do not keep the barycentrics of an accepted candidate, so the (rare) textured triangle is
// tested once more out of line - only values that are live in the walk anyway cross the call.
static __device__ __noinline__ bool alpha_any_retest(const DeviceAlpha* __restrict__ Dp, const float4* __restrict__ tris, long long i, float4 o_tmax, float4 s_k) {
    const float4 ra = __ldg(tris + 4 * i), rb = __ldg(tris + 4 * i + 1), rc = __ldg(tris + 4 * i + 2);
    const V3 p0 = mk(ra.x, ra.y, ra.z), p1 = mk(ra.w, rb.x, rb.y), p2 = mk(rb.z, rb.w, rc.x);
    const uint32_t prim = __float_as_uint(rc.y);
    TriCtx tc;
    const int pk = __float_as_int(s_k.w);
    tc.kx = pk & 3; tc.ky = (pk >> 2) & 3; tc.kz = (pk >> 4) & 3; tc.sx = s_k.x; tc.sy = s_k.y; tc.sz = s_k.z;
    float t, b0, b1, b2;
    if (!triangle_test(mk(o_tmax.x, o_tmax.y, o_tmax.z), tc, o_tmax.w, p0, p1, p2, &t, &b0, &b1, &b2)) return false;  // cannot happen: same inputs as the caller's test
    return alpha_tex_accepts_inl(Dp, prim, b0, b1, b2, true);
}
template <bool kTex = true>
B2_D bool alpha_ok_any(const DeviceAccel& A, uint32_t flags, long long tri_index, V3 o, const TriCtx& tc, float t_max) {
    if (flags & 6u) return false;
    if (!kTex) return true;
    if (!(flags & B200PT_PRIM_ALPHA_TEXTURE)) return true;
    return alpha_any_retest(A.alpha, A.tris, tri_index, make_float4(o.x, o.y, o.z, t_max), make_float4(tc.sx, tc.sy, tc.sz, __int_as_float(tc.kx | (tc.ky << 2) | (tc.kz << 4))));
}

struct HitOut {
    float t;
    uint32_t prim;
    float b0, b1;
    float b2;  // third barycentric exactly as the test produced it (e2 * inv_det); kept for the path tracer
};

#ifndef B2_STACK
#define B2_STACK 64  // reference stack depth, mod.rs:185
#endif

// Closest hit, wide-node layout.  ANY = any-hit (intersect_p) variant.
template <bool ANY>
B2_D bool traverse_wide(const DeviceAccel& A, const Ray32& ray, HitOut* out) {
    out->t = __int_as_float(0x7f800000);
    out->prim = 0xffffffffu;
    out->b0 = 0.0f; out->b1 = 0.0f; out->b2 = 0.0f;
    if (A.root_code == B2_EMPTY_ROOT) return false;
    RayCtx r;
    r.ox = ray.ox; r.oy = ray.oy; r.oz = ray.oz;
    r.ix = 1.0f / ray.dx; r.iy = 1.0f / ray.dy; r.iz = 1.0f / ray.dz;
    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
    float t_max = ray.tmax;
    float te;
    if (!(slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) &&
          te < t_max))
        return false;
    const TriCtx tc = make_tri_ctx(ray.dx, ray.dy, ray.dz);
    const V3 o = mk(ray.ox, ray.oy, ray.oz);
    int stack_code[B2_STACK];
    float stack_t[B2_STACK];
    int sp = 0;
    int cur = A.root_code;
    bool hit = false;
    for (;;) {
        if (cur >= 0) {
            const float4* q = A.wide + 4ll * cur;
            float4 q0, q1, q2, q3;
            ldg8(q, &q0, &q1);
            ldg8(q + 2, &q2, &q3);
            float t0, t1;
            bool h0 = slab(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) && t0 < t_max;
            bool h1 = slab(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) && t1 < t_max;
            int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y), axis = __float_as_int(q3.z);
            int neg = axis == 0 ? r.nx : (axis == 1 ? r.ny : r.nz);
            // reference order: neg ? (second first, push first) : (first first, push second)
            int near_c = neg ? c1 : c0, far_c = neg ? c0 : c1;
            bool near_h = neg ? h1 : h0, far_h = neg ? h0 : h1;
            float far_t = neg ? t0 : t1;
            if (near_h) {
                if (far_h) { stack_code[sp] = far_c; stack_t[sp] = far_t; ++sp; }
                cur = near_c;
                continue;
            }
            if (far_h) { cur = far_c; continue; }
        } else {
            long long first = (long long)(~cur);
            V3 p0, p1, p2;
            uint32_t prim, flags, leaf_n;
            load_tri(A.tris, first, &p0, &p1, &p2, &prim, &flags, &leaf_n);
            for (uint32_t i = 0;;) {
                float t, b0, b1, b2;
                if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, first + i)) {
                    if (ANY) {
                        if (alpha_ok_any(A, flags, first + i, o, tc, t_max)) return true;  // alpha / shadow-alpha == 0 reject (triangle.rs:886-899)
                    } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {          // alpha == 0 reject (triangle.rs:587-607)
                        hit = true;
                        t_max = t;
                        out->t = t; out->prim = prim; out->b0 = b0; out->b1 = b1; out->b2 = b2;
                    }
                }
                if (++i >= leaf_n) break;
                uint32_t dummy;
                load_tri(A.tris, first + i, &p0, &p1, &p2, &prim, &flags, &dummy);
            }
        }
        // pop
        for (;;) {
            if (sp == 0) return hit;
            --sp;
            cur = stack_code[sp];
            if (ANY || stack_t[sp] < t_max) break;
        }
    }
}

// Baseline: literal walk over the 32-byte LinearBVHNode array (2 x float4 per
// node, node tested when visited), kept for A/B measurement (variant 1).
template <bool ANY>
B2_D bool traverse_ref(const DeviceAccel& A, const Ray32& ray, HitOut* out) {
    out->t = __int_as_float(0x7f800000);
    out->prim = 0xffffffffu;
    out->b0 = 0.0f; out->b1 = 0.0f; out->b2 = 0.0f;
    if (A.root_code == B2_EMPTY_ROOT) return false;
    RayCtx r;
    r.ox = ray.ox; r.oy = ray.oy; r.oz = ray.oz;
    r.ix = 1.0f / ray.dx; r.iy = 1.0f / ray.dy; r.iz = 1.0f / ray.dz;
    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
    float t_max = ray.tmax;
    const TriCtx tc = make_tri_ctx(ray.dx, ray.dy, ray.dz);
    const V3 o = mk(ray.ox, ray.oy, ray.oz);
    int stack[B2_STACK];
    int sp = 0, cur = 0;
    bool hit = false;
    for (;;) {
        float4 n0, n1;
        ldg8(A.ref_nodes + 2ll * cur, &n0, &n1);
        float te;
        uint32_t offset = __float_as_uint(n1.z), meta = __float_as_uint(n1.w);
        uint32_t nprims = meta & 0xffffu, axis = (meta >> 16) & 0xffu;
        if (slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, &te) && te < t_max) {
            if (nprims > 0) {
                for (uint32_t i = 0; i < nprims; ++i) {
                    V3 p0, p1, p2;
                    uint32_t prim, flags, dummy;
                    load_tri(A.tris, (long long)offset + i, &p0, &p1, &p2, &prim, &flags, &dummy);
                    float t, b0, b1, b2;
                    if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)offset + i)) {
                        if (ANY) {
                            if (alpha_ok_any(A, flags, (long long)offset + i, o, tc, t_max)) return true;
                        } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                            hit = true;
                            t_max = t;
                            out->t = t; out->prim = prim; out->b0 = b0; out->b1 = b1; out->b2 = b2;
                        }
                    }
                }
                if (sp == 0) break;
                cur = stack[--sp];
            } else {
                int neg = axis == 0 ? r.nx : (axis == 1 ? r.ny : r.nz);
                if (neg) { stack[sp++] = cur + 1; cur = (int)offset; }
                else { stack[sp++] = (int)offset; cur = cur + 1; }
            }
        } else {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return hit;
}

}  // namespace b2
