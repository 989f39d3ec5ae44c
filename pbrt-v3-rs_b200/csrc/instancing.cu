// Two-level traversal: TransformedPrimitive (core/src/primitives/transformed_primitive.rs:43-73) with a static
// transform among the primitives of the scene aggregate, each referring to an object's own BVHAccel
// (api/src/lib.rs:937-987).  Config C5 (ecosys-style instancing).
//
// Device layout: ONE wide-node array and ONE 64-byte record array hold the top-level BVH followed by every object's
// BVH with global indices, so the single-level walk (traverse_wide) is reused unchanged for the nested call.  A
// top-level leaf record is either a triangle or — flag bit 31 — an instance (id in the primitive field).
// TransformedPrimitive::intersect semantics kept: the ray is taken to instance space by Transform::transform_ray
// (origin nudged by its error bound, t_max shortened by the same dt, transform.rs:451-476), the nested aggregate is
// intersected, and the INSTANCE-space t_max is written back to the world ray (transformed_primitive.rs:52-56).
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "instancing.cuh"
#include "traverse_phased.cuh"

namespace b2 {

// Transform::transform_ray (transform.rs:307-331, 373-380, 451-476) with the full 4x4 world_to_instance: the
// reference inverts with Gauss-Jordan, so the last row of the inverse is not always exactly (0,0,0,1) and the
// homogeneous divide `if wp == 1 { p } else { p / wp }` is kept.  m0..m3 = rows 0..3.
B2_D Ray32 xf_ray(float4 m0, float4 m1, float4 m2, float4 m3, const Ray32& r) {
    float x = r.ox, y = r.oy, z = r.oz;
    float xp = (m0.x * x + m0.y * y) + (m0.z * z + m0.w);
    float yp = (m1.x * x + m1.y * y) + (m1.z * z + m1.w);
    float zp = (m2.x * x + m2.y * y) + (m2.z * z + m2.w);
    float wp = (m3.x * x + m3.y * y) + (m3.z * z + m3.w);
    float xs = pabs(m0.x * x) + pabs(m0.y * y) + pabs(m0.z * z) + pabs(m0.w);
    float ys = pabs(m1.x * x) + pabs(m1.y * y) + pabs(m1.z * z) + pabs(m1.w);
    float zs = pabs(m2.x * x) + pabs(m2.y * y) + pabs(m2.z * z) + pabs(m2.w);
    V3 o_err = kGamma3 * mk(xs, ys, zs);
    V3 o = mk(xp, yp, zp);
    if (!(wp == 1.0f)) o = o / wp;
    V3 d = mk(m0.x * r.dx + m0.y * r.dy + m0.z * r.dz, m1.x * r.dx + m1.y * r.dy + m1.z * r.dz, m2.x * r.dx + m2.y * r.dy + m2.z * r.dz);
    float l2 = length_squared(d);
    float t_max = r.tmax;
    if (l2 > 0.0f) {
        float dt = dot(vabs(d), o_err) / l2;
        o = o + d * dt;
        t_max -= dt;
    }
    Ray32 q;
    q.ox = o.x; q.oy = o.y; q.oz = o.z; q.tmax = t_max; q.dx = d.x; q.dy = d.y; q.dz = d.z; q.time = r.time;
    return q;
}

// BVHAccel::intersect / intersect_p over the scene aggregate (same walk as traverse_wide; leaves may hold instances).
template <bool ANY>
B2_D bool traverse_top(const DeviceAccel2& A2, const Ray32& ray, HitOut* out, int* inst_out) {
    const DeviceAccel& A = A2.top;
    out->t = __int_as_float(0x7f800000);
    out->prim = 0xffffffffu;
    out->b0 = 0.0f; out->b1 = 0.0f; out->b2 = 0.0f;
    *inst_out = -1;
    if (A.root_code == B2_EMPTY_ROOT) return false;
    RayCtx r;
    r.ox = ray.ox; r.oy = ray.oy; r.oz = ray.oz;
    r.ix = 1.0f / ray.dx; r.iy = 1.0f / ray.dy; r.iz = 1.0f / ray.dz;
    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
    float t_max = ray.tmax;
    float te;
    if (!(slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max)) return false;
    const TriCtx tc = make_tri_ctx(ray.dx, ray.dy, ray.dz);
    const V3 o = mk(ray.ox, ray.oy, ray.oz);
    int stack_code[B2_STACK];
    float stack_t[B2_STACK];
    int sp = 0, cur = A.root_code;
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
                if (flags & 0x80000000u) {  // TransformedPrimitive
                    const float4* T = A2.inst_trav + 6ll * prim;
                    const float4 m0 = __ldg(T), m1 = __ldg(T + 1), m2 = __ldg(T + 2), m3 = __ldg(T + 3), b0q = __ldg(T + 4), b1q = __ldg(T + 5);
                    Ray32 wr = ray;
                    wr.tmax = t_max;
                    Ray32 ir = xf_ray(m0, m1, m2, m3, wr);
                    DeviceAccel oa = A;
                    oa.root_code = __float_as_int(b1q.z);
                    oa.root_bounds[0] = b0q.x; oa.root_bounds[1] = b0q.y; oa.root_bounds[2] = b0q.z;
                    oa.root_bounds[3] = b0q.w; oa.root_bounds[4] = b1q.x; oa.root_bounds[5] = b1q.y;
                    HitOut h2;
                    if (traverse_wide<ANY>(oa, ir, &h2)) {
                        if (ANY) return true;
                        hit = true;
                        t_max = h2.t;  // r.t_max = ray.t_max (instance-space value), transformed_primitive.rs:55
                        *out = h2;
                        *inst_out = (int)prim;
                    }
                } else {
                    float t, b0, b1, b2;
                    if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, first + i)) {
                        if (ANY) {
                            if (alpha_ok_any(A, flags, first + i, o, tc, t_max)) return true;
                        } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                            hit = true;
                            t_max = t;
                            out->t = t; out->prim = prim; out->b0 = b0; out->b1 = b1; out->b2 = b2;
                            *inst_out = -1;
                        }
                    }
                }
                if (++i >= leaf_n) break;
                uint32_t dummy;
                load_tri(A.tris, first + i, &p0, &p1, &p2, &prim, &flags, &dummy);
            }
        }
        for (;;) {
            if (sp == 0) return hit;
            --sp;
            cur = stack_code[sp];
            if (ANY || stack_t[sp] < t_max) break;
        }
    }
}

template <bool ANY>
__global__ void __launch_bounds__(128) k_trace_twolevel(DeviceAccel2 A, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                        float* __restrict__ b2_out, int* __restrict__ inst_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r0 = __ldg(rays + 2 * i), r1 = __ldg(rays + 2 * i + 1);
    Ray32 ray{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    HitOut h;
    int inst;
    bool hit = traverse_top<ANY>(A, ray, &h, &inst);
    if (ANY) ((uint8_t*)out)[i] = hit ? 1 : 0;
    else {
        ((float4*)out)[i] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1);
        if (b2_out) b2_out[i] = h.b2;
        if (inst_out) inst_out[i] = inst;
    }
}

// Phase-scheduled persistent two-level walk (variant 0, default): the kernel of traverse_phased.cuh with the instance
// entered and left inside the same loop, so a warp keeps lanes that walk the scene aggregate and lanes that walk an
// object's BVH in the same NODE / TRI phases (both levels share one node array and one record array).
// Per lane: `sp_base` is the stack height at which the instance was entered (the nested walk pops down to it),
// the interrupted top-level leaf is kept in (saved_cur, saved_i, saved_left) and the world ray is re-read from the
// input when the instance is left.  Per ray the order of node tests, triangle tests and t_max updates is
// TransformedPrimitive::intersect's (transformed_primitive.rs:51-64): on a hit the INSTANCE-space t_max stays in the
// world ray, without one the world t_max is restored.
#define B2_STACK2 96  // scene aggregate + object walk share one stack (the reference has 64 entries per level)

template <bool ANY, int kSwitch, int kRefill, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks) k_trace_phased2(DeviceAccel2 A2, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                                unsigned long long* __restrict__ counter, float* __restrict__ b2_out, int* __restrict__ inst_out,
                                                                 const int* __restrict__ n_dev) {
    const DeviceAccel& A = A2.top;
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    int stack_code[B2_STACK2];
    float stack_t[B2_STACK2];

    long long ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f, world_t_max = 0.0f;
    int cur = kIdle, sp = 0, sp_base = 0;
    long long tri_i = 0, saved_i = 0;
    uint32_t tri_left = 0, saved_left = 0;
    int saved_cur = 0;
    int in_inst = -1;        // instance being walked, -1 at the top level
    bool inst_hit = false;   // the current instance produced a hit
    bool hit = false;
    HitOut h;
    int h_inst = -1;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;
    bool node_phase = true;

    auto set_ray = [&](float ox, float oy, float oz, float dx, float dy, float dz) {
        r.ox = ox; r.oy = oy; r.oz = oz;
        r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;
        r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
        tc = make_tri_ctx(dx, dy, dz);
        o = mk(ox, oy, oz);
    };

    for (;;) {
        const unsigned idle_mask = __ballot_sync(0xffffffffu, cur == kIdle);
        if (idle_mask == 0xffffffffu && exhausted) break;
        if (!exhausted && __popc(idle_mask) >= kRefill) {
            const int want = __popc(idle_mask);
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(counter, (unsigned long long)want);
            b = __shfl_sync(0xffffffffu, b, 0);
            const long long n_rays = ray_count(n, n_dev);
            if ((long long)b + want >= n_rays) exhausted = true;
            if (cur == kIdle) {
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n_rays) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = id;
                    set_ray(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                    t_max = r0.w;
                    sp = 0; sp_base = 0; hit = false; tri_left = 0; in_inst = -1; h_inst = -1;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
                    float te;
                    bool enter = A.root_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) cur = A.root_code;
                    else {
                        if (ANY) ((uint8_t*)out)[id] = 0;
                        else {
                            ((float4*)out)[id] = make_float4(h.t, __uint_as_float(h.prim), 0.0f, 0.0f);
                            if (b2_out) b2_out[id] = 0.0f;
                            if (inst_out) inst_out[id] = -1;
                        }
                    }
                }
            }
        }
        for (;;) {
            const unsigned m_node = __ballot_sync(0xffffffffu, cur >= 0);
            const unsigned m_tri = __ballot_sync(0xffffffffu, cur < 0 && cur != kIdle);
            if (!(m_node | m_tri)) break;
            if (!exhausted && __popc(~(m_node | m_tri)) >= kRefill) break;
            const int nn = __popc(m_node), nt = __popc(m_tri);
            if (node_phase) { if (nn < kSwitch && nt > nn) node_phase = false; }
            else            { if (nt < kSwitch && nn > nt) node_phase = true; }
            if (nt == 0) node_phase = true;
            if (nn == 0) node_phase = false;

            bool retire = false;
            if (node_phase) {
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
                    int near_c = neg ? c1 : c0, far_c = neg ? c0 : c1;
                    bool near_h = neg ? h1 : h0, far_h = neg ? h0 : h1;
                    float far_t = neg ? t0 : t1;
                    if (near_h) {
                        if (far_h) { stack_code[sp] = far_c; stack_t[sp] = far_t; ++sp; }
                        cur = near_c;
                    } else if (far_h) {
                        cur = far_c;
                    } else {
                        retire = true;
                    }
                    tri_left = 0;
                }
            } else if (cur < 0 && cur != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = (long long)(~cur);
                load_tri(A.tris, tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                ++tri_i;
                --tri_left;
                if (flags & 0x80000000u) {
                    // TransformedPrimitive: take the ray to instance space and start the object's walk at its root.
                    const float4* T = A2.inst_trav + 6ll * prim;
                    const float4 m0 = __ldg(T), m1 = __ldg(T + 1), m2 = __ldg(T + 2), m3 = __ldg(T + 3), b0q = __ldg(T + 4), b1q = __ldg(T + 5);
                    const float4 w1 = __ldg(rays + 2 * ray_id + 1);
                    Ray32 wr{o.x, o.y, o.z, t_max, w1.x, w1.y, w1.z, w1.w};
                    const Ray32 ir = xf_ray(m0, m1, m2, m3, wr);
                    RayCtx ri;
                    ri.ox = ir.ox; ri.oy = ir.oy; ri.oz = ir.oz;
                    ri.ix = 1.0f / ir.dx; ri.iy = 1.0f / ir.dy; ri.iz = 1.0f / ir.dz;
                    ri.nx = ri.ix < 0.0f; ri.ny = ri.iy < 0.0f; ri.nz = ri.iz < 0.0f;
                    float te;
                    const int root = __float_as_int(b1q.z);
                    if (root != B2_EMPTY_ROOT && slab(ri, b0q.x, b0q.y, b0q.z, b0q.w, b1q.x, b1q.y, &te) && te < ir.tmax) {
                        saved_cur = cur; saved_i = tri_i; saved_left = tri_left;
                        world_t_max = t_max;
                        in_inst = (int)prim; inst_hit = false;
                        sp_base = sp;
                        r = ri;
                        tc = make_tri_ctx(ir.dx, ir.dy, ir.dz);
                        o = mk(ir.ox, ir.oy, ir.oz);
                        t_max = ir.tmax;
                        cur = root;
                        tri_left = 0;
                    } else if (tri_left == 0) retire = true;
                } else {
                    float t, b0, b1, b2;
                    if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, tri_i - 1)) {
                        if (ANY) {
                            if (alpha_ok_any(A, flags, tri_i - 1, o, tc, t_max)) { hit = true; sp = 0; sp_base = 0; in_inst = -1; tri_left = 0; }
                        } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                            hit = true;
                            t_max = t;
                            h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                            h_inst = in_inst;
                            inst_hit = true;
                        }
                    }
                    if (tri_left == 0) retire = true;
                }
            }
            if (retire) {
                for (;;) {
                    cur = kIdle;
                    while (sp > sp_base) {
                        --sp;
                        if (ANY || stack_t[sp] < t_max) { cur = stack_code[sp]; break; }
                    }
                    tri_left = 0;
                    if (cur != kIdle || in_inst < 0) break;
                    // the object's walk is finished: back to the interrupted leaf of the scene aggregate
                    if (!inst_hit) t_max = world_t_max;
                    in_inst = -1;
                    sp_base = 0;
                    const float4 w0 = __ldg(rays + 2 * ray_id), w1 = __ldg(rays + 2 * ray_id + 1);
                    set_ray(w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
                    if (saved_left > 0) { cur = saved_cur; tri_i = saved_i; tri_left = saved_left; break; }
                }
                if (cur == kIdle) {
                    if (ANY) ((uint8_t*)out)[ray_id] = hit ? 1 : 0;
                    else {
                        ((float4*)out)[ray_id] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1);
                        if (b2_out) b2_out[ray_id] = h.b2;
                        if (inst_out) inst_out[ray_id] = h_inst;
                    }
                }
            }
        }
    }
}

// Loop-free postponed-leaf form of the two-level walk (variant 0, default; the single-level kernel is k_trace_spec2 in
// traverse_spec.cuh, where the equivalence argument is written down).  Inside an object's BVH a lane parks the leaf it
// reached and keeps walking; at the scene-aggregate level a leaf is NOT walked past (cur = kHold): its records may be
// TransformedPrimitives, and entering one replaces the lane's ray.  Entering spills the register-held stack top and
// records sp_base; the object's walk pops down to sp_base only; leaving restores the world ray from the input, the
// interrupted leaf (saved_*) and the stack top.
template <bool ANY, int kSwitch, int kRefill, int kBlocks, bool kTex = true>
__global__ void __launch_bounds__(128, kBlocks) k_trace_spec2_2l(DeviceAccel2 A2, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                                 unsigned long long* __restrict__ counter, float* __restrict__ b2_out, int* __restrict__ inst_out,
                                                                 const int* __restrict__ n_dev) {
    const DeviceAccel& A = A2.top;
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    const int kRetry = (int)0x80000001;  // pop (again) in the next NODE step
    const int kHold = (int)0x80000002;   // scene-aggregate level: wait until the parked leaf has been processed, then pop
    StackEntry<ANY> stack[B2_STACK2];
    // The world-space ray context (origin, reciprocals, watertight-test constants) while the lane walks an object: kept in
    // local memory (volatile: not promoted to registers) instead of being re-derived from the input ray on return.
    // profiles/r2_ncu_full_c5_k_trace_spec2_2l_sass.csv.gz: the re-derivation (three divisions, make_tri_ctx: ~250
    // instructions) ran with 1.0 of 32 lanes, 19.6 M times per 16.6 M rays = 22 % of the kernel's issued instructions.
    volatile float wsave[10];

    int ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f, world_t_max = 0.0f;
    int cur = kIdle;
    float cur_t = 0.0f;
    int pend = kIdle;
    int sp = 0, sp_base = 0;
    int top_code = kIdle;
    float top_t = 0.0f;
    int negmask = 0;
    int tri_i = 0, saved_i = 0;
    uint32_t tri_left = 0, saved_left = 0;
    int in_inst = -1;
    bool inst_hit = false;
    HitOut h;
    int h_inst = -1;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;
    bool node_phase = true;

    auto set_ray = [&](float ox, float oy, float oz, float dx, float dy, float dz) {
        r.ox = ox; r.oy = oy; r.oz = oz;
        r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;
        r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
        negmask = r.nx | (r.ny << 1) | (r.nz << 2);
        tc = make_tri_ctx(dx, dy, dz);
        o = mk(ox, oy, oz);
    };

    for (;;) {
        const unsigned idle_mask = __ballot_sync(0xffffffffu, cur == kIdle && pend == kIdle);
        if (idle_mask == 0xffffffffu && exhausted) break;
        if (!exhausted && __popc(idle_mask) >= kRefill) {
            const int want = __popc(idle_mask);
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(counter, (unsigned long long)want);
            b = __shfl_sync(0xffffffffu, b, 0);
            const long long n_rays = ray_count(n, n_dev);
            if ((long long)b + want >= n_rays) exhausted = true;
            if (cur == kIdle && pend == kIdle) {
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n_rays) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = (int)id;
                    set_ray(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                    t_max = r0.w;
                    sp = 0; sp_base = 0; tri_left = 0; top_code = kIdle; in_inst = -1; h_inst = -1;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
                    float te;
                    bool enter = A.root_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) {
                        if (A.root_code >= 0) { cur = A.root_code; cur_t = te; }
                        else { pend = A.root_code; cur = kHold; }
                    } else if (ANY) {
                        ((uint8_t*)out)[id] = 0;
                    } else {
                        ((float4*)out)[id] = make_float4(h.t, __uint_as_float(h.prim), 0.0f, 0.0f);
                        if (b2_out) b2_out[id] = 0.0f;
                        if (inst_out) inst_out[id] = -1;
                    }
                }
            }
        }
        for (;;) {
            const unsigned m_node = __ballot_sync(0xffffffffu, cur >= 0 || cur == kRetry);
            const unsigned m_tri = __ballot_sync(0xffffffffu, pend != kIdle);
            if (!(m_node | m_tri)) break;
            if (!exhausted && __popc(~(m_node | m_tri)) >= kRefill) break;
            const int nn = __popc(m_node), nt = __popc(m_tri);
            if (node_phase) { if (nn < kSwitch && nt > nn) node_phase = false; }
            else            { if (nt < kSwitch && nn > nt) node_phase = true; }
            if (nt == 0) node_phase = true;
            if (nn == 0) node_phase = false;

            bool fin = false;  // this level's walk is finished (cur and pend both empty)
            if (node_phase) {
                bool need_pop = cur == kRetry;
                if (cur >= 0) {
                    const float4* q = A.wide + 4ll * cur;
                    float4 q0, q1, q2, q3;
                    ldg8(q, &q0, &q1);
                    ldg8(q + 2, &q2, &q3);
                    float t0, t1;
                    // literal box test: an instance-space ray may be axis-parallel where the world ray is not, and the
                    // min/max form is chosen per warp at refills only
                    const bool h0 = slab_bf(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                    const bool h1 = slab_bf(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                    const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y), axis = __float_as_int(q3.z);
                    const bool neg = (negmask >> axis) & 1;
                    const int near_c = neg ? c1 : c0, far_c = neg ? c0 : c1;
                    const bool near_h = neg ? h1 : h0, far_h = neg ? h0 : h1;
                    const float near_t = neg ? t1 : t0, far_t = neg ? t0 : t1;
                    const bool push = near_h & far_h;
                    const bool spill = push & (top_code != kIdle);
                    if (spill) stack[sp].set(top_code, top_t);
                    sp += spill ? 1 : 0;
                    top_code = push ? far_c : top_code;
                    top_t = push ? far_t : top_t;
                    cur = near_h ? near_c : far_c;
                    cur_t = near_h ? near_t : far_t;
                    need_pop = !(near_h | far_h);
                    const bool park = !need_pop & (cur < 0) & (pend == kIdle);
                    pend = park ? cur : pend;
                    tri_left = park ? 0u : tri_left;
                    if (park) { if (in_inst >= 0) need_pop = true; else cur = kHold; }
                }
                if (need_pop) {
                    const int c = top_code;
                    const float t = top_t;
                    const bool have = c != kIdle;
                    const bool refill = have & (sp > sp_base);
                    sp -= refill ? 1 : 0;
                    StackEntry<ANY> e;
                    e.set(kIdle, 0.0f);
                    if (refill) e = stack[sp];
                    top_code = e.code(); top_t = e.t();
                    const bool valid = have & (ANY || t < t_max);
                    cur = valid ? c : (have ? kRetry : kIdle);
                    cur_t = t;
                    const bool park = valid & (c < 0) & (pend == kIdle);
                    pend = park ? c : pend;
                    tri_left = park ? 0u : tri_left;
                    cur = park ? (in_inst >= 0 ? kRetry : kHold) : cur;
                    fin = (cur == kIdle) & (pend == kIdle);
                }
            } else if (pend != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = ~pend;
                load_tri(A.tris, (long long)tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                ++tri_i;
                --tri_left;
                bool entered = false;
                if (flags & 0x80000000u) {
                    // TransformedPrimitive (only in leaves of the scene aggregate, where cur == kHold)
                    const float4* T = A2.inst_trav + 6ll * prim;
                    const float4 m0 = __ldg(T), m1 = __ldg(T + 1), m2 = __ldg(T + 2), m3 = __ldg(T + 3), b0q = __ldg(T + 4), b1q = __ldg(T + 5);
                    const float4 w1 = __ldg(rays + 2ll * ray_id + 1);
                    Ray32 wr{o.x, o.y, o.z, t_max, w1.x, w1.y, w1.z, w1.w};
                    const Ray32 ir = xf_ray(m0, m1, m2, m3, wr);
                    RayCtx ri;
                    ri.ox = ir.ox; ri.oy = ir.oy; ri.oz = ir.oz;
                    ri.ix = 1.0f / ir.dx; ri.iy = 1.0f / ir.dy; ri.iz = 1.0f / ir.dz;
                    ri.nx = ri.ix < 0.0f; ri.ny = ri.iy < 0.0f; ri.nz = ri.iz < 0.0f;
                    float te;
                    const int root = __float_as_int(b1q.z);
                    if (root != B2_EMPTY_ROOT && slab(ri, b0q.x, b0q.y, b0q.z, b0q.w, b1q.x, b1q.y, &te) && te < ir.tmax) {
                        entered = true;
                        saved_i = tri_i; saved_left = tri_left;
                        world_t_max = t_max;
                        in_inst = (int)prim; inst_hit = false;
                        if (top_code != kIdle) { stack[sp].set(top_code, top_t); ++sp; top_code = kIdle; }
                        sp_base = sp;
                        wsave[0] = r.ox; wsave[1] = r.oy; wsave[2] = r.oz; wsave[3] = r.ix; wsave[4] = r.iy; wsave[5] = r.iz;
                        wsave[6] = tc.sx; wsave[7] = tc.sy; wsave[8] = tc.sz;
                        wsave[9] = __int_as_float(tc.kx | (tc.ky << 2) | (tc.kz << 4) | (negmask << 6));
                        r = ri;
                        negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                        tc = make_tri_ctx(ir.dx, ir.dy, ir.dz);
                        o = mk(ir.ox, ir.oy, ir.oz);
                        t_max = ir.tmax;
                        tri_left = 0;
                        if (root >= 0) { cur = root; cur_t = te; pend = kIdle; }
                        else { pend = root; cur = kRetry; }  // single-leaf object
                    }
                } else {
                    float t, b0, b1, b2;
                    if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)tri_i - 1)) {
                        if (ANY) {
                            if (alpha_ok_any<kTex>(A, flags, (long long)tri_i - 1, o, tc, t_max)) { h.prim = 0u; cur = kIdle; top_code = kIdle; sp = 0; sp_base = 0; in_inst = -1; tri_left = 0; }
                        } else if (alpha_ok<false, kTex>(A, flags, prim, b0, b1, b2)) {
                            t_max = t;
                            h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                            h_inst = in_inst;
                            inst_hit = true;
                        }
                    }
                }
                if (!entered && tri_left == 0) {
                    // leaf done, t_max current again: re-validate what was reached speculatively (object level only)
                    pend = kIdle;
                    const bool live = cur != kIdle && cur != kRetry && cur != kHold;
                    if (cur == kHold) cur = kRetry;
                    else if (!ANY && live && !(cur_t < t_max)) cur = kRetry;
                    else if (live && cur < 0) { pend = cur; cur = kRetry; }
                    fin = cur == kIdle;
                }
            }
            if (fin && in_inst >= 0) {
                // the object's walk is finished: back to the interrupted leaf of the scene aggregate
                fin = false;
                if (!inst_hit) t_max = world_t_max;
                in_inst = -1;
                sp_base = 0;
                {
                    r.ox = wsave[0]; r.oy = wsave[1]; r.oz = wsave[2]; r.ix = wsave[3]; r.iy = wsave[4]; r.iz = wsave[5];
                    tc.sx = wsave[6]; tc.sy = wsave[7]; tc.sz = wsave[8];
                    const int pk = __float_as_int(wsave[9]);
                    tc.kx = pk & 3; tc.ky = (pk >> 2) & 3; tc.kz = (pk >> 4) & 3;
                    negmask = (pk >> 6) & 7;
                    r.nx = negmask & 1; r.ny = (negmask >> 1) & 1; r.nz = (negmask >> 2) & 1;
                    o = mk(r.ox, r.oy, r.oz);
                }
                if (sp > 0) { --sp; const StackEntry<ANY> e = stack[sp]; top_code = e.code(); top_t = e.t(); }
                if (saved_left > 0) { pend = -1; tri_i = saved_i; tri_left = saved_left; cur = kHold; }  // pend: any leaf code, tri_i / tri_left carry the position
                else cur = kRetry;
            }
            if (fin) {
                if (ANY) ((uint8_t*)out)[ray_id] = h.prim != 0xffffffffu ? 1 : 0;
                else {
                    ((float4*)out)[ray_id] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1);
                    if (b2_out) b2_out[ray_id] = h.b2;
                    if (inst_out) inst_out[ray_id] = h_inst;
                }
            }
        }
    }
}

}  // namespace b2
#include "instancing_spec3.cuh"
namespace b2 {

int trace_work_counter(int device, const TraceLaunch* tl, cudaStream_t s, unsigned long long** out);  // traverse_kernels.cu

template <bool ANY>
static int launch_phased2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, float* d_b2, int* d_inst, int variant, const TraceLaunch* tl) {
    DevCtx* c = dev_ctx(A.top.device);
    if (!c) { b200pt_set_error("two-level traversal: accelerator on a device that was never initialised"); return B200PT_ERR_NO_DEVICE; }
    if (n >= 0x7fffffffLL) { b200pt_set_error("two-level traversal: at most 2^31-2 rays per launch"); return B200PT_ERR_INVALID; }
    unsigned long long* ctr = nullptr;
    int rc = trace_work_counter(A.top.device, tl, s, &ctr);
    if (rc) return rc;
    const int* n_dev = tl ? tl->n_dev : nullptr;
    int64_t want = (n + 127) / 128;
    if (variant == 4) {
        constexpr int kBlocks = 5;
        int grid = (int)std::min<int64_t>(want, (int64_t)c->sm_count * kBlocks);
        k_trace_phased2<ANY, 16, 16, kBlocks><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
    } else {
        // CTAs per SM (register cap): closest-hit 7 (72 registers, 62 B of spills: 3 % faster on C5 than 6 without spills), any-hit 7; B200PT_2L_BLOCKS = 5..8 overrides the closest-hit choice (A/B)
        static const int closest_blocks = [] { const char* e = std::getenv("B200PT_2L_BLOCKS"); int v = e ? std::atoi(e) : 7; return (v >= 5 && v <= 8) ? v : 7; }();
        const int kb = ANY ? 7 : closest_blocks;
        int grid = (int)std::min<int64_t>(want, (int64_t)c->sm_count * kb);
        // B200PT_2L_KERNEL: 3 (default) = three-phase kernel (instancing_spec3.cuh), 2 = k_trace_spec2_2l; B200PT_2L_TUNE picks the
        // phase-switch / refill thresholds of the three-phase kernel (A/B)
        static const int kernel = [] { const char* e = std::getenv("B200PT_2L_KERNEL"); return e ? std::atoi(e) : 2; }();
        static const int tune = [] { const char* e = std::getenv("B200PT_2L_TUNE"); return e ? std::atoi(e) : 4; }();  // profiles/r2_c5_ab.txt: (12, 8) is 5 % faster on C5 than (20, 20)
        if (kernel == 3 && kb == 7) {
            if (tune == 1) k_trace_spec3_2l<ANY, 16, 12, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
            else if (tune == 2) k_trace_spec3_2l<ANY, 12, 8, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
            else if (tune == 3) k_trace_spec3_2l<ANY, 24, 12, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
            else k_trace_spec3_2l<ANY, 20, ANY ? 16 : 20, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
            g_launches.fetch_add(1);
            cudaError_t e3 = cudaGetLastError();
            return e3 == cudaSuccess ? B200PT_OK : cuda_fail(e3, "k_trace_spec3_2l launch");
        }
        if (kb == 7 && tune == 1) k_trace_spec2_2l<ANY, 16, 12, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else if (kb == 7 && tune == 2) k_trace_spec2_2l<ANY, 20, 12, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else if (kb == 7 && tune == 3) k_trace_spec2_2l<ANY, 24, 16, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else if (kb == 7 && tune == 4) {  // default
            if (A.top.alpha) k_trace_spec2_2l<ANY, 12, 8, 7, true><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
            else k_trace_spec2_2l<ANY, 12, 8, 7, false><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        }
        else if (kb == 5) k_trace_spec2_2l<ANY, 20, ANY ? 16 : 20, 5><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else if (kb == 7) k_trace_spec2_2l<ANY, 20, ANY ? 16 : 20, 7><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else if (kb == 8) k_trace_spec2_2l<ANY, 20, ANY ? 16 : 20, 8><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
        else k_trace_spec2_2l<ANY, 20, ANY ? 16 : 20, 6><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, ctr, d_b2, d_inst, n_dev);
    }
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200PT_OK : cuda_fail(e, "k_trace_phased2 launch");
}

int launch_intersect2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, float* d_b2, int* d_inst, int variant, const TraceLaunch* tl) {
    if (n <= 0) return B200PT_OK;
    if (variant == 0 || variant == 4) return launch_phased2<false>(A, d_rays, n, d_hits, s, d_b2, d_inst, variant, tl);
    if (tl && tl->n_dev) { b200pt_set_error("two-level traversal: device-resident ray counts need a persistent kernel variant"); return B200PT_ERR_INVALID; }
    k_trace_twolevel<false><<<(int)((n + 127) / 128), 128, 0, s>>>(A, (const float4*)d_rays, n, d_hits, d_b2, d_inst);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200PT_OK : cuda_fail(e, "k_trace_twolevel launch");
}
int launch_occluded2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant, const TraceLaunch* tl) {
    if (n <= 0) return B200PT_OK;
    if (variant == 0 || variant == 4) return launch_phased2<true>(A, d_rays, n, d_out, s, nullptr, nullptr, variant, tl);
    if (tl && tl->n_dev) { b200pt_set_error("two-level traversal: device-resident ray counts need a persistent kernel variant"); return B200PT_ERR_INVALID; }
    k_trace_twolevel<true><<<(int)((n + 127) / 128), 128, 0, s>>>(A, (const float4*)d_rays, n, d_out, nullptr, nullptr);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200PT_OK : cuda_fail(e, "k_trace_twolevel launch");
}

// Appends one BVH (wide nodes + leaf records) to the shared arrays; returns its root code.
template <class RecordFn>
static int append_bvh(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered, int64_t n_ordered, RecordFn record,
                      std::vector<float4>& wide, std::vector<float4>& recs, int* root_code) {
    const int64_t wbase = (int64_t)wide.size() / 4, rbase = (int64_t)recs.size() / 4;
    std::vector<int32_t> wide_of((size_t)n_nodes, -1);
    int64_t n_wide = 0;
    for (int64_t i = 0; i < n_nodes; ++i)
        if (nodes[i].n_primitives == 0) wide_of[(size_t)i] = (int32_t)(wbase + n_wide++);
    auto code_of = [&](int64_t i) -> int32_t { return nodes[i].n_primitives == 0 ? wide_of[(size_t)i] : ~(int32_t)(rbase + nodes[i].offset); };
    wide.resize((size_t)(wbase + n_wide) * 4);
    for (int64_t i = 0; i < n_nodes; ++i) {
        if (nodes[i].n_primitives != 0) continue;
        int64_t c0 = i + 1, c1 = nodes[i].offset;
        if (c1 <= i || c1 >= n_nodes || c0 >= n_nodes) { b200pt_set_error("b200pt_scene_create: malformed node array"); return B200PT_ERR_INVALID; }
        const float* a0 = nodes[c0].bounds;
        const float* a1 = nodes[c1].bounds;
        float4* q = &wide[(size_t)wide_of[(size_t)i] * 4];
        q[0] = make_float4(a0[0], a0[1], a0[2], a0[3]);
        q[1] = make_float4(a0[4], a0[5], a1[0], a1[1]);
        q[2] = make_float4(a1[2], a1[3], a1[4], a1[5]);
        int32_t k0 = code_of(c0), k1 = code_of(c1), ax = nodes[i].axis;
        float f0, f1, f2;
        std::memcpy(&f0, &k0, 4); std::memcpy(&f1, &k1, 4); std::memcpy(&f2, &ax, 4);
        q[3] = make_float4(f0, f1, f2, 0.0f);
    }
    recs.resize((size_t)(rbase + n_ordered) * 4);
    for (int64_t j = 0; j < n_ordered; ++j) {
        int rc = record(ordered[j], &recs[(size_t)(rbase + j) * 4]);
        if (rc) return rc;
    }
    for (int64_t i = 0; i < n_nodes; ++i) {
        if (nodes[i].n_primitives == 0) continue;
        uint32_t cnt = nodes[i].n_primitives;
        if ((int64_t)nodes[i].offset + cnt > n_ordered) { b200pt_set_error("b200pt_scene_create: leaf range out of bounds"); return B200PT_ERR_INVALID; }
        float fc;
        std::memcpy(&fc, &cnt, 4);
        recs[(size_t)(rbase + nodes[i].offset) * 4 + 2].w = fc;
    }
    *root_code = n_nodes > 0 ? code_of(0) : B2_EMPTY_ROOT;
    return B200PT_OK;
}

int accel2_build_device(const b200pt_scene_desc* d, Accel2Impl* a) {
    {   // the scene aggregate's and the deepest object's pending entries share one stack of B2_STACK2 entries
        const int top = bvh_max_depth(d->nodes, d->n_nodes);
        int obj = 0;
        for (int o = 0; o < d->n_objects; ++o) {
            const int dd = bvh_max_depth(d->objects[o].nodes, d->objects[o].n_nodes);
            if (dd < 0) { b200pt_set_error("b200pt_scene_create: malformed object node array"); return B200PT_ERR_INVALID; }
            obj = std::max(obj, dd);
        }
        if (top < 0) { b200pt_set_error("b200pt_scene_create: malformed node array"); return B200PT_ERR_INVALID; }
        if (top > 64 || obj > 64 || top + obj + 2 > B2_STACK2) {
            b200pt_set_error("b200pt_scene_create: scene aggregate + object BVH are deeper than the traversal stack holds (64 levels each in the reference, mod.rs:185; 94 together here)");
            return B200PT_ERR_UNSUPPORTED;
        }
    }
    std::vector<float4> wide, recs;
    auto tri_record = [&](int64_t gp, float4* q) {
        const float* v = d->tri_verts + 9 * (size_t)gp;
        uint32_t p = (uint32_t)gp, fl = d->prim_flags ? d->prim_flags[gp] & 0x7fffffffu : 0u, zero = 0u;
        float fp, ff, fz;
        std::memcpy(&fp, &p, 4); std::memcpy(&ff, &fl, 4); std::memcpy(&fz, &zero, 4);
        q[0] = make_float4(v[0], v[1], v[2], v[3]);
        q[1] = make_float4(v[4], v[5], v[6], v[7]);
        q[2] = make_float4(v[8], fp, ff, fz);
        q[3] = record_duv(d->tri_uvs && (fl & B200PT_PRIM_HAS_UV) ? d->tri_uvs + 6 * (size_t)gp : nullptr);
    };
    const int64_t n_top = d->n_top_tris + d->n_instances;
    std::memset(&a->dev, 0, sizeof(a->dev));
    int root = B2_EMPTY_ROOT;
    int rc = append_bvh(d->nodes, d->n_nodes, d->ordered_prims, n_top, [&](uint32_t p, float4* q) -> int {
        if ((int64_t)p < d->n_top_tris) { tri_record(p, q); return 0; }
        int64_t inst = (int64_t)p - d->n_top_tris;
        if (inst >= d->n_instances) { b200pt_set_error("b200pt_scene_create: top-level primitive index out of range"); return B200PT_ERR_INVALID; }
        uint32_t id = (uint32_t)inst, fl = 0x80000000u, zero = 0u;
        float fp, ff, fz;
        std::memcpy(&fp, &id, 4); std::memcpy(&ff, &fl, 4); std::memcpy(&fz, &zero, 4);
        q[0] = q[1] = q[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        q[2] = make_float4(0.0f, fp, ff, fz);
        return 0;
    }, wide, recs, &root);
    if (rc) return rc;
    a->dev.top.root_code = root;
    a->dev.top.device = current_device();
    if (d->n_nodes > 0) std::memcpy(a->dev.top.root_bounds, d->nodes[0].bounds, 24);
    a->dev.top.n_nodes = (int)d->n_nodes;
    a->dev.top.n_prims = d->n_prims;
    struct ObjRoot { float bounds[6]; int root_code; };
    std::vector<ObjRoot> objs((size_t)d->n_objects);
    for (int o = 0; o < d->n_objects; ++o) {
        const b200pt_object& ob = d->objects[o];
        if (ob.n_nodes < 1 || ob.first_prim < d->n_top_tris || ob.first_prim + ob.n_prims > d->n_prims) {
            b200pt_set_error("b200pt_scene_create: object without a BVH or with a primitive range outside [n_top_tris, n_prims)");
            return B200PT_ERR_INVALID;
        }
        int oroot;
        rc = append_bvh(ob.nodes, ob.n_nodes, ob.ordered_prims, ob.n_prims, [&](uint32_t p, float4* q) -> int {
            if ((int64_t)p >= ob.n_prims) { b200pt_set_error("b200pt_scene_create: object ordered_prims index out of range"); return B200PT_ERR_INVALID; }
            tri_record(ob.first_prim + p, q);
            return 0;
        }, wide, recs, &oroot);
        if (rc) return rc;
        objs[(size_t)o].root_code = oroot;
        std::memcpy(objs[(size_t)o].bounds, ob.nodes[0].bounds, 24);
    }
    std::vector<DInstance> insts((size_t)d->n_instances);
    std::vector<float4> trav((size_t)d->n_instances * 6);
    for (int i = 0; i < d->n_instances; ++i) {
        const b200pt_instance& in = d->instances[i];
        const float* w = in.world_to_instance;
        const float* m = in.instance_to_world;
        if (in.object < 0 || in.object >= d->n_objects) {
            b200pt_set_error("b200pt_scene_create: instance with a bad object index");
            return B200PT_ERR_INVALID;
        }
        DInstance& I = insts[(size_t)i];
        std::memcpy(I.w2i, w, 64);
        std::memcpy(I.i2w, m, 64);
        I.object = in.object;
        bool ident = true;
        for (int k = 0; k < 16; ++k) if (m[k] != ((k % 5 == 0) ? 1.0f : 0.0f)) ident = false;
        I.identity = ident ? 1 : 0;  // Transform::is_identity, transformed_primitive.rs:57
        I.pad[0] = I.pad[1] = 0;
        float4* T = &trav[(size_t)i * 6];
        for (int k = 0; k < 4; ++k) T[k] = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        const ObjRoot& ob = objs[(size_t)in.object];
        float frc, fob;
        std::memcpy(&frc, &ob.root_code, 4); std::memcpy(&fob, &in.object, 4);
        T[4] = make_float4(ob.bounds[0], ob.bounds[1], ob.bounds[2], ob.bounds[3]);
        T[5] = make_float4(ob.bounds[4], ob.bounds[5], frc, fob);
    }
    B2_CUDA(cudaMalloc(&a->d_wide, std::max<size_t>(wide.size(), 4) * sizeof(float4)));
    B2_CUDA(cudaMalloc(&a->d_recs, std::max<size_t>(recs.size(), 4) * sizeof(float4)));
    B2_CUDA(cudaMalloc(&a->d_trav, std::max<size_t>(trav.size(), 6) * sizeof(float4)));
    B2_CUDA(cudaMalloc(&a->d_insts, std::max<size_t>(insts.size(), 1) * sizeof(DInstance)));
    if (!wide.empty()) B2_CUDA(cudaMemcpy(a->d_wide, wide.data(), wide.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (!recs.empty()) B2_CUDA(cudaMemcpy(a->d_recs, recs.data(), recs.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (!trav.empty()) B2_CUDA(cudaMemcpy(a->d_trav, trav.data(), trav.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (!insts.empty()) B2_CUDA(cudaMemcpy(a->d_insts, insts.data(), insts.size() * sizeof(DInstance), cudaMemcpyHostToDevice));
    a->dev.top.wide = a->d_wide;
    a->dev.top.tris = a->d_recs;
    a->dev.top.ref_nodes = nullptr;
    a->dev.inst_trav = a->d_trav;
    a->dev.instances = a->d_insts;
    return B200PT_OK;
}

void accel2_free_device(Accel2Impl* a) {
    for (void* p : {(void*)a->d_wide, (void*)a->d_recs, (void*)a->d_trav, (void*)a->d_insts})
        if (p) cudaFree(p);
    a->d_wide = a->d_recs = a->d_trav = nullptr; a->d_insts = nullptr;
}

}  // namespace b2
