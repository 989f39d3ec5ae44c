// Phase-scheduled persistent traversal with a POSTPONED LEAF (A/B variant 9).
//
// In k_trace_phased a lane that reaches a leaf waits for the warp's next TRI phase, and the NODE phase runs with
// ~17 of 32 lanes on incoherent rays.  Here a lane parks the leaf in `pend` and keeps walking (pops the next stack
// entry) until it reaches a second leaf; the TRI phase then serves every lane that carries a parked leaf.
//
// Results stay bit-identical to the reference walk:
//   * A leaf is parked only when no other leaf is parked, i.e. when every triangle the reference would have tested
//     before it HAS been tested and ray.t_max is current: the leaf was reached exactly as in the reference.
//   * Everything walked while a leaf is parked uses a t_max that may be too LARGE.  The only t_max-dependent term of
//     Bounds3::intersect_p_inv is `t_entry < ray.t_max`; all other comparisons and all arithmetic are unchanged, so the
//     speculative walk visits a superset of the reference's nodes in the reference's order.
//   * A node reached speculatively carries its entry distance (cur_t, or the stacked t of far children).  When the
//     parked leaf has been tested (t_max current again) the blocked node / every later pop re-applies
//     `t_entry < t_max`.  Child boxes lie inside their parent's box and f32 subtraction / multiplication are monotonic,
//     so t_entry(child) >= t_entry(parent): a node the reference would have culled with the current t_max fails that
//     re-check itself, together with everything below it.
//   * Leaves are tested in walk order (the blocked leaf waits for the parked one), so ties still go to the last
//     tested primitive.
#pragma once
#include "traverse_phased.cuh"

namespace b2 {

template <bool ANY, int kSwitch, int kRefill, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks) k_trace_spec(DeviceAccel A, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                             unsigned long long* __restrict__ counter, float* __restrict__ b2_out, const int* __restrict__ n_dev) {
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    StackEntry<ANY> stack[B2_STACK];

    int ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f;
    int cur = kIdle;           // next node: >= 0 interior, < 0 leaf blocked behind `pend`, kIdle: walk finished
    float cur_t = 0.0f;        // entry distance of `cur`
    int pend = kIdle;          // parked leaf (being tested in TRI phases)
    int sp = 0;
    int top_code = kIdle;
    float top_t = 0.0f;
    int negmask = 0;
    int tri_i = 0;
    uint32_t tri_left = 0;
    HitOut h;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;
    bool node_phase = true;
    bool lane_slow = false, warp_slow = false;

    // next stack entry whose entry distance is still in range (t_max may be stale-large: conservative)
    auto pop = [&]() {
        cur = kIdle;
        while (top_code != kIdle) {
            const int c = top_code;
            const float t = top_t;
            if (sp > 0) { --sp; const StackEntry<ANY> e = stack[sp]; top_code = e.code(); top_t = e.t(); }
            else top_code = kIdle;
            if (ANY || t < t_max) { cur = c; cur_t = t; break; }
        }
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
                lane_slow = false;
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n_rays) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = (int)id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                    t_max = r0.w;
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
                    sp = 0; tri_left = 0; top_code = kIdle;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
                    float te;
                    bool enter = A.root_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) {
                        if (A.root_code >= 0) { cur = A.root_code; cur_t = te; }
                        else pend = A.root_code;  // single-leaf tree
                        lane_slow = !slab_fast_ok(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz);
                    } else if (ANY) {
                        ((uint8_t*)out)[id] = 0;
                    } else {
                        ((float4*)out)[id] = make_float4(h.t, __uint_as_float(h.prim), 0.0f, 0.0f);
                        if (b2_out) b2_out[id] = 0.0f;
                    }
                }
            }
            warp_slow = __any_sync(0xffffffffu, lane_slow);
        }
        for (;;) {
            const unsigned m_node = __ballot_sync(0xffffffffu, cur >= 0);
            const unsigned m_tri = __ballot_sync(0xffffffffu, pend != kIdle);
            if (!(m_node | m_tri)) break;
            if (!exhausted && __popc(~(m_node | m_tri)) >= kRefill) break;
            const int nn = __popc(m_node), nt = __popc(m_tri);
            if (node_phase) { if (nn < kSwitch && nt > nn) node_phase = false; }
            else            { if (nt < kSwitch && nn > nt) node_phase = true; }
            if (nt == 0) node_phase = true;
            if (nn == 0) node_phase = false;

            bool done = false;
            if (node_phase) {
                if (cur >= 0) {
                    const float4* q = A.wide + 4ll * cur;
                    float4 q0, q1, q2, q3;
                    ldg8(q, &q0, &q1);
                    ldg8(q + 2, &q2, &q3);
                    float t0, t1;
                    bool h0, h1;
                    if (!warp_slow) {
                        h0 = slab_fast(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_fast(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                    } else {
                        h0 = slab_bf(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_bf(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                    }
                    const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y), axis = __float_as_int(q3.z);
                    const bool neg = (negmask >> axis) & 1;
                    const int near_c = neg ? c1 : c0, far_c = neg ? c0 : c1;
                    const bool near_h = neg ? h1 : h0, far_h = neg ? h0 : h1;
                    const float near_t = neg ? t1 : t0, far_t = neg ? t0 : t1;
                    bool need_pop = false;
                    if (near_h) {
                        if (far_h) {
                            if (top_code != kIdle) { stack[sp].set(top_code, top_t); ++sp; }
                            top_code = far_c; top_t = far_t;
                        }
                        cur = near_c; cur_t = near_t;
                    } else if (far_h) {
                        cur = far_c; cur_t = far_t;
                    } else {
                        need_pop = true;
                    }
                    // a leaf reached with no leaf parked was reached with the current t_max: park it and walk on
                    if (!need_pop && cur < 0 && pend == kIdle) { pend = cur; tri_left = 0; need_pop = true; }
                    if (need_pop) {
                        pop();
                        if (cur < 0 && cur != kIdle && pend == kIdle) { pend = cur; tri_left = 0; pop(); }  // popped straight into a leaf, nothing parked
                        done = cur == kIdle && pend == kIdle;
                    }
                }
            } else if (pend != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = ~pend;
                load_tri(A.tris, (long long)tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                float t, b0, b1, b2;
                if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)tri_i)) {
                    if (ANY) {
                        if (alpha_ok_any(A, flags, (long long)tri_i, o, tc, t_max)) { h.prim = 0u; cur = kIdle; top_code = kIdle; sp = 0; tri_left = 1; }  // occluded: drop the rest of the walk
                    } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                        t_max = t;
                        h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                    }
                }
                ++tri_i;
                if (--tri_left == 0) {
                    // t_max is current again: re-validate what was reached speculatively
                    pend = kIdle;
                    if (!ANY && cur != kIdle && !(cur_t < t_max)) pop();
                    if (cur < 0 && cur != kIdle) {
                        // the blocked leaf (validated against the current t_max) becomes the parked one
                        pend = cur;
                        pop();
                    }
                    done = cur == kIdle && pend == kIdle;
                }
            }
            if (done) {
                if (ANY) ((uint8_t*)out)[ray_id] = h.prim != 0xffffffffu ? 1 : 0;
                else {
                    ((float4*)out)[ray_id] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1);
                    if (b2_out) b2_out[ray_id] = h.b2;
                }
            }
        }
    }
}

// Loop-free form of the same walk (A/B variant 12): ONE predicated pop attempt per NODE step.  A lane whose popped
// entry fails `t_entry < t_max`, or that has just parked the leaf it popped, stays in state kRetry and pops again in
// the next NODE step instead of making the warp wait in a 2-3-lane loop with a dependent local-memory load
// (profiles/r1: those loops were ~16 % of the issued instructions).  kReps node steps run per phase vote.
template <bool ANY, int kSwitch, int kRefill, int kBlocks, int kReps, bool kTex = true>
__global__ void __launch_bounds__(128, kBlocks) k_trace_spec2(DeviceAccel A, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                              unsigned long long* __restrict__ counter, float* __restrict__ b2_out, const int* __restrict__ n_dev) {
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    const int kRetry = (int)0x80000001;  // pop (again) in the next NODE step; never a leaf code (~first with first < 2^31 - 2)
    StackEntry<ANY> stack[B2_STACK];

    int ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f;
    int cur = kIdle;
    float cur_t = 0.0f;
    int pend = kIdle;
    int sp = 0;
    int top_code = kIdle;
    float top_t = 0.0f;
    int negmask = 0;
    int tri_i = 0;
    uint32_t tri_left = 0;
    HitOut h;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;
    bool node_phase = true;
    bool lane_slow = false, warp_slow = false;

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
                lane_slow = false;
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n_rays) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = (int)id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                    t_max = r0.w;
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
                    sp = 0; tri_left = 0; top_code = kIdle;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
                    float te;
                    bool enter = A.root_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) {
                        if (A.root_code >= 0) { cur = A.root_code; cur_t = te; }
                        else pend = A.root_code;
                        lane_slow = !slab_fast_ok(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz);
                    } else if (ANY) {
                        ((uint8_t*)out)[id] = 0;
                    } else {
                        ((float4*)out)[id] = make_float4(h.t, __uint_as_float(h.prim), 0.0f, 0.0f);
                        if (b2_out) b2_out[id] = 0.0f;
                    }
                }
            }
            warp_slow = __any_sync(0xffffffffu, lane_slow);
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

            bool done = false;
            if (node_phase) {
#pragma unroll
                for (int rep = 0; rep < kReps; ++rep) {
                    bool need_pop = cur == kRetry;
                    if (cur >= 0) {
                        const float4* q = A.wide + 4ll * cur;
                        float4 q0, q1, q2, q3;
                        ldg8(q, &q0, &q1);
                        ldg8(q + 2, &q2, &q3);
                        float t0, t1;
                        bool h0, h1;
                        if (!warp_slow) {
                            h0 = slab_fast(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                            h1 = slab_fast(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                        } else {
                            h0 = slab_bf(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                            h1 = slab_bf(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                        }
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
                        // a leaf reached with no leaf parked was reached with the current t_max: park it and walk on
                        const bool park = !need_pop & (cur < 0) & (pend == kIdle);
                        pend = park ? cur : pend;
                        tri_left = park ? 0u : tri_left;
                        need_pop |= park;
                    }
                    if (need_pop) {
                        const int c = top_code;
                        const float t = top_t;
                        const bool have = c != kIdle;
                        const bool refill = have & (sp > 0);
                        sp -= refill ? 1 : 0;
                        StackEntry<ANY> e;
                        e.set(kIdle, 0.0f);
                        if (refill) e = stack[sp];
                        top_code = e.code(); top_t = e.t();
                        const bool valid = have & (ANY || t < t_max);
                        cur = valid ? c : (have ? kRetry : kIdle);
                        cur_t = t;
                        const bool park = valid & (c < 0) & (pend == kIdle);  // popped straight into a leaf, nothing parked
                        pend = park ? c : pend;
                        tri_left = park ? 0u : tri_left;
                        cur = park ? kRetry : cur;
                        done = (cur == kIdle) & (pend == kIdle);
                    }
                    if (done) break;
                }
            } else if (pend != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = ~pend;
                load_tri(A.tris, (long long)tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                float t, b0, b1, b2;
                if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)tri_i)) {
                    if (ANY) {
                        if (alpha_ok_any<kTex>(A, flags, (long long)tri_i, o, tc, t_max)) { h.prim = 0u; cur = kIdle; top_code = kIdle; sp = 0; tri_left = 1; }
                    } else if (alpha_ok<false, kTex>(A, flags, prim, b0, b1, b2)) {
                        t_max = t;
                        h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                    }
                }
                ++tri_i;
                if (--tri_left == 0) {
                    // t_max is current again: re-validate what was reached speculatively
                    pend = kIdle;
                    const bool live = cur != kIdle && cur != kRetry;
                    if (!ANY && live && !(cur_t < t_max)) cur = kRetry;
                    else if (live && cur < 0) { pend = cur; cur = kRetry; }  // the blocked leaf becomes the parked one
                    done = cur == kIdle;
                }
            }
            if (done) {
                if (ANY) ((uint8_t*)out)[ray_id] = h.prim != 0xffffffffu ? 1 : 0;
                else {
                    ((float4*)out)[ray_id] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1);
                    if (b2_out) b2_out[ray_id] = h.b2;
                }
            }
        }
    }
}

}  // namespace b2
