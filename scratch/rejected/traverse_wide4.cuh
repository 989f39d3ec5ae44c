// Any-hit walk over 4-WIDE nodes (A/B variant 6, shadow rays only).
//
// BVHAccel::intersect_p answers "is anything hit before t_max"; the answer does not depend on the order in which
// leaves are visited, only on WHICH leaves are tested.  A leaf is tested by the reference iff the boxes of all its
// ancestors pass Bounds3::intersect_p_inv.  A child's box lies inside its parent's and f32 subtraction / multiplication
// are monotonic, so a box that passes implies that every enclosing box passes: testing only every
// second level of the same SAH tree reaches exactly the same leaves.  Each 128-byte node therefore holds the boxes of the
// (up to four) grandchildren of a reference node - a leaf child stands for itself - and one step does four box tests
// behind ONE dependent fetch instead of two.  Closest-hit rays keep the two-wide walk: there the visiting order decides
// equal-t ties.
//
// Node record (8 float4): q0 = lo0.xyz hi0.x | q1 = hi0.yz lo1.xy | q2 = lo1.z hi1.xyz | q3 = lo2.xyz hi2.x |
// q4 = hi2.yz lo3.xy | q5 = lo3.z hi3.xyz | q6 = child codes (>= 0 wide node, < 0 ~first triangle, kIdle = none) | q7 = -.
// Scheduling is k_trace_spec2's: persistent warps, NODE / TRI phase votes, a parked leaf while the walk goes on, the
// newest stack entry in registers.
#pragma once
#include "traverse_spec.cuh"

namespace b2 {

#define B2_STACK_W4 128

template <int kSwitch, int kRefill, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks) k_occl_wide4(DeviceAccel A, const float4* __restrict__ rays, long long n, uint8_t* __restrict__ out,
                                                             unsigned long long* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    const int kRetry = (int)0x80000001;
    int stack[B2_STACK_W4];

    int ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f;
    int cur = kIdle;
    int pend = kIdle;
    int sp = 0;
    int top_code = kIdle;
    int tri_i = 0;
    uint32_t tri_left = 0;
    bool hit_any = false;
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
            if ((long long)b + want >= n) exhausted = true;
            if (cur == kIdle && pend == kIdle) {
                lane_slow = false;
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = (int)id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    t_max = r0.w;
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
                    sp = 0; tri_left = 0; top_code = kIdle; hit_any = false;
                    float te;
                    bool enter = A.root4_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) {
                        if (A.root4_code >= 0) cur = A.root4_code;
                        else pend = A.root4_code;  // single-leaf tree
                        lane_slow = !slab_fast_ok(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz);
                    } else {
                        out[id] = 0;
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
                bool need_pop = cur == kRetry;
                if (cur >= 0) {
                    const float4* q = A.wide4 + 8ll * cur;
                    float4 q0, q1, q2, q3, q4, q5, q6, q7;
                    ldg8(q, &q0, &q1);
                    ldg8(q + 2, &q2, &q3);
                    ldg8(q + 4, &q4, &q5);
                    ldg8(q + 6, &q6, &q7);
                    float t0, t1, t2, t3;
                    bool h0, h1, h2, h3;
                    if (!warp_slow) {
                        h0 = slab_fast(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_fast(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                        h2 = slab_fast(r, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, &t2) & (t2 < t_max);
                        h3 = slab_fast(r, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, &t3) & (t3 < t_max);
                    } else {
                        h0 = slab_bf(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_bf(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                        h2 = slab_bf(r, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, &t2) & (t2 < t_max);
                        h3 = slab_bf(r, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, &t3) & (t3 < t_max);
                    }
                    const int c0 = __float_as_int(q6.x), c1 = __float_as_int(q6.y), c2 = __float_as_int(q6.z), c3 = __float_as_int(q6.w);
                    h0 &= c0 != kIdle; h1 &= c1 != kIdle; h2 &= c2 != kIdle; h3 &= c3 != kIdle;
                    // the first hit child is walked next, the others are stacked (any order is correct for any-hit)
                    int next = kIdle;
                    auto take = [&](bool h, int c) {
                        if (h) {
                            if (next == kIdle) next = c;
                            else {
                                if (top_code != kIdle) { stack[sp] = top_code; ++sp; }
                                top_code = c;
                            }
                        }
                    };
                    take(h0, c0); take(h1, c1); take(h2, c2); take(h3, c3);
                    cur = next;
                    need_pop = next == kIdle;
                    const bool park = !need_pop & (cur < 0) & (pend == kIdle);
                    pend = park ? cur : pend;
                    tri_left = park ? 0u : tri_left;
                    need_pop |= park;
                }
                if (need_pop) {
                    const int c = top_code;
                    const bool have = c != kIdle;
                    const bool refill = have & (sp > 0);
                    sp -= refill ? 1 : 0;
                    top_code = refill ? stack[sp] : kIdle;
                    cur = have ? c : kIdle;
                    const bool park = have & (c < 0) & (pend == kIdle);
                    pend = park ? c : pend;
                    tri_left = park ? 0u : tri_left;
                    cur = park ? kRetry : cur;
                    done = (cur == kIdle) & (pend == kIdle);
                }
            } else if (pend != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = ~pend;
                load_tri(A.tris, (long long)tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                float t, b0, b1, b2;
                if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)tri_i)) {
                    if (!(flags & 6u)) { hit_any = true; cur = kIdle; top_code = kIdle; sp = 0; tri_left = 1; }
                }
                ++tri_i;
                if (--tri_left == 0) {
                    pend = kIdle;
                    const bool live = cur != kIdle && cur != kRetry;
                    if (live && cur < 0) { pend = cur; cur = kRetry; }  // the blocked leaf becomes the parked one
                    done = cur == kIdle;
                }
            }
            if (done) out[ray_id] = hit_any ? 1 : 0;
        }
    }
}

}  // namespace b2
