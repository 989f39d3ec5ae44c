// Phase-scheduled persistent traversal (variant 0, the default).
//
// The first persistent kernel (variant 0) lets every lane do "whatever it needs next" each iteration —
// a box-pair test or a leaf's triangle tests — so the warp executes both code paths almost every
// iteration with a handful of lanes in the triangle path: ncu shows 11.5 of 32 threads active per
// instruction and the kernel is issue-bound (profiles/r1_*).  Here the warp votes on a PHASE and only
// that code path runs:
//   NODE phase  — lanes sitting on an interior node test its two child boxes; lanes that reached a leaf wait.
//   TRI  phase  — lanes sitting on a leaf test ONE triangle; lanes on interior nodes wait.
// The phase flips when fewer than kSwitch lanes can still make progress in it, and idle lanes are refilled
// when at least kRefill of them are free.  Refills come from a per-warp chunk of kChunk consecutive rays
// reserved with one atomicAdd on the global counter, so rays that share a warp stay neighbours in the input
// (coherent primary rays keep hitting the same nodes / L1 lines); incoherent input is unaffected.
// Measured on C2 (profiles/r1_variants.txt): kSwitch 6 -> 16 raised the node-phase lane utilisation and is
// +15 % on incoherent closest-hit, +13 % on any-hit.  Per ray the visiting order, every
// comparison and every arithmetic operation is unchanged (see traverse.cuh), so results stay bit-identical.
#pragma once
#include "traverse.cuh"

namespace b2 {

template <bool ANY> struct StackEntry;
template <> struct StackEntry<false> {
    int2 v;
    B2_D void set(int code, float t) { v = make_int2(code, __float_as_int(t)); }
    B2_D int code() const { return v.x; }
    B2_D float t() const { return __int_as_float(v.y); }
};
template <> struct StackEntry<true> {
    int v;
    B2_D void set(int code, float) { v = code; }
    B2_D int code() const { return v; }
    B2_D float t() const { return 0.0f; }
};

// Stack: the newest entry lives in registers (top_code / top_t); a pop takes it from there and only PREFETCHES the
// entry below it from local memory, so the local-memory load latency is off the pop -> next-node-fetch critical path
// (the pop section runs with 3-4 active lanes, profiles/r1_*).  Entries are 8 bytes {code, bits(t_entry)}: one
// STL.64 / LDL.64 per spill / refill.  The box tests have no early-out branches (slab_bf).
// kOpt bit 0: warps in which every live ray has finite non-zero direction components take the min/max box test
//             (slab_fast, 25 vs 37 instructions per box); a warp holding an axis-parallel ray keeps the literal one.
template <bool ANY, int kSwitch, int kRefill, int kChunk, int kBlocks, int kOpt = 1>
__global__ void __launch_bounds__(128, kBlocks) k_trace_phased(DeviceAccel A, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                         unsigned long long* __restrict__ counter, float* __restrict__ b2_out, const int* __restrict__ n_dev) {
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;  // no ray in this lane (leaf codes are ~first >= -2^31 + 1)
    StackEntry<ANY> stack[B2_STACK];    // closest: {code, bits(t_entry)}; any-hit: code only (t_max never shrinks)

    long long ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f;
    int cur = kIdle, sp = 0;
    int top_code = kIdle;      // newest stack entry (kIdle = none)
    float top_t = 0.0f;
    int negmask = 0;           // dir_is_neg bits
    long long tri_i = 0;       // next triangle of the current leaf
    uint32_t tri_left = 0;     // triangles left in the current leaf (0 = leaf header not read yet)
    bool hit = false;
    HitOut h;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;  // warp-uniform: the global ray counter ran past n
    bool node_phase = true;
    bool lane_slow = false;  // this lane's ray needs the literal box test
    bool warp_slow = false;  // warp-uniform: some lane's ray does
    long long chunk_next = 0, chunk_end = 0;  // warp-uniform: this warp's reserved ray range

    for (;;) {
        // ---- refill ---------------------------------------------------------------------------
        const unsigned idle_mask = __ballot_sync(0xffffffffu, cur == kIdle);
        if (idle_mask == 0xffffffffu && exhausted) break;
        if (!exhausted && __popc(idle_mask) >= kRefill) {
            if (chunk_next >= chunk_end) {
                const int want = kChunk > 0 ? kChunk : __popc(idle_mask);  // kChunk == 0: take exactly the idle lanes' worth
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(counter, (unsigned long long)want);
                b = __shfl_sync(0xffffffffu, b, 0);
                chunk_next = (long long)b;
                const long long n_rays = ray_count(n, n_dev);
                chunk_end = (long long)b + want < n_rays ? (long long)b + want : n_rays;
                if (chunk_next >= n_rays) { exhausted = true; chunk_end = chunk_next; }
            }
            const long long base = chunk_next;
            const int rank = __popc(idle_mask & ((1u << lane) - 1u));
            const long long take = (chunk_end - chunk_next) < (long long)__popc(idle_mask) ? (chunk_end - chunk_next) : (long long)__popc(idle_mask);
            chunk_next += take;
            if (cur == kIdle) {
                lane_slow = false;
                const long long id = base + rank;
                if (rank < take) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                    t_max = r0.w;
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
                    sp = 0; hit = false; tri_left = 0; top_code = kIdle;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
                    float te;
                    bool enter = A.root_code != B2_EMPTY_ROOT &&
                                 slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) && te < t_max;
                    if (enter) {
                        cur = A.root_code;
                        if (kOpt & 1) lane_slow = !slab_fast_ok(r.ox, r.oy, r.oz, r.ix, r.iy, r.iz);
                    } else {
                        if (ANY) ((uint8_t*)out)[id] = 0;
                        else { ((float4*)out)[id] = make_float4(h.t, __uint_as_float(h.prim), 0.0f, 0.0f); if (b2_out) b2_out[id] = 0.0f; }
                    }
                }
            }
            if (kOpt & 1) warp_slow = __any_sync(0xffffffffu, lane_slow);
        }
        // ---- run phases until enough lanes went idle ------------------------------------------------
        for (;;) {
            const unsigned m_node = __ballot_sync(0xffffffffu, cur >= 0);
            const unsigned m_tri = __ballot_sync(0xffffffffu, cur < 0 && cur != kIdle);
            if (!(m_node | m_tri)) break;
            if (!exhausted && __popc(~(m_node | m_tri)) >= kRefill) break;  // go refill
            // phase vote with hysteresis
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
                    bool h0, h1;
                    if ((kOpt & 1) && !warp_slow) {
                        h0 = slab_fast(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_fast(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                    } else {
                        h0 = slab_bf(r, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &t0) & (t0 < t_max);
                        h1 = slab_bf(r, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &t1) & (t1 < t_max);
                    }
                    const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y), axis = __float_as_int(q3.z);
                    const bool neg = (negmask >> axis) & 1;
                    // reference order: neg ? (second first, push first) : (first first, push second)
                    const int near_c = neg ? c1 : c0, far_c = neg ? c0 : c1;
                    const bool near_h = neg ? h1 : h0, far_h = neg ? h0 : h1;
                    const float far_t = neg ? t0 : t1;
                    if (near_h) {
                        if (far_h) {
                            if (top_code != kIdle) { stack[sp].set(top_code, top_t); ++sp; }
                            top_code = far_c; top_t = far_t;
                        }
                        cur = near_c;
                    } else if (far_h) {
                        cur = far_c;
                    } else {
                        retire = true;  // pop
                    }
                    tri_left = 0;
                }
            } else if (cur < 0 && cur != kIdle) {
                V3 p0, p1, p2;
                uint32_t prim, flags, leaf_n;
                if (tri_left == 0) tri_i = (long long)(~cur);
                load_tri(A.tris, tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                if (tri_left == 0) tri_left = leaf_n;
                float t, b0, b1, b2;
                if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, tri_i)) {
                    if (ANY) {
                        if (alpha_ok_any(A, flags, tri_i, o, tc, t_max)) { hit = true; sp = 0; top_code = kIdle; tri_left = 1; }
                    } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                        hit = true;
                        t_max = t;
                        h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                    }
                }
                ++tri_i;
                if (--tri_left == 0) retire = true;  // leaf done: pop
            }
            if (retire) {
                cur = kIdle;
                while (top_code != kIdle) {
                    const int c = top_code;
                    const float t = top_t;
                    if (sp > 0) { --sp; const StackEntry<ANY> e = stack[sp]; top_code = e.code(); top_t = e.t(); }
                    else top_code = kIdle;
                    if (ANY || t < t_max) { cur = c; break; }
                }
                tri_left = 0;
                if (cur == kIdle) {
                    if (ANY) ((uint8_t*)out)[ray_id] = hit ? 1 : 0;
                    else { ((float4*)out)[ray_id] = make_float4(h.t, __uint_as_float(h.prim), h.b0, h.b1); if (b2_out) b2_out[ray_id] = h.b2; }
                }
            }
        }
    }
}

}  // namespace b2
