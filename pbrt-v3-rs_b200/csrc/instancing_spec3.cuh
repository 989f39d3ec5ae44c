// Three-phase form of the two-level walk (default closest-hit / any-hit kernel of instanced scenes).
//
// profiles/r2_ncu_c5_*: in k_trace_spec2_2l (instancing.cu) the TransformedPrimitive entry (transform_ray, three
// divisions, root box test, triangle context: ~300 instructions) and the return to the scene aggregate (world ray
// re-derived: ~120) sit inside the TRI / pop paths and ran with 1.0 and 2.0 of 32 lanes on C5 - 20 % of all issued
// instructions.  Here a lane that reaches an instance record, or whose object walk has finished, only posts a crossing
// request (xreq) and waits; the warp votes among NODE, TRI and CROSS phases, and one CROSS phase serves every waiting
// lane with the shared part (ray load, reciprocals, triangle context) non-divergent.  The order in which one ray's
// nodes, leaves and instances are visited is unchanged (a lane does nothing between posting the request and the
// crossing), so results stay bit-identical (TransformedPrimitive::intersect, transformed_primitive.rs:43-73).
#pragma once

namespace b2 {

template <bool ANY, int kSwitch, int kRefill, int kBlocks>
__global__ void __launch_bounds__(128, kBlocks) k_trace_spec3_2l(DeviceAccel2 A2, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                                 unsigned long long* __restrict__ counter, float* __restrict__ b2_out, int* __restrict__ inst_out,
                                                                 const int* __restrict__ n_dev) {
    const DeviceAccel& A = A2.top;
    const unsigned lane = threadIdx.x & 31u;
    const int kIdle = (int)0x80000000;
    const int kRetry = (int)0x80000001;  // pop (again) in the next NODE step
    const int kHold = (int)0x80000002;   // scene-aggregate level: wait until the parked leaf has been processed, then pop
    const int kNoX = -1, kLeave = -2;    // xreq: none / leave the object / (>= 0) enter this instance record
    StackEntry<ANY> stack[B2_STACK2];

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
    int xreq = kNoX;
    bool inst_hit = false;
    HitOut h;
    int h_inst = -1;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = h.b1 = h.b2 = 0.0f;
    bool exhausted = false;
    int phase = 0;  // 0 NODE, 1 TRI, 2 CROSS

    for (;;) {
        const unsigned idle_mask = __ballot_sync(0xffffffffu, cur == kIdle && pend == kIdle && xreq == kNoX);
        if (idle_mask == 0xffffffffu && exhausted) break;
        if (!exhausted && __popc(idle_mask) >= kRefill) {
            const int want = __popc(idle_mask);
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(counter, (unsigned long long)want);
            b = __shfl_sync(0xffffffffu, b, 0);
            const long long n_rays = ray_count(n, n_dev);
            if ((long long)b + want >= n_rays) exhausted = true;
            if (cur == kIdle && pend == kIdle && xreq == kNoX) {
                const long long id = (long long)b + __popc(idle_mask & ((1u << lane) - 1u));
                if (id < n_rays) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = (int)id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
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
            const bool waits = xreq != kNoX;
            const unsigned m_x = __ballot_sync(0xffffffffu, waits);
            const unsigned m_node = __ballot_sync(0xffffffffu, !waits && (cur >= 0 || cur == kRetry));
            const unsigned m_tri = __ballot_sync(0xffffffffu, !waits && pend != kIdle);
            if (!(m_node | m_tri | m_x)) break;
            if (!exhausted && __popc(~(m_node | m_tri | m_x)) >= kRefill) break;
            const int nn = __popc(m_node), nt = __popc(m_tri), nx = __popc(m_x);
            const int n_cur = phase == 0 ? nn : (phase == 1 ? nt : nx);
            if (n_cur < kSwitch) phase = (nx > nn && nx >= nt) ? 2 : (nt > nn ? 1 : 0);
            if (phase == 0 && nn == 0) phase = nt >= nx ? 1 : 2;

            bool fin = false;        // this level's walk is finished (cur and pend both empty)
            bool leaf_done = false;  // the parked leaf has been processed completely
            if (phase == 0) {
                if (!waits) {
                    bool need_pop = cur == kRetry;
                    if (cur >= 0) {
                        const float4* q = A.wide + 4ll * cur;
                        float4 q0, q1, q2, q3;
                        ldg8(q, &q0, &q1);
                        ldg8(q + 2, &q2, &q3);
                        float t0, t1;
                        // literal box test: an instance-space ray may be axis-parallel where the world ray is not
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
                }
            } else if (phase == 1) {
                if (!waits && pend != kIdle) {
                    V3 p0, p1, p2;
                    uint32_t prim, flags, leaf_n;
                    if (tri_left == 0) tri_i = ~pend;
                    load_tri(A.tris, (long long)tri_i, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                    if (tri_left == 0) tri_left = leaf_n;
                    ++tri_i;
                    --tri_left;
                    if (flags & 0x80000000u) {
                        xreq = (int)prim;  // TransformedPrimitive (only in leaves of the scene aggregate, where cur == kHold)
                    } else {
                        float t, b0, b1, b2;
                        if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)tri_i - 1)) {
                            if (ANY) {
                                if (alpha_ok_any(A, flags, (long long)tri_i - 1, o, tc, t_max)) { h.prim = 0u; cur = kIdle; top_code = kIdle; sp = 0; sp_base = 0; in_inst = -1; tri_left = 0; }
                            } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                                t_max = t;
                                h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
                                h_inst = in_inst;
                                inst_hit = true;
                            }
                        }
                        leaf_done = tri_left == 0;
                    }
                }
            } else if (waits) {
                // CROSS: enter an instance or return to the scene aggregate
                const bool entering = xreq >= 0;
                const float4 w0 = __ldg(rays + 2ll * ray_id), w1 = __ldg(rays + 2ll * ray_id + 1);
                float ox = w0.x, oy = w0.y, oz = w0.z, dx = w1.x, dy = w1.y, dz = w1.z, tm = t_max;
                float4 b0q = make_float4(0.0f, 0.0f, 0.0f, 0.0f), b1q = b0q;
                if (entering) {
                    const float4* T = A2.inst_trav + 6ll * xreq;
                    const float4 m0 = __ldg(T), m1 = __ldg(T + 1), m2 = __ldg(T + 2), m3 = __ldg(T + 3);
                    b0q = __ldg(T + 4); b1q = __ldg(T + 5);
                    Ray32 wr{w0.x, w0.y, w0.z, t_max, w1.x, w1.y, w1.z, w1.w};
                    const Ray32 ir = xf_ray(m0, m1, m2, m3, wr);
                    ox = ir.ox; oy = ir.oy; oz = ir.oz; dx = ir.dx; dy = ir.dy; dz = ir.dz; tm = ir.tmax;
                }
                RayCtx rn;
                rn.ox = ox; rn.oy = oy; rn.oz = oz;
                rn.ix = 1.0f / dx; rn.iy = 1.0f / dy; rn.iz = 1.0f / dz;
                rn.nx = rn.ix < 0.0f; rn.ny = rn.iy < 0.0f; rn.nz = rn.iz < 0.0f;
                bool go = true;
                float te = 0.0f;
                const int root = __float_as_int(b1q.z);
                if (entering) go = root != B2_EMPTY_ROOT && slab(rn, b0q.x, b0q.y, b0q.z, b0q.w, b1q.x, b1q.y, &te) && te < tm;
                if (go) {
                    r = rn;
                    negmask = r.nx | (r.ny << 1) | (r.nz << 2);
                    tc = make_tri_ctx(dx, dy, dz);
                    o = mk(ox, oy, oz);
                }
                if (entering) {
                    if (go) {
                        saved_i = tri_i; saved_left = tri_left;
                        world_t_max = t_max;
                        in_inst = xreq; inst_hit = false;
                        if (top_code != kIdle) { stack[sp].set(top_code, top_t); ++sp; top_code = kIdle; }
                        sp_base = sp;
                        t_max = tm;
                        tri_left = 0;
                        if (root >= 0) { cur = root; cur_t = te; pend = kIdle; }
                        else { pend = root; cur = kRetry; }  // single-leaf object
                    } else {
                        leaf_done = tri_left == 0;
                    }
                } else {
                    // the object's walk is finished: back to the interrupted leaf of the scene aggregate
                    if (!inst_hit) t_max = world_t_max;
                    in_inst = -1;
                    sp_base = 0;
                    if (sp > 0) { --sp; const StackEntry<ANY> e = stack[sp]; top_code = e.code(); top_t = e.t(); }
                    if (saved_left > 0) { pend = -1; tri_i = saved_i; tri_left = saved_left; cur = kHold; }  // pend: any leaf code, tri_i / tri_left carry the position
                    else cur = kRetry;
                }
                xreq = kNoX;
            }
            if (leaf_done) {
                // leaf done, t_max current again: re-validate what was reached speculatively (object level only)
                pend = kIdle;
                const bool live = cur != kIdle && cur != kRetry && cur != kHold;
                if (cur == kHold) cur = kRetry;
                else if (!ANY && live && !(cur_t < t_max)) cur = kRetry;
                else if (live && cur < 0) { pend = cur; cur = kRetry; }
                fin = cur == kIdle;
            }
            if (fin && in_inst >= 0) { fin = false; xreq = kLeave; }
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
