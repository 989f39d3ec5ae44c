// Traversal kernels (K2a closest-hit, K2b any-hit; K3 triangle test inlined).
// See traverse.cuh for the data layout and the equivalence argument.
#include "common.cuh"

namespace b2 {

// variant 2: wide nodes, one thread per ray.
// variant 1: reference LinearBVHNode walk, one thread per ray (baseline).
template <bool ANY, int VARIANT>
__global__ void __launch_bounds__(128) k_trace_simple(DeviceAccel A, const float4* __restrict__ rays, long long n, void* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r0 = __ldg(rays + 2 * i), r1 = __ldg(rays + 2 * i + 1);
    Ray32 ray{r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    HitOut h;
    bool hit = (VARIANT == 1) ? traverse_ref<ANY>(A, ray, &h) : traverse_wide<ANY>(A, ray, &h);
    if (ANY) {
        ((uint8_t*)out)[i] = hit ? 1 : 0;
    } else {
        float4 o;
        o.x = h.t; o.y = __uint_as_float(h.prim); o.z = h.b0; o.w = h.b1;
        ((float4*)out)[i] = o;
    }
}

// variant 0 (default): persistent warps with dynamic ray fetch.
// Incoherent rays finish after very different numbers of steps, so in the
// one-thread-per-ray kernel a warp runs until its slowest ray is done with most
// lanes idle.  Here each warp keeps pulling work: whenever the number of lanes
// that still hold a live ray drops below a threshold, the idle lanes grab new
// ray indices from a global counter (one atomicAdd per warp, distributed with
// ballot/popc prefix) and the whole warp re-enters the traversal loop.  The
// grid is sized to the machine (SMs x resident CTAs), not to the ray count.
struct TravState {
    RayCtx r;
    TriCtx tc;
    float t_max;
    int cur;
    int sp;
    bool hit;
    long long ray_id;
    HitOut h;
};

template <bool ANY>
__global__ void __launch_bounds__(128, 4) k_trace_persistent(DeviceAccel A, const float4* __restrict__ rays, long long n,
                                                              void* __restrict__ out, unsigned long long* __restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    int stack_code[B2_STACK];
    float stack_t[B2_STACK];
    const int kDone = B2_EMPTY_ROOT;  // sentinel for "lane has no ray"

    long long ray_id = -1;
    RayCtx r;
    TriCtx tc;
    V3 o;
    float t_max = 0.0f;
    int cur = kDone, sp = 0;
    bool hit = false;
    HitOut h;
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = 0.0f; h.b1 = 0.0f;
    bool exhausted = false;  // warp-uniform: the global queue is empty

    for (;;) {
        // ---- refill idle lanes ------------------------------------------------
        const bool idle = (cur == kDone);
        const unsigned idle_mask = __ballot_sync(0xffffffffu, idle);
        if (idle_mask == 0xffffffffu && exhausted) break;
        if (!exhausted && __popc(idle_mask) >= 8) {  // refill when >= 25% of the warp is idle
            const int n_idle = __popc(idle_mask);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)n_idle);
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((long long)base + n_idle >= n) exhausted = true;
            if (idle) {
                const int my = __popc(idle_mask & ((1u << lane) - 1u));
                const long long id = (long long)base + my;
                if (id < n) {
                    float4 r0 = __ldg(rays + 2 * id), r1 = __ldg(rays + 2 * id + 1);
                    ray_id = id;
                    r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
                    r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
                    r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
                    t_max = r0.w;
                    tc = make_tri_ctx(r1.x, r1.y, r1.z);
                    o = mk(r0.x, r0.y, r0.z);
                    sp = 0;
                    hit = false;
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = 0.0f; h.b1 = 0.0f;
                    cur = A.root_code;
                    float te;
                    if (cur == kDone ||
                        !(slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) &&
                          te < t_max)) {
                        // miss at the root: retire immediately
                        if (ANY) ((uint8_t*)out)[id] = 0;
                        else { float4 ov; ov.x = h.t; ov.y = __uint_as_float(h.prim); ov.z = 0.0f; ov.w = 0.0f; ((float4*)out)[id] = ov; }
                        cur = kDone;
                    }
                }
            }
        }
        // ---- traverse until enough lanes went idle ------------------------------
        while (true) {
            if (cur != kDone) {
                bool finished = false;
                if (cur >= 0) {
                    const float4* q = A.wide + 4ll * cur;
                    float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2), q3 = ldg4(q + 3);
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
                        cur = kDone;  // pop below
                        finished = true;
                    }
                } else {
                    long long first = (long long)(~cur);
                    V3 p0, p1, p2;
                    uint32_t prim, flags, leaf_n;
                    load_tri(A.tris, first, &p0, &p1, &p2, &prim, &flags, &leaf_n);
                    bool any_done = false;
                    for (uint32_t i = 0;;) {
                        float t, b0, b1, b2;
                        if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2)) {
                            if (ANY) {
                                if (!(flags & 6u)) { any_done = true; break; }
                            } else if (!(flags & 2u)) {
                                hit = true;
                                t_max = t;
                                h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1;
                            }
                        }
                        if (++i >= leaf_n) break;
                        uint32_t dummy;
                        load_tri(A.tris, first + i, &p0, &p1, &p2, &prim, &flags, &dummy);
                    }
                    if (ANY && any_done) { hit = true; sp = 0; }
                    cur = kDone;
                    finished = true;
                }
                if (finished) {
                    // pop the next live entry
                    for (;;) {
                        if (sp == 0) { cur = kDone; break; }
                        --sp;
                        if (ANY || stack_t[sp] < t_max) { cur = stack_code[sp]; break; }
                    }
                    if (cur == kDone) {  // ray retired: write result
                        if (ANY) ((uint8_t*)out)[ray_id] = hit ? 1 : 0;
                        else { float4 ov; ov.x = h.t; ov.y = __uint_as_float(h.prim); ov.z = h.b0; ov.w = h.b1; ((float4*)out)[ray_id] = ov; }
                    }
                }
            }
            const unsigned live = __ballot_sync(0xffffffffu, cur != kDone);
            if (live == 0u) break;
            if (!exhausted && __popc(live) <= 24) break;  // go refill
        }
    }
}

static int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

int launch_intersect(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, int variant) {
    if (n <= 0) return B200PT_OK;
    const int block = 128;
    if (variant == 1) k_trace_simple<false, 1><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_hits);
    else k_trace_simple<false, 2><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_hits);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_trace launch");
    return B200PT_OK;
}

int launch_occluded(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant) {
    if (n <= 0) return B200PT_OK;
    const int block = 128;
    if (variant == 1) k_trace_simple<true, 1><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_out);
    else k_trace_simple<true, 2><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_out);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_trace launch");
    return B200PT_OK;
}

}  // namespace b2
