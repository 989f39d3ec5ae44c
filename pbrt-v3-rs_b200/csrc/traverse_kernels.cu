// Traversal kernels (K2a closest-hit, K2b any-hit; K3 triangle test inlined).
// See traverse.cuh for the data layout and the equivalence argument.
#include <cstdlib>
#include "common.cuh"
#include "traverse_phased.cuh"
#include "traverse_spec.cuh"

namespace b2 {

// variant 2: wide nodes, one thread per ray.
// variant 1: reference LinearBVHNode walk, one thread per ray (baseline).
template <bool ANY, int VARIANT>
__global__ void __launch_bounds__(128) k_trace_simple(DeviceAccel A, const float4* __restrict__ rays, long long n, void* __restrict__ out,
                                                      float* __restrict__ b2_out) {
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
        if (b2_out) b2_out[i] = h.b2;
    }
}

// variant 3: persistent warps with dynamic ray fetch, every lane does what it needs next ("if-if").
// Kept as the A/B baseline of the phase-scheduled kernel in traverse_phased.cuh (variant 0, default).
// Incoherent rays finish after very different numbers of steps, so in the
// one-thread-per-ray kernel a warp runs until its slowest ray is done with most
// lanes idle.  Here each warp keeps pulling work: whenever the number of lanes
// that still hold a live ray drops below a threshold, the idle lanes grab new
// ray indices from a global counter (one atomicAdd per warp, distributed with
// ballot/popc prefix) and the whole warp re-enters the traversal loop.  The
// grid is sized to the machine (SMs x resident CTAs), not to the ray count.

template <bool ANY>
__global__ void __launch_bounds__(128, 4) k_trace_persistent(DeviceAccel A, const float4* __restrict__ rays, long long n,
                                                              void* __restrict__ out, unsigned long long* __restrict__ counter, float* __restrict__ b2_out,
                                                              const int* __restrict__ n_dev) {
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
    h.t = 0.0f; h.prim = 0xffffffffu; h.b0 = 0.0f; h.b1 = 0.0f; h.b2 = 0.0f;
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
            const long long n_rays = ray_count(n, n_dev);
            if ((long long)base + n_idle >= n_rays) exhausted = true;
            if (idle) {
                const int my = __popc(idle_mask & ((1u << lane) - 1u));
                const long long id = (long long)base + my;
                if (id < n_rays) {
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
                    h.t = __int_as_float(0x7f800000); h.prim = 0xffffffffu; h.b0 = 0.0f; h.b1 = 0.0f; h.b2 = 0.0f;
                    cur = A.root_code;
                    float te;
                    if (cur == kDone ||
                        !(slab(r, A.root_bounds[0], A.root_bounds[1], A.root_bounds[2], A.root_bounds[3], A.root_bounds[4], A.root_bounds[5], &te) &&
                          te < t_max)) {
                        // miss at the root: retire immediately
                        if (ANY) ((uint8_t*)out)[id] = 0;
                        else { float4 ov; ov.x = h.t; ov.y = __uint_as_float(h.prim); ov.z = 0.0f; ov.w = 0.0f; ((float4*)out)[id] = ov; if (b2_out) b2_out[id] = 0.0f; }
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
                        if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, first + i)) {
                            if (ANY) {
                                if (alpha_ok_any(A, flags, first + i, o, tc, t_max)) { any_done = true; break; }
                            } else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) {
                                hit = true;
                                t_max = t;
                                h.t = t; h.prim = prim; h.b0 = b0; h.b1 = b1; h.b2 = b2;
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
                        else { float4 ov; ov.x = h.t; ov.y = __uint_as_float(h.prim); ov.z = h.b0; ov.w = h.b1; ((float4*)out)[ray_id] = ov; if (b2_out) b2_out[ray_id] = h.b2; }
                    }
                }
            }
            const unsigned live = __ballot_sync(0xffffffffu, cur != kDone);
            if (live == 0u) break;
            if (!exhausted && __popc(live) <= 24) break;  // go refill
        }
    }
}

// Work accounting for the roofline: walks the 32-byte LinearBVHNode array in the
// reference order and counts, per ray, the nodes whose bounds are tested and the triangle tests —
// the implementation-independent N_node / N_tri of BVHAccel::intersect / intersect_p.
template <bool ANY>
__global__ void __launch_bounds__(128) k_count(DeviceAccel A, const float4* __restrict__ rays, long long n, unsigned long long* __restrict__ totals,
                                               uint2* __restrict__ per_ray) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nn = 0, nt = 0;
    if (i < n && A.root_code != B2_EMPTY_ROOT) {
        float4 r0 = __ldg(rays + 2 * i), r1 = __ldg(rays + 2 * i + 1);
        RayCtx r;
        r.ox = r0.x; r.oy = r0.y; r.oz = r0.z;
        r.ix = 1.0f / r1.x; r.iy = 1.0f / r1.y; r.iz = 1.0f / r1.z;
        r.nx = r.ix < 0.0f; r.ny = r.iy < 0.0f; r.nz = r.iz < 0.0f;
        float t_max = r0.w;
        const TriCtx tc = make_tri_ctx(r1.x, r1.y, r1.z);
        const V3 o = mk(r0.x, r0.y, r0.z);
        int stack[B2_STACK];
        int sp = 0, cur = 0;
        bool done = false;
        while (!done) {
            float4 n0, n1;
            ldg8(A.ref_nodes + 2ll * cur, &n0, &n1);
            float te;
            uint32_t offset = __float_as_uint(n1.z), meta = __float_as_uint(n1.w);
            uint32_t nprims = meta & 0xffffu, axis = (meta >> 16) & 0xffu;
            ++nn;
            if (slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, &te) && te < t_max) {
                if (nprims > 0) {
                    for (uint32_t k = 0; k < nprims && !done; ++k) {
                        V3 p0, p1, p2;
                        uint32_t prim, flags, dummy;
                        load_tri(A.tris, (long long)offset + k, &p0, &p1, &p2, &prim, &flags, &dummy);
                        float t, b0, b1, b2;
                        ++nt;
                        if (triangle_test(o, tc, t_max, p0, p1, p2, &t, &b0, &b1, &b2) && triangle_nondegenerate(p0, p1, p2, A.tris, (long long)offset + k)) {
                            if (ANY) { if (alpha_ok_any(A, flags, (long long)offset + k, o, tc, t_max)) done = true; }
                            else if (alpha_ok<false>(A, flags, prim, b0, b1, b2)) t_max = t;
                        }
                    }
                    if (sp == 0) done = true; else cur = stack[--sp];
                } else {
                    int neg = axis == 0 ? r.nx : (axis == 1 ? r.ny : r.nz);
                    if (neg) { stack[sp++] = cur + 1; cur = (int)offset; }
                    else { stack[sp++] = (int)offset; cur = cur + 1; }
                }
            } else {
                if (sp == 0) done = true; else cur = stack[--sp];
            }
        }
        if (per_ray) per_ray[i] = make_uint2(nn, nt);
    }
    // warp-aggregate, one atomic pair per warp
    for (int d = 16; d > 0; d >>= 1) { nn += __shfl_down_sync(0xffffffffu, nn, d); nt += __shfl_down_sync(0xffffffffu, nt, d); }
    if ((threadIdx.x & 31) == 0 && (nn | nt)) { atomicAdd(totals, (unsigned long long)nn); atomicAdd(totals + 1, (unsigned long long)nt); }
}

int launch_count_work(const DeviceAccel& A, const void* d_rays, int64_t n, int any_hit, unsigned long long* d_totals, void* d_per_ray, cudaStream_t s) {
    if (n <= 0) return B200PT_OK;
    int grid = (int)((n + 127) / 128);
    if (any_hit) k_count<true><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_totals, (uint2*)d_per_ray);
    else k_count<false><<<grid, 128, 0, s>>>(A, (const float4*)d_rays, n, d_totals, (uint2*)d_per_ray);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_count launch");
    return B200PT_OK;
}

static int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

// First use on a device: the work-counter ring and the resident grid of each persistent kernel.
static int persistent_setup(DevCtx* c) {
    std::lock_guard<std::mutex> g(c->mu);
    if (c->ring) return B200PT_OK;
    unsigned long long* ring = nullptr;
    B2_CUDA(cudaMalloc(&ring, DevCtx::kRing * sizeof(unsigned long long)));
    int nb = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_persistent<false>, 128, 0));
    c->persist_grid[0] = c->sm_count * (nb > 0 ? nb : 1);
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_persistent<true>, 128, 0));
    c->persist_grid[1] = c->sm_count * (nb > 0 ? nb : 1);
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_spec2<false, 20, 20, 7, 1, false>, 128, 0));
    c->persist_grid[2] = c->sm_count * (nb > 0 ? nb : 1);
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_spec2<true, 20, 16, 8, 1, false>, 128, 0));
    c->persist_grid[3] = c->sm_count * (nb > 0 ? nb : 1);
    c->ring = ring;
    return B200PT_OK;
}

// A zeroed work counter for one persistent launch on stream s (the caller's own, or the next ring slot).
int trace_work_counter(int device, const TraceLaunch* tl, cudaStream_t s, unsigned long long** out) {
    if (tl && tl->work_ctr) { *out = tl->work_ctr; return B200PT_OK; }
    DevCtx* c = dev_ctx(device);
    if (!c) { b200pt_set_error("traversal: accelerator on a device that was never initialised"); return B200PT_ERR_NO_DEVICE; }
    if (!c->ring) { int rc = persistent_setup(c); if (rc) return rc; }
    unsigned long long* ctr = c->ring + (c->ring_next.fetch_add(1) % DevCtx::kRing);
    B2_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
    *out = ctr;
    return B200PT_OK;
}

template <bool ANY>
static int launch_any(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant, float* d_b2, const TraceLaunch* tl) {
    if (n <= 0) return B200PT_OK;
    const int block = 128;
    const int* n_dev = tl ? tl->n_dev : nullptr;
    if (n_dev && (variant == 1 || variant == 2)) { b200pt_set_error("traversal: device-resident ray counts need a persistent kernel variant"); return B200PT_ERR_INVALID; }
    if (variant == 1) {
        k_trace_simple<ANY, 1><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_out, d_b2);
    } else if (variant == 2) {
        k_trace_simple<ANY, 2><<<grid_for(n, block), block, 0, s>>>(A, (const float4*)d_rays, n, d_out, d_b2);
    } else {
        // Persistent kernels.  0 (default) = loop-free postponed-leaf walk (traverse_spec.cuh); A/B baselines:
        // 3 = "if-if" persistent warps, 4 = phase-scheduled without postponement, 5 = postponed leaf with pop loops.
        DevCtx* c = dev_ctx(A.device);
        if (!c) { b200pt_set_error("traversal: accelerator on a device that was never initialised"); return B200PT_ERR_NO_DEVICE; }
        if (!c->ring) { int rc = persistent_setup(c); if (rc) return rc; }
        if (n >= 0x7fffffffLL) { b200pt_set_error("traversal: at most 2^31-2 rays per launch"); return B200PT_ERR_INVALID; }
        unsigned long long* ctr = nullptr;
        int rc = trace_work_counter(A.device, tl, s, &ctr);
        if (rc) return rc;
        int grid = c->persist_grid[(variant != 3 ? 2 : 0) + (ANY ? 1 : 0)];
        int need = grid_for(n, block);
        if (need < grid) grid = need;
        const float4* R = (const float4*)d_rays;
        // closest-hit: <= 72 registers -> 7 CTAs/SM; any-hit: 64 registers -> 8 CTAs/SM (profiles/r1_variants.txt)
        if (variant == 3) k_trace_persistent<ANY><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
        else if (variant == 4) {
            if (ANY) k_trace_phased<true, 16, 16, 0, 8><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
            else k_trace_phased<false, 16, 16, 0, 7><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
        } else if (variant == 5) {
            k_trace_spec<ANY, 16, 16, ANY ? 8 : 7><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
        } else {
            // B200PT_TRACE_TUNE (A/B): phase-switch / refill thresholds other than the round-1 choice (20, 20 / 16)
            static const int tune = [] { const char* e = std::getenv("B200PT_TRACE_TUNE"); return e ? std::atoi(e) : 0; }();
            if (tune == 1) {
                if (ANY) k_trace_spec2<true, 16, 12, 8, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
                else k_trace_spec2<false, 16, 12, 7, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
            } else if (tune == 2) {
                if (ANY) k_trace_spec2<true, 12, 8, 8, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
                else k_trace_spec2<false, 12, 8, 7, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
            } else if (tune == 3) {
                if (ANY) k_trace_spec2<true, 20, 8, 8, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
                else k_trace_spec2<false, 20, 8, 7, 1><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
            } else {
            // the scene has no alpha textures (the usual case): instantiation without the texture call
            if (ANY) { if (A.alpha) k_trace_spec2<true, 20, 16, 8, 1, true><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
                       else k_trace_spec2<true, 20, 16, 8, 1, false><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev); }
            else { if (A.alpha) k_trace_spec2<false, 20, 20, 7, 1, true><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev);
                   else k_trace_spec2<false, 20, 20, 7, 1, false><<<grid, block, 0, s>>>(A, R, n, d_out, ctr, d_b2, n_dev); }
            }
        }
    }
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_trace launch");
    return B200PT_OK;
}

int launch_intersect(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, int variant, float* d_b2, const TraceLaunch* tl) {
    return launch_any<false>(A, d_rays, n, d_hits, s, variant, d_b2, tl);
}
int launch_occluded(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant, const TraceLaunch* tl) {
    return launch_any<true>(A, d_rays, n, d_out, s, variant, nullptr, tl);
}

}  // namespace b2
