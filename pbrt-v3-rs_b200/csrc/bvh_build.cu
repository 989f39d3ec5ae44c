// GPU construction of the reference's SAH BVH (BVHAccel::new with
// SplitMethod::SAH: accelerators/src/bvh/mod.rs:43-153, sah.rs:26-367,
// common.rs:66-224).  The node array and the primitive order it returns are
// byte-identical to what the reference's single-threaded recursion (and
// host_bvh.cpp) produce; only the schedule is different.
//
// Why this can be parallel and still identical:
//   * every quantity a split decision depends on is a min/max reduction
//     (node bounds, centroid bounds, bucket bounds) or an integer count, i.e.
//     order-independent.  Floats are reduced as order-preserving u32 keys with
//     atomicMin/atomicMax (warp redux.sync -> shared -> one global atomic per
//     block when a block lies inside one node);
//   * the rounding operations (centroid, bucket index, surface area, cost)
//     are evaluated per primitive / per node with the reference's operand
//     order (-fmad=false, IEEE divide);
//   * `itertools::partition` (sah.rs:354) is a front/back swap partition whose
//     result is a pure function of the predicate sequence: with m = number of
//     primitives that satisfy it, the k-th failing primitive among the first m
//     positions trades places with the k-th satisfying primitive counted from
//     the back.  Both ranks come from one prefix sum of the predicate;
//   * leaves own the range [start, end) they were built from, so
//     first_prim_offset == start and ordered_prims is the final permutation;
//   * depth-first node indices follow from subtree sizes (bottom-up) and
//     "first child = me + 1, second child = me + 1 + size(first)" (top-down).
//   * the sign of a zero in a box is order-dependent in the reference
//     (min(a, b) = a < b ? a : b, core/src/pbrt/common.rs:83-108); it never
//     changes a decision, but it is part of the bytes, so the final bounds are
//     produced bottom-up exactly as the reference does: a leaf folds its
//     primitives in order, an interior node is union(child0, child1)
//     (common.rs:150-159).
//
// Schedule: phase A walks the tree level by level with one thread per
// primitive position (nodes with more than kSmall primitives); ranges of at
// most kSmall primitives are finished by ONE thread each running the
// sequential algorithm on its range (phase B, 10^5-10^6 independent threads);
// phases C/D/E compute sizes + bounds bottom-up, depth-first indices top-down
// and emit the 32-byte LinearBVHNode records.
//
// HBM-bound integer/min-max work: nothing here is a contraction.
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_hlbvh.h"

namespace {

constexpr int kBins = 12;     // sah.rs:11
constexpr int kSmall = 32;    // ranges with <= kSmall primitives are finished by one thread
constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

enum { kErrPool = 1, kErrEmptySide = 2, kErrBigLeaf = 3 };

// a = (lo.x, lo.y, lo.z, hi.x), b = (hi.y, hi.z, prim id bits, 0)
struct __align__(32) Item {
    float4 a, b;
};

struct Box {
    float lo[3], hi[3];
};

__host__ __device__ inline float fmin_ref(float a, float b) { return a < b ? a : b; }  // core/src/pbrt/common.rs:83-92
__host__ __device__ inline float fmax_ref(float a, float b) { return a > b ? a : b; }  // :99-108
__device__ inline Box empty_box() {  // bounds3.rs:26-29
    Box b;
    b.lo[0] = b.lo[1] = b.lo[2] = FLT_MAX;
    b.hi[0] = b.hi[1] = b.hi[2] = -FLT_MAX;
    return b;
}
__device__ inline void grow(Box& b, const Box& o) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_ref(b.lo[k], o.lo[k]); b.hi[k] = fmax_ref(b.hi[k], o.hi[k]); }
}
__device__ inline void grow_pt(Box& b, const float* p) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_ref(b.lo[k], p[k]); b.hi[k] = fmax_ref(b.hi[k], p[k]); }
}
__device__ inline float area(const Box& b) {  // bounds3.rs:94-105
    if (b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2]) return 0.0f;
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    float h = dx * dy + dx * dz + dy * dz;
    return h + h;
}
__device__ inline int widest_axis(const Box& b) {  // bounds3.rs:122-134
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (dx > dy && dx > dz) return 0;
    return dy > dz ? 1 : 2;
}
__device__ inline Box item_box(const Item& it) {
    Box b;
    b.lo[0] = it.a.x; b.lo[1] = it.a.y; b.lo[2] = it.a.z;
    b.hi[0] = it.a.w; b.hi[1] = it.b.x; b.hi[2] = it.b.y;
    return b;
}
// BVHPrimitiveInfo::new: centroid = 0.5 * (p_min + p_max)  (common.rs:86)
__device__ inline float centroid(const Box& b, int k) { return 0.5f * (b.lo[k] + b.hi[k]); }

// (12 * Bounds3::offset(c)[dim]) as usize, 12 -> 11  (sah.rs:305-313, bounds3.rs:153-168)
__device__ inline int bin_of(float c, float cb_lo, float cb_hi) {
    float o = c - cb_lo;
    if (cb_hi > cb_lo) o /= cb_hi - cb_lo;
    float v = (float)kBins * o;
    int b = (!(v == v) || v <= 0.0f) ? 0 : (v >= (float)kBins ? kBins : (int)v);  // Rust saturating cast, then 12 -> 11
    return b >= kBins ? kBins - 1 : b;
}

// order-preserving float <-> u32 key (total order, -0 < +0)
__device__ inline uint32_t fkey(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline float funkey(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Phase-A node: a range with more than kSmall primitives.
struct ANode {
    uint32_t start, end;
    int32_t child[2];   // >= 0: ANode id; < 0: ~start of a small range
    uint32_t kind;      // 0 = interior, 1 = leaf
    uint32_t dim, best, m;
    uint32_t size, dfs;
    float cb_lo, cb_hi;  // centroid bounds along dim
    float bounds[6];     // reduction result (decisions); replaced by the reference-order bounds in phase C
    uint32_t acc[12];    // keys: bounds lo[3] (min), hi[3] (max), centroid lo[3] (min), hi[3] (max)
};

struct Bins {
    uint32_t count[kBins];
    uint32_t key[kBins][6];  // lo[3] (min), hi[3] (max)
};

struct Ctl {
    uint32_t n_anodes;   // allocated ANodes
    uint32_t n_small;    // entries in the small list
    uint32_t error;
    uint32_t pad;
};

__device__ inline void init_acc(ANode& nd) {
    const uint32_t kmax = fkey(FLT_MAX), kmin = fkey(-FLT_MAX);
#pragma unroll
    for (int k = 0; k < 3; ++k) { nd.acc[k] = kmax; nd.acc[3 + k] = kmin; nd.acc[6 + k] = kmax; nd.acc[9 + k] = kmin; }
}

// ---------------------------------------------------------------- setup
__global__ void k_tri_bounds(const float* __restrict__ tri_verts, int64_t n, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* v = tri_verts + 9 * i;  // shapes/src/triangle.rs:427-431
    Box b;
    for (int k = 0; k < 3; ++k) b.lo[k] = b.hi[k] = v[k];
    grow_pt(b, v + 3);
    grow_pt(b, v + 6);
    float* o = out + 6 * i;
    o[0] = b.lo[0]; o[1] = b.lo[1]; o[2] = b.lo[2]; o[3] = b.hi[0]; o[4] = b.hi[1]; o[5] = b.hi[2];
}

__global__ void k_make_items(const float* __restrict__ prim_bounds, uint32_t n, Item* __restrict__ items, int32_t* __restrict__ seg,
                             int32_t seg0) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* pb = prim_bounds + 6 * (size_t)i;
    Item it;
    it.a = make_float4(pb[0], pb[1], pb[2], pb[3]);
    it.b = make_float4(pb[4], pb[5], __uint_as_float(i), 0.0f);
    items[i] = it;
    seg[i] = seg0;
}

__global__ void k_init_root(ANode* nodes, Ctl* ctl, uint2* small_list, uint32_t n) {
    ctl->error = 0;
    if (n > (uint32_t)kSmall) {
        ANode& r = nodes[0];
        r.start = 0; r.end = n; r.child[0] = r.child[1] = 0; r.kind = 0; r.size = 0; r.dfs = 0;
        init_acc(r);
        ctl->n_anodes = 1;
        ctl->n_small = 0;
    } else {
        small_list[0] = make_uint2(0, n);
        ctl->n_anodes = 0;
        ctl->n_small = 1;
    }
}

// ---------------------------------------------------------------- phase A, per level
// 1. node bounds + centroid bounds of every active node
__global__ void __launch_bounds__(kBlock) k_level_reduce(const Item* __restrict__ items, const int32_t* __restrict__ seg, ANode* nodes, uint32_t n) {
    __shared__ uint32_t sh[12];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t sb = seg[blockIdx.x * blockDim.x];
    const int32_t s = i < n ? seg[i] : sb;
    const bool blk_uni = __syncthreads_and(s == sb);
    if (blk_uni && sb < 0) return;
    uint32_t key[12];
    if (i < n && s >= 0) {
        Item it = items[i];
        Box b = item_box(it);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float c = centroid(b, k);
            key[k] = fkey(b.lo[k]); key[3 + k] = fkey(b.hi[k]); key[6 + k] = fkey(c); key[9 + k] = fkey(c);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { key[k] = key[6 + k] = 0xffffffffu; key[3 + k] = key[9 + k] = 0u; }
    }
    if (blk_uni) {
        if (threadIdx.x < 12) sh[threadIdx.x] = ((threadIdx.x / 3) & 1) ? 0u : 0xffffffffu;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const bool is_max = (k / 3) & 1;
            uint32_t r = is_max ? __reduce_max_sync(kFull, key[k]) : __reduce_min_sync(kFull, key[k]);
            if ((threadIdx.x & 31) == 0) { if (is_max) atomicMax(&sh[k], r); else atomicMin(&sh[k], r); }
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            const bool is_max = (threadIdx.x / 3) & 1;
            if (is_max) atomicMax(&nodes[sb].acc[threadIdx.x], sh[threadIdx.x]); else atomicMin(&nodes[sb].acc[threadIdx.x], sh[threadIdx.x]);
        }
        return;
    }
    const int32_t s0 = __shfl_sync(kFull, s, 0);
    const bool warp_uni = __all_sync(kFull, s == s0);
    if (warp_uni) {
        if (s0 < 0) return;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const bool is_max = (k / 3) & 1;
            uint32_t r = is_max ? __reduce_max_sync(kFull, key[k]) : __reduce_min_sync(kFull, key[k]);
            if ((threadIdx.x & 31) == 0) { if (is_max) atomicMax(&nodes[s0].acc[k], r); else atomicMin(&nodes[s0].acc[k], r); }
        }
        return;
    }
    if (i < n && s >= 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            if ((k / 3) & 1) atomicMax(&nodes[s].acc[k], key[k]); else atomicMin(&nodes[s].acc[k], key[k]);
        }
    }
}

// 2. split axis, zero-extent leaves (sah.rs:54-63); clears the level's buckets
__global__ void k_level_decide1(ANode* nodes, uint32_t first, uint32_t count, Bins* bins, Ctl* ctl) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ANode& nd = nodes[first + j];
    Box cb;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        nd.bounds[k] = funkey(nd.acc[k]);
        nd.bounds[3 + k] = funkey(nd.acc[3 + k]);
        cb.lo[k] = funkey(nd.acc[6 + k]);
        cb.hi[k] = funkey(nd.acc[9 + k]);
    }
    const int dim = widest_axis(cb);
    nd.dim = dim;
    nd.cb_lo = cb.lo[dim];
    nd.cb_hi = cb.hi[dim];
    if (cb.hi[dim] == cb.lo[dim]) {
        nd.kind = 1;
        if (nd.end - nd.start >= 65536u) atomicMax(&ctl->error, (uint32_t)kErrBigLeaf);  // mod.rs:137 asserts
        return;
    }
    Bins& B = bins[j];
    const uint32_t kmax = fkey(FLT_MAX), kmin = fkey(-FLT_MAX);
    for (int b = 0; b < kBins; ++b) {
        B.count[b] = 0;
        for (int k = 0; k < 3; ++k) { B.key[b][k] = kmax; B.key[b][3 + k] = kmin; }
    }
}

// 3. bucket counts and bounds (sah.rs:303-316)
__global__ void __launch_bounds__(kBlock) k_level_bin(const Item* __restrict__ items, const int32_t* __restrict__ seg, const ANode* __restrict__ nodes,
                                                       uint32_t first, Bins* bins, uint32_t n) {
    __shared__ Bins sh;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t sb = seg[blockIdx.x * blockDim.x];
    const int32_t s = i < n ? seg[i] : sb;
    const bool blk_uni = __syncthreads_and(s == sb);
    if (blk_uni && (sb < 0 || nodes[sb].kind != 0)) return;
    const bool active = i < n && s >= 0 && nodes[s].kind == 0;
    int b = 0;
    uint32_t key[6];
    if (active) {
        const ANode& nd = nodes[s];
        Item it = items[i];
        Box bx = item_box(it);
        const int dim = nd.dim;
        b = bin_of(centroid(bx, dim), nd.cb_lo, nd.cb_hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) { key[k] = fkey(bx.lo[k]); key[3 + k] = fkey(bx.hi[k]); }
    }
    if (blk_uni) {
        uint32_t* w = reinterpret_cast<uint32_t*>(&sh);
        const uint32_t kmax = fkey(FLT_MAX), kmin = fkey(-FLT_MAX);
        for (int t = threadIdx.x; t < kBins * 7; t += blockDim.x) {
            if (t < kBins) w[t] = 0;
            else { int k = (t - kBins) % 6; w[t] = k < 3 ? kmax : kmin; }
        }
        __syncthreads();
        if (active) {
            atomicAdd(&sh.count[b], 1u);
#pragma unroll
            for (int k = 0; k < 3; ++k) { atomicMin(&sh.key[b][k], key[k]); atomicMax(&sh.key[b][3 + k], key[3 + k]); }
        }
        __syncthreads();
        Bins& G = bins[sb - first];
        for (int t = threadIdx.x; t < kBins * 7; t += blockDim.x) {
            if (t < kBins) { if (sh.count[t]) atomicAdd(&G.count[t], sh.count[t]); }
            else {
                int bb = (t - kBins) / 6, k = (t - kBins) % 6;
                if (sh.count[bb]) { if (k < 3) atomicMin(&G.key[bb][k], sh.key[bb][k]); else atomicMax(&G.key[bb][k], sh.key[bb][k]); }
            }
        }
        return;
    }
    if (active) {
        Bins& G = bins[s - first];
        atomicAdd(&G.count[b], 1u);
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&G.key[b][k], key[k]); atomicMax(&G.key[b][3 + k], key[3 + k]); }
    }
}

// 4. SAH cost of the 11 candidate planes, split / leaf decision, child allocation (sah.rs:318-366)
__global__ void k_level_decide2(ANode* nodes, uint32_t first, uint32_t count, const Bins* bins, Ctl* ctl, uint2* small_list,
                                uint32_t pool_cap, int max_prims) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ANode& nd = nodes[first + j];
    if (nd.kind != 0) return;
    const Bins& B = bins[j];
    Box bin_box[kBins];
    uint32_t bin_count[kBins];
    for (int b = 0; b < kBins; ++b) {
        bin_count[b] = B.count[b];
        for (int k = 0; k < 3; ++k) { bin_box[b].lo[k] = funkey(B.key[b][k]); bin_box[b].hi[k] = funkey(B.key[b][3 + k]); }
    }
    // prefix / suffix sweeps (exact: unions are min/max, counts are integers)
    Box right[kBins - 1];
    uint32_t nr[kBins - 1];
    Box acc = empty_box();
    uint32_t cnt = 0;
    for (int b = kBins - 1; b >= 1; --b) { grow(acc, bin_box[b]); cnt += bin_count[b]; right[b - 1] = acc; nr[b - 1] = cnt; }
    Box bb;
    for (int k = 0; k < 3; ++k) { bb.lo[k] = nd.bounds[k]; bb.hi[k] = nd.bounds[3 + k]; }
    const float total_area = area(bb);
    acc = empty_box();
    cnt = 0;
    float best = 0.0f;
    int best_bin = 0;
    uint32_t m = 0;
    for (int b = 0; b < kBins - 1; ++b) {  // sah.rs:321-347: first minimum wins
        grow(acc, bin_box[b]);
        cnt += bin_count[b];
        float cost = 1.0f + ((float)cnt * area(acc) + (float)nr[b] * area(right[b])) / total_area;
        if (b == 0 || cost < best) { best = cost; best_bin = b; m = cnt; }
    }
    const uint32_t n = nd.end - nd.start;
    if (!(n > (uint32_t)max_prims || best < (float)n)) {  // sah.rs:351 (n <= 255 here)
        nd.kind = 1;
        return;
    }
    if (m == 0 || m == n) {  // the reference recurses on an empty range and panics (sah.rs:37)
        atomicMax(&ctl->error, (uint32_t)kErrEmptySide);
        nd.kind = 1;
        return;
    }
    nd.best = best_bin;
    nd.m = m;
    for (int c = 0; c < 2; ++c) {
        const uint32_t cs = c == 0 ? nd.start : nd.start + m, ce = c == 0 ? nd.start + m : nd.end;
        if (ce - cs <= (uint32_t)kSmall) {
            small_list[atomicAdd(&ctl->n_small, 1u)] = make_uint2(cs, ce);
            nd.child[c] = ~(int32_t)cs;
        } else {
            uint32_t id = atomicAdd(&ctl->n_anodes, 1u);
            if (id >= pool_cap) { atomicMax(&ctl->error, (uint32_t)kErrPool); nd.child[c] = -1; nd.kind = 1; return; }
            ANode& ch = nodes[id];
            ch.start = cs; ch.end = ce; ch.child[0] = ch.child[1] = 0; ch.kind = 0; ch.size = 0; ch.dfs = 0;
            init_acc(ch);
            nd.child[c] = (int32_t)id;
        }
    }
}

// 5. partition predicate (sah.rs:354-361) + per-block counts
__global__ void __launch_bounds__(kBlock) k_level_pred(const Item* __restrict__ items, const int32_t* __restrict__ seg, const ANode* __restrict__ nodes,
                                                        uint8_t* __restrict__ pred, uint32_t* __restrict__ block_sum, uint32_t n) {
    __shared__ uint32_t warp_sum[kBlock / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = 0;
    if (i < n) {
        const int32_t s = seg[i];
        if (s >= 0) {
            const ANode& nd = nodes[s];
            if (nd.kind == 0) {
                Item it = items[i];
                Box bx = item_box(it);
                p = bin_of(centroid(bx, nd.dim), nd.cb_lo, nd.cb_hi) <= (int)nd.best ? 1u : 0u;
            }
        }
        pred[i] = (uint8_t)p;
    }
    uint32_t c = __popc(__ballot_sync(kFull, p));
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kBlock / 32; ++w) t += warp_sum[w];
        block_sum[blockIdx.x] = t;
    }
}

// 6. exclusive scan of the block counts (one block)
__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t* block_sum, uint32_t n_blocks) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sum[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(kFull, x, d); if ((threadIdx.x & 31) >= d) x += y; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = wsum[threadIdx.x], xs = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(kFull, xs, d); if (threadIdx.x >= d) xs += y; }
            wsum[threadIdx.x] = xs - w;  // exclusive
        }
        __syncthreads();
        const uint32_t excl = carry + wsum[threadIdx.x >> 5] + x - v;
        if (i < n_blocks) block_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

// Exclusive scan of a LARGE array in place (the radix sort's 64 x n_blocks histogram: 2.5 M counters for 10 M primitives, which the
// one-block scan above walked in 2 442 sequential steps = 1.6 ms per pass, 8 of the build's 13 ms): sums of 1024-element chunks,
// k_scan_blocks over those few sums, then every chunk scans itself and adds its offset.
__global__ void __launch_bounds__(1024) k_scan_chunk_sums(const uint32_t* __restrict__ a, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t wsum[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    uint32_t v = i < n ? a[i] : 0;
    v = __reduce_add_sync(kFull, v);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t t = __reduce_add_sync(kFull, wsum[threadIdx.x]);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t* __restrict__ a, uint32_t n, const uint32_t* __restrict__ sums_excl) {
    __shared__ uint32_t wsum[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    const uint32_t v = i < n ? a[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(kFull, x, d); if ((threadIdx.x & 31) >= d) x += y; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = wsum[threadIdx.x], xs = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(kFull, xs, d); if (threadIdx.x >= d) xs += y; }
        wsum[threadIdx.x] = xs - w;  // exclusive
    }
    __syncthreads();
    if (i < n) a[i] = sums_excl[blockIdx.x] + wsum[threadIdx.x >> 5] + x - v;
}

// 7. T[i] = number of predicate-true positions before i (whole array)
__global__ void __launch_bounds__(kBlock) k_scan_apply(const uint8_t* __restrict__ pred, const uint32_t* __restrict__ block_sum, uint32_t* __restrict__ T,
                                                        uint32_t n) {
    __shared__ uint32_t warp_sum[kBlock / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t p = i < n ? pred[i] : 0;
    const uint32_t bal = __ballot_sync(kFull, p);
    const uint32_t lane = threadIdx.x & 31;
    if (lane == 0) warp_sum[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    uint32_t off = block_sum[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += warp_sum[w];
    if (i < n) T[i] = off + __popc(bal & ((1u << lane) - 1u));
}

// 8. pair lists: k-th misplaced "false" from the front, k-th misplaced "true" from the back
__global__ void __launch_bounds__(kBlock) k_level_pairs(const int32_t* __restrict__ seg, const ANode* __restrict__ nodes, const uint8_t* __restrict__ pred,
                                                         const uint32_t* __restrict__ T, uint32_t* __restrict__ list_f, uint32_t* __restrict__ list_t,
                                                         uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = seg[i];
    if (s < 0) return;
    const ANode& nd = nodes[s];
    if (nd.kind != 0) return;
    const uint32_t p = i - nd.start, tb = T[i] - T[nd.start], m = nd.m;
    const bool pr = pred[i] != 0;
    if (p < m && !pr) list_f[nd.start + (p - tb)] = i;
    else if (p >= m && pr) list_t[nd.start + (m - tb - 1)] = i;
}

// 9. swaps + segment ids of the next level
__global__ void __launch_bounds__(kBlock) k_level_swap(Item* __restrict__ items, int32_t* __restrict__ seg, const ANode* __restrict__ nodes,
                                                        const uint32_t* __restrict__ T, const uint32_t* __restrict__ list_f,
                                                        const uint32_t* __restrict__ list_t, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = seg[i];
    if (s < 0) return;
    const ANode& nd = nodes[s];
    if (nd.kind != 0) { seg[i] = -1; return; }
    const uint32_t j = i - nd.start, m = nd.m;
    const uint32_t n_pairs = m - (T[nd.start + m] - T[nd.start]);
    if (j < n_pairs) {
        const uint32_t a = list_f[i], b = list_t[i];
        Item ia = items[a], ib = items[b];
        items[a] = ib;
        items[b] = ia;
    }
    const int32_t c = nd.child[j < m ? 0 : 1];
    seg[i] = c >= 0 ? c : -1;
}

// ---------------------------------------------------------------- phase B: one thread per small range
__device__ inline void store_node(b200pt_bvh_node* dst, const Box& b, uint32_t offset, uint32_t n_prims, uint32_t axis) {
    float4* d = reinterpret_cast<float4*>(dst);
    d[0] = make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]);
    d[1] = make_float4(b.hi[1], b.hi[2], __uint_as_float(offset), __uint_as_float((n_prims & 0xffffu) | (axis << 16)));
}
__device__ inline Box node_box(const b200pt_bvh_node* nd) {
    Box b;
    for (int k = 0; k < 3; ++k) { b.lo[k] = nd->bounds[k]; b.hi[k] = nd->bounds[3 + k]; }
    return b;
}

__global__ void __launch_bounds__(128) k_small_build(Item* __restrict__ items, const uint2* __restrict__ small_list, uint32_t n_small, int max_prims,
                                                      b200pt_bvh_node* __restrict__ tmp, uint32_t* __restrict__ small_size, Ctl* ctl) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_small) return;
    const uint2 range = small_list[t];
    b200pt_bvh_node* out = tmp + 2 * (size_t)range.x;
    struct Job { uint32_t b, e; int32_t parent; };
    Job stack[kSmall + 2];
    int sp = 0;
    stack[sp++] = Job{range.x, range.y, -1};
    uint32_t n_nodes = 0;
    while (sp > 0) {
        const Job job = stack[--sp];
        const uint32_t me = n_nodes++;
        if (job.parent >= 0) out[job.parent].offset = me;  // local index; rebased at emit
        Box bb = empty_box();
        for (uint32_t i = job.b; i < job.e; ++i) grow(bb, item_box(items[i]));
        const uint32_t count = job.e - job.b;
        if (count == 1) { store_node(out + me, bb, job.b, count, 0); continue; }  // sah.rs:47-49
        Box cb = empty_box();
        for (uint32_t i = job.b; i < job.e; ++i) {
            Box bx = item_box(items[i]);
            float c[3] = {centroid(bx, 0), centroid(bx, 1), centroid(bx, 2)};
            grow_pt(cb, c);
        }
        const int dim = widest_axis(cb);
        if (cb.hi[dim] == cb.lo[dim]) { store_node(out + me, bb, job.b, count, 0); continue; }  // sah.rs:61
        uint32_t mid;
        if (count <= 2) {
            // sah.rs:81-83: equal counts; the smaller centroid comes first
            mid = (job.b + job.e) / 2;
            Item i0 = items[job.b], i1 = items[job.e - 1];
            if (centroid(item_box(i1), dim) < centroid(item_box(i0), dim)) { items[job.b] = i1; items[job.e - 1] = i0; }
        } else {
            uint32_t bin_count[kBins];
            Box bin_box[kBins];
            for (int b = 0; b < kBins; ++b) { bin_count[b] = 0; bin_box[b] = empty_box(); }
            uint32_t bucket_of[kSmall];  // bucket of every primitive of the range (count <= kSmall)
            for (uint32_t i = job.b; i < job.e; ++i) {
                Box bx = item_box(items[i]);
                int b = bin_of(centroid(bx, dim), cb.lo[dim], cb.hi[dim]);
                bucket_of[i - job.b] = b;
                bin_count[b] += 1;
                grow(bin_box[b], bx);
            }
            Box right[kBins - 1];
            uint32_t nr[kBins - 1];
            Box acc = empty_box();
            uint32_t cnt = 0;
            for (int b = kBins - 1; b >= 1; --b) { grow(acc, bin_box[b]); cnt += bin_count[b]; right[b - 1] = acc; nr[b - 1] = cnt; }
            const float total_area = area(bb);
            acc = empty_box();
            cnt = 0;
            float best = 0.0f;
            int best_bin = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                grow(acc, bin_box[b]);
                cnt += bin_count[b];
                float cost = 1.0f + ((float)cnt * area(acc) + (float)nr[b] * area(right[b])) / total_area;
                if (b == 0 || cost < best) { best = cost; best_bin = b; }
            }
            if (!(count > (uint32_t)max_prims || best < (float)count)) { store_node(out + me, bb, job.b, count, 0); continue; }
            // itertools::partition, front/back swap (the bucket of a primitive moves with it)
            uint32_t f = job.b, bk = job.e, split = 0;
            while (f < bk) {
                const uint32_t front = f++;
                if (!((int)bucket_of[front - job.b] <= best_bin)) {
                    bool found = false;
                    while (bk > f) {
                        --bk;
                        if ((int)bucket_of[bk - job.b] <= best_bin) { found = true; break; }
                    }
                    if (!found) break;
                    Item ia = items[front], ib = items[bk];
                    items[front] = ib;
                    items[bk] = ia;
                    uint32_t tb = bucket_of[front - job.b];
                    bucket_of[front - job.b] = bucket_of[bk - job.b];
                    bucket_of[bk - job.b] = tb;
                }
                ++split;
            }
            mid = job.b + split;
        }
        if (mid == job.b || mid == job.e) {
            atomicMax(&ctl->error, (uint32_t)kErrEmptySide);
            store_node(out + me, bb, job.b, count, 0);
            continue;
        }
        store_node(out + me, bb, 0, 0, (uint32_t)dim);
        stack[sp++] = Job{mid, job.e, (int32_t)me};  // second child: after the whole first subtree
        stack[sp++] = Job{job.b, mid, -1};           // first child: me + 1
    }
    // interior bounds the way the reference forms them: union(child0, child1) (common.rs:150-159)
    for (int32_t i = (int32_t)n_nodes - 1; i >= 0; --i) {
        if (out[i].n_primitives != 0) continue;
        Box a = node_box(out + i + 1), b = node_box(out + out[i].offset);
        grow(a, b);
        for (int k = 0; k < 3; ++k) { out[i].bounds[k] = a.lo[k]; out[i].bounds[3 + k] = a.hi[k]; }
    }
    small_size[range.x] = n_nodes;
}

// ---------------------------------------------------------------- phase C: sizes + reference-order bounds, bottom-up
__global__ void k_finish_up(ANode* nodes, uint32_t first, uint32_t count, const Item* __restrict__ items, const b200pt_bvh_node* __restrict__ tmp,
                            const uint32_t* __restrict__ small_size) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ANode& nd = nodes[first + j];
    if (nd.kind != 0) {
        Box bb = empty_box();
        for (uint32_t i = nd.start; i < nd.end; ++i) grow(bb, item_box(items[i]));
        for (int k = 0; k < 3; ++k) { nd.bounds[k] = bb.lo[k]; nd.bounds[3 + k] = bb.hi[k]; }
        nd.size = 1;
        return;
    }
    uint32_t size = 1;
    Box cbx[2];
    for (int c = 0; c < 2; ++c) {
        const int32_t code = nd.child[c];
        if (code >= 0) {
            const ANode& ch = nodes[code];
            size += ch.size;
            for (int k = 0; k < 3; ++k) { cbx[c].lo[k] = ch.bounds[k]; cbx[c].hi[k] = ch.bounds[3 + k]; }
        } else {
            const uint32_t cs = (uint32_t)~code;
            size += small_size[cs];
            cbx[c] = node_box(tmp + 2 * (size_t)cs);
        }
    }
    grow(cbx[0], cbx[1]);
    for (int k = 0; k < 3; ++k) { nd.bounds[k] = cbx[0].lo[k]; nd.bounds[3 + k] = cbx[0].hi[k]; }
    nd.size = size;
}

// ---------------------------------------------------------------- phase D: depth-first indices top-down + emit of phase-A nodes
__global__ void k_finish_down(ANode* nodes, uint32_t first, uint32_t count, const uint32_t* __restrict__ small_size, uint32_t* __restrict__ small_dfs,
                              b200pt_bvh_node* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    ANode& nd = nodes[first + j];
    Box bb;
    for (int k = 0; k < 3; ++k) { bb.lo[k] = nd.bounds[k]; bb.hi[k] = nd.bounds[3 + k]; }
    if (nd.kind != 0) { store_node(out + nd.dfs, bb, nd.start, nd.end - nd.start, 0); return; }
    uint32_t at = nd.dfs + 1, second = 0;
    for (int c = 0; c < 2; ++c) {
        const int32_t code = nd.child[c];
        if (c == 1) second = at;
        if (code >= 0) { nodes[code].dfs = at; at += nodes[code].size; }
        else { const uint32_t cs = (uint32_t)~code; small_dfs[cs] = at; at += small_size[cs]; }
    }
    store_node(out + nd.dfs, bb, second, 0, nd.dim);
}

// ---------------------------------------------------------------- phase E: emit the small subtrees (one warp per subtree)
__global__ void __launch_bounds__(256) k_emit_small(const uint2* __restrict__ small_list, uint32_t n_small, const b200pt_bvh_node* __restrict__ tmp,
                                                     const uint32_t* __restrict__ small_size, const uint32_t* __restrict__ small_dfs,
                                                     b200pt_bvh_node* __restrict__ out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_small) return;
    const uint32_t cs = small_list[w].x, size = small_size[cs], base = small_dfs[cs];
    const float4* src = reinterpret_cast<const float4*>(tmp + 2 * (size_t)cs);
    float4* dst = reinterpret_cast<float4*>(out + base);
    for (uint32_t q = lane; q < 2 * size; q += 32) {
        float4 v = src[q];
        if (q & 1) {
            const uint32_t meta = __float_as_uint(v.w);
            if ((meta & 0xffffu) == 0) v.z = __uint_as_float(__float_as_uint(v.z) + base);  // interior: second child, rebased
        }
        dst[q] = v;
    }
}

__global__ void k_ordered(const Item* __restrict__ items, uint32_t n, uint32_t* __restrict__ ordered) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ordered[i] = __float_as_uint(items[i].b.z);
}

// ================================================================ SplitMethod::HLBVH on the GPU
// hlbvh.rs:33-449 / morton.rs:37-120 with the reference behaviours kept by host_hlbvh.cpp (float-bit-pattern Morton
// codes, treelets emitted in order).  Stages: scene bounds (key atomics) -> Morton codes -> 5 stable 6-bit LSD radix
// passes (the reference's own schedule: per-block digit histograms, one scan over digit-major counts, ranked scatter)
// -> treelet starts (flag + scan) -> one thread per treelet emits its LBVH in pre-order into its slice of a scratch
// array -> the <= 4096 treelet roots go to the host for the upper SAH layout (b200pt_hlbvh_upper_layout; tiny and
// sequential) -> blocks and upper nodes are written to their final places.
__global__ void __launch_bounds__(kBlock) k_hl_bounds(const float* __restrict__ pb, uint32_t n, uint32_t* acc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) { key[k] = i < n ? fkey(pb[6 * (size_t)i + k]) : 0xffffffffu; key[3 + k] = i < n ? fkey(pb[6 * (size_t)i + 3 + k]) : 0u; }
    // warp reduction, then one thread per block folds the warps' results: 6 global atomics per block instead of per warp
    // (10 M primitives: 1.9 M atomics on six addresses took 1.27 ms of a 3.9 ms build)
    __shared__ uint32_t part[kBlock / 32][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        uint32_t r = k < 3 ? __reduce_min_sync(kFull, key[k]) : __reduce_max_sync(kFull, key[k]);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        uint32_t r = part[0][k];
        for (int w = 1; w < kBlock / 32; ++w) r = k < 3 ? min(r, part[w][k]) : max(r, part[w][k]);
        if (k < 3) atomicMin(&acc[k], r); else atomicMax(&acc[k], r);
    }
}
__device__ inline uint32_t spread3(uint32_t x) {  // left_shift_3 (morton.rs:102-120), debug_assert compiled out
    if (x == (1u << 10)) x -= 1;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__global__ void __launch_bounds__(kBlock) k_hl_morton(const float* __restrict__ pb, uint32_t n, const uint32_t* __restrict__ acc, uint32_t* __restrict__ code,
                                                       uint32_t* __restrict__ val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = 0;
#pragma unroll
    for (int k = 2; k >= 0; --k) {
        const float lo = funkey(acc[k]), hi = funkey(acc[3 + k]);
        const float cen = 0.5f * (pb[6 * (size_t)i + k] + pb[6 * (size_t)i + 3 + k]);  // common.rs:86
        float o = cen - lo;                                                         // Bounds3::offset
        if (hi > lo) o /= hi - lo;
        c |= spread3(__float_as_uint(o * 1024.0f)) << k;                            // morton.rs:43-49: float_to_bits
    }
    code[i] = c;
    val[i] = i;
}
// radix pass, step 1: digit histogram of every block, digit-major (hist[d * n_blocks + block])
__global__ void __launch_bounds__(kBlock) k_hl_hist(const uint32_t* __restrict__ code, uint32_t n, int shift, uint32_t* __restrict__ hist, uint32_t n_blocks) {
    __shared__ uint32_t h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&h[(code[i] >> shift) & 63u], 1u);
    __syncthreads();
    if (threadIdx.x < 64) hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}
// step 3 (after the exclusive scan of hist): stable scatter.  Rank inside the block = elements of the same digit in
// earlier warps + earlier lanes of the same warp (match_any), which is the order the sequential pass visits them in.
__global__ void __launch_bounds__(kBlock) k_hl_scatter(const uint32_t* __restrict__ code, const uint32_t* __restrict__ val, uint32_t n, int shift,
                                                        const uint32_t* __restrict__ hist, uint32_t n_blocks, uint32_t* __restrict__ code_out,
                                                        uint32_t* __restrict__ val_out) {
    __shared__ uint32_t wc[kBlock / 32][64];
    for (int t = threadIdx.x; t < (kBlock / 32) * 64; t += blockDim.x) (&wc[0][0])[t] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = i < n;
    const uint32_t c = live ? code[i] : 0, d = live ? ((c >> shift) & 63u) : 64u + lane;  // dead lanes match nobody
    const uint32_t peers = __match_any_sync(kFull, d);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (live && rank == 0) wc[warp][d] = __popc(peers);
    __syncthreads();
    if (!live) return;
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += wc[w][d];
    const uint32_t dst = hist[d * n_blocks + blockIdx.x] + before + rank;
    code_out[dst] = c;
    val_out[dst] = val[i];
}
// treelet starts: runs of equal top-12 bits (hlbvh.rs:53-69)
__global__ void __launch_bounds__(kBlock) k_hl_flags(const uint32_t* __restrict__ code, uint32_t n, uint8_t* __restrict__ flag, uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t warp_sum[kBlock / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = 0;
    if (i < n) { p = (i == 0 || ((code[i] ^ code[i - 1]) & 0x3FFC0000u) != 0) ? 1u : 0u; flag[i] = (uint8_t)p; }
    const uint32_t c = __popc(__ballot_sync(kFull, p));
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kBlock / 32; ++w) t += warp_sum[w]; block_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(kBlock) k_hl_starts(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ T, uint32_t n, uint32_t* __restrict__ starts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) starts[T[i]] = i;
}
// emit_lbvh (hlbvh.rs:243-345) for one treelet per thread, iteratively, straight into pre-order
__global__ void __launch_bounds__(64) k_hl_treelets(const float* __restrict__ pb, const uint32_t* __restrict__ code, const uint32_t* __restrict__ val, uint32_t n,
                                                     const uint32_t* __restrict__ starts, uint32_t n_treelets, int max_prims, b200pt_bvh_node* __restrict__ tmp,
                                                     uint32_t* __restrict__ sizes, Ctl* ctl) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_treelets) return;
    const uint32_t first = starts[t], end = t + 1 < n_treelets ? starts[t + 1] : n;
    b200pt_bvh_node* out = tmp + 2 * (size_t)first;
    struct Job { uint32_t first, count; int32_t bit, parent; };
    Job stack[48];  // at most one pending second child per split bit (18) plus the current first child
    int sp = 0;
    stack[sp++] = Job{first, end - first, 30 - 1 - 12, -1};
    uint32_t n_nodes = 0;
    while (sp > 0) {
        Job j = stack[--sp];
        while (j.bit >= 0 && j.count >= (uint32_t)max_prims && ((code[j.first] ^ code[j.first + j.count - 1]) & (1u << j.bit)) == 0) --j.bit;  // hlbvh.rs:278-291
        const uint32_t me = n_nodes++;
        if (j.parent >= 0) out[j.parent].offset = me;
        if (j.bit == -1 || j.count < (uint32_t)max_prims) {  // hlbvh.rs:257-273
            Box b = empty_box();
            for (uint32_t i = 0; i < j.count; ++i) {
                const float* q = pb + 6 * (size_t)val[j.first + i];
                Box o;
                o.lo[0] = q[0]; o.lo[1] = q[1]; o.lo[2] = q[2]; o.hi[0] = q[3]; o.hi[1] = q[4]; o.hi[2] = q[5];
                grow(b, o);
            }
            if (j.count >= 65536u) atomicMax(&ctl->error, (uint32_t)kErrBigLeaf);
            store_node(out + me, b, j.first, j.count, 0);
            continue;
        }
        const uint32_t mask = 1u << j.bit;
        uint32_t lo = 0, hi = j.count - 1;
        while (lo + 1 != hi) {  // hlbvh.rs:293-316
            const uint32_t mid = (lo + hi) / 2;
            if (((code[j.first + lo] ^ code[j.first + mid]) & mask) == 0) lo = mid; else hi = mid;
        }
        store_node(out + me, empty_box(), 0, 0, (uint32_t)(j.bit % 3));
        stack[sp++] = Job{j.first + hi, j.count - hi, j.bit - 1, (int32_t)me};
        stack[sp++] = Job{j.first, hi, j.bit - 1, -1};
    }
    for (int32_t i = (int32_t)n_nodes - 1; i >= 0; --i) {  // interior bounds = union(child0, child1), common.rs:150-159
        if (out[i].n_primitives != 0) continue;
        Box a = node_box(out + i + 1), b = node_box(out + out[i].offset);
        grow(a, b);
        for (int k = 0; k < 3; ++k) { out[i].bounds[k] = a.lo[k]; out[i].bounds[3 + k] = a.hi[k]; }
    }
    sizes[t] = n_nodes;
}
// ---- the same treelets, level-parallel (default) ---------------------------------------------------------------------
// emit_lbvh splits a range at the highest Morton bit (<= the parent's bit - 1) in which its codes differ, so a node's children
// are a pure function of (first, count, bit): all nodes of one depth are independent.  One launch per depth (the bit falls by at
// least one per level: 19 launches) lets one thread create each node and allocate its two children with one atomicAdd; the
// ids of a depth are contiguous, so the next launch's work list is the id range itself.  Pre-order positions follow from
// subtree sizes (bottom-up, together with the bounds: union(child0, child1) as common.rs:150-159) and "first child = me + 1,
// second child = me + 1 + size(first)" (top-down).  Output: the same `tmp` / `sizes` the one-thread-per-treelet kernel writes.
struct HlLevels {
    uint32_t *first, *count, *c0, *size, *pos, *tre;
    int32_t* bit;
    float* box;          // 6 floats per node
    uint32_t* counter;   // nodes allocated
    uint32_t* lvl;       // lvl[d] = first id of depth d
};
__global__ void k_hlp_init(HlLevels T, const uint32_t* __restrict__ starts, uint32_t n_treelets, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { *T.counter = n_treelets; T.lvl[0] = 0; T.lvl[1] = n_treelets; }
    if (t >= n_treelets) return;
    const uint32_t first = starts[t], end = t + 1 < n_treelets ? starts[t + 1] : n;
    T.first[t] = first; T.count[t] = end - first; T.bit[t] = 30 - 1 - 12; T.tre[t] = t;
}
__global__ void __launch_bounds__(256) k_hlp_level(HlLevels T, int depth, const float* __restrict__ pb, const uint32_t* __restrict__ code, const uint32_t* __restrict__ val,
                                                    int max_prims, Ctl* ctl) {
    const uint32_t lo_id = T.lvl[depth], hi_id = T.lvl[depth + 1];
    for (uint32_t id = lo_id + blockIdx.x * blockDim.x + threadIdx.x; id < hi_id; id += gridDim.x * blockDim.x) {
        const uint32_t first = T.first[id], count = T.count[id];
        int bit = T.bit[id];
        while (bit >= 0 && count >= (uint32_t)max_prims && ((code[first] ^ code[first + count - 1]) & (1u << bit)) == 0) --bit;  // hlbvh.rs:278-291
        T.bit[id] = bit;
        if (bit == -1 || count < (uint32_t)max_prims) {  // hlbvh.rs:257-273
            Box b = empty_box();
            for (uint32_t i = 0; i < count; ++i) {
                const float* q = pb + 6 * (size_t)val[first + i];
                Box o;
                o.lo[0] = q[0]; o.lo[1] = q[1]; o.lo[2] = q[2]; o.hi[0] = q[3]; o.hi[1] = q[4]; o.hi[2] = q[5];
                grow(b, o);
            }
            if (count >= 65536u) atomicMax(&ctl->error, (uint32_t)kErrBigLeaf);
            float* bx = T.box + 6 * (size_t)id;
            for (int k = 0; k < 3; ++k) { bx[k] = b.lo[k]; bx[3 + k] = b.hi[k]; }
            T.c0[id] = 0xffffffffu;
            T.size[id] = 1;
            continue;
        }
        const uint32_t mask = 1u << bit;
        uint32_t lo = 0, hi = count - 1;
        while (lo + 1 != hi) {  // hlbvh.rs:293-316
            const uint32_t mid = (lo + hi) / 2;
            if (((code[first + lo] ^ code[first + mid]) & mask) == 0) lo = mid; else hi = mid;
        }
        const uint32_t c = atomicAdd(T.counter, 2u);
        T.c0[id] = c;
        const uint32_t tre = T.tre[id];
        T.first[c] = first; T.count[c] = hi; T.bit[c] = bit - 1; T.tre[c] = tre;
        T.first[c + 1] = first + hi; T.count[c + 1] = count - hi; T.bit[c + 1] = bit - 1; T.tre[c + 1] = tre;
    }
}
__global__ void k_hlp_mark(HlLevels T, int depth) { T.lvl[depth + 2] = *T.counter; }
__global__ void __launch_bounds__(256) k_hlp_up(HlLevels T, int depth) {
    const uint32_t lo_id = T.lvl[depth], hi_id = T.lvl[depth + 1];
    for (uint32_t id = lo_id + blockIdx.x * blockDim.x + threadIdx.x; id < hi_id; id += gridDim.x * blockDim.x) {
        const uint32_t c = T.c0[id];
        if (c == 0xffffffffu) continue;
        T.size[id] = 1 + T.size[c] + T.size[c + 1];
        const float *a = T.box + 6 * (size_t)c, *b = a + 6;
        float* o = T.box + 6 * (size_t)id;
        for (int k = 0; k < 3; ++k) { o[k] = fmin_ref(a[k], b[k]); o[3 + k] = fmax_ref(a[3 + k], b[3 + k]); }
    }
}
__global__ void __launch_bounds__(256) k_hlp_down(HlLevels T, int depth) {
    const uint32_t lo_id = T.lvl[depth], hi_id = T.lvl[depth + 1];
    for (uint32_t id = lo_id + blockIdx.x * blockDim.x + threadIdx.x; id < hi_id; id += gridDim.x * blockDim.x) {
        if (depth == 0) T.pos[id] = 0;
        const uint32_t c = T.c0[id];
        if (c == 0xffffffffu) continue;
        const uint32_t me = T.pos[id];
        T.pos[c] = me + 1;
        T.pos[c + 1] = me + 1 + T.size[c];
    }
}
__global__ void __launch_bounds__(256) k_hlp_emit(HlLevels T, const uint32_t* __restrict__ starts, uint32_t n_treelets, b200pt_bvh_node* __restrict__ tmp, uint32_t* __restrict__ sizes) {
    const uint32_t total = *T.counter;
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
        b200pt_bvh_node* out = tmp + 2 * (size_t)starts[T.tre[id]] + T.pos[id];
        const float* bx = T.box + 6 * (size_t)id;
        Box b;
        for (int k = 0; k < 3; ++k) { b.lo[k] = bx[k]; b.hi[k] = bx[3 + k]; }
        const uint32_t c = T.c0[id];
        if (c == 0xffffffffu) store_node(out, b, T.first[id], T.count[id], 0);
        else store_node(out, b, T.pos[c + 1], 0, (uint32_t)(T.bit[id] % 3));
        if (id < n_treelets) sizes[id] = T.size[id];
    }
}
__global__ void k_hl_gather_roots(const b200pt_bvh_node* __restrict__ tmp, const uint32_t* __restrict__ starts, uint32_t n_treelets, b200pt_bvh_node* __restrict__ roots) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_treelets) roots[t] = tmp[2 * (size_t)starts[t]];
}
// final placement: one warp per treelet block (second-child indices rebased), one thread per upper node
__global__ void __launch_bounds__(256) k_hl_emit(const b200pt_bvh_node* __restrict__ tmp, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ sizes,
                                                  const long long* __restrict__ base, uint32_t n_treelets, b200pt_bvh_node* __restrict__ out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_treelets) return;
    const uint32_t size = sizes[w], b = (uint32_t)base[w];
    const float4* src = reinterpret_cast<const float4*>(tmp + 2 * (size_t)starts[w]);
    float4* dst = reinterpret_cast<float4*>(out + b);
    for (uint32_t q = lane; q < 2 * size; q += 32) {
        float4 v = src[q];
        if ((q & 1) && (__float_as_uint(v.w) & 0xffffu) == 0) v.z = __uint_as_float(__float_as_uint(v.z) + b);
        dst[q] = v;
    }
}
__global__ void k_hl_emit_upper(const b200pt_bvh_node* __restrict__ upper, const long long* __restrict__ index, uint32_t n_upper, b200pt_bvh_node* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_upper) out[index[k]] = upper[k];
}

// Build scratch: one cached device allocation, grown on demand and carved up by a bump allocator, so that a build
// costs no cudaMalloc / cudaFree (they dominated: 40 ms of a 46 ms build of 1 M triangles).  Builds are serialised by
// the mutex; b200pt_bvh_build_release() gives the memory back.
struct Workspace {
    std::mutex mu;
    char* base = nullptr;
    size_t cap = 0;
};
Workspace g_ws_all[65];  // one cached scratch arena per device (the last slot is never filled: no device bound)
inline Workspace& ws() { const int d = b2::current_device(); return g_ws_all[d >= 0 && d < 64 ? d : 64]; }

struct Arena {
    char* base;
    size_t cap, used = 0;
    template <typename T> T* take(size_t count) {
        used = (used + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + used);
        used += std::max<size_t>(count, 1) * sizeof(T);
        return used <= cap ? p : nullptr;
    }
};
inline size_t padded(size_t bytes) { return ((bytes + 255) & ~(size_t)255) + 256; }

size_t build_scratch_bytes(uint32_t n) {
    const size_t nb = ((size_t)n + kBlock - 1) / kBlock;
    return padded(sizeof(Item) * n) + padded(4 * nb * kBlock) + padded(sizeof(ANode) * ((size_t)n / 8 + 4096)) + padded(sizeof(Bins) * ((size_t)n / kSmall + 2)) +
           padded(sizeof(Ctl)) + padded(8 * (size_t)n) + padded(n) + padded(4 * nb) + 5 * padded(4 * (size_t)n) + padded(64 * (size_t)n) + padded(4 * (nb / 1024 + 2));
}

int workspace_reserve(size_t bytes) {  // caller holds ws().mu
    if (ws().cap >= bytes) return B200PT_OK;
    if (ws().base) { cudaFree(ws().base); ws().base = nullptr; ws().cap = 0; }
    void* q = nullptr;
    B2_CUDA(cudaMalloc(&q, bytes));
    ws().base = static_cast<char*>(q);
    ws().cap = bytes;
    return B200PT_OK;
}

inline unsigned blocks(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

namespace b2 {

// d_prim_bounds: n x 6 floats (device); d_nodes: room for 2n-1 nodes (device); d_ordered: n indices (device).
int build_in_arena(Arena& A, const float* d_prim_bounds, int64_t n64, int max_prims, b200pt_bvh_node* d_nodes, int64_t* n_nodes_out,
                          uint32_t* d_ordered, cudaStream_t st) {
    const uint32_t n = (uint32_t)n64;
    max_prims &= 0xff;  // reference stores it as u8 (mod.rs:357)
    const uint32_t pool_cap = n / 8 + 4096;
    const uint32_t level_cap = n / kSmall + 2;
    const uint32_t n_blocks = blocks(n, kBlock);

    Item* items = A.take<Item>(n);
    int32_t* seg = A.take<int32_t>((size_t)n_blocks * kBlock);
    ANode* nodes = A.take<ANode>(pool_cap);
    Bins* bins = A.take<Bins>(level_cap);
    Ctl* ctl = A.take<Ctl>(1);
    uint2* small_list = A.take<uint2>(n);
    uint8_t* pred = A.take<uint8_t>(n);
    uint32_t* block_sum = A.take<uint32_t>(n_blocks);
    const uint32_t n_chunks = (n_blocks + 1023) / 1024;
    uint32_t* chunk_sums = A.take<uint32_t>(n_chunks + 1);
    uint32_t* T = A.take<uint32_t>(n);
    uint32_t* list_f = A.take<uint32_t>(n);
    uint32_t* list_t = A.take<uint32_t>(n);
    uint32_t* small_size = A.take<uint32_t>(n);
    uint32_t* small_dfs = A.take<uint32_t>(n);
    b200pt_bvh_node* tmp = A.take<b200pt_bvh_node>(2 * (size_t)n);
    if (!tmp) { b200pt_set_error("b200pt_bvh_build_sah_device: internal error: scratch arena too small"); return B200PT_ERR_INVALID; }

    int64_t launches = 0;
    k_make_items<<<n_blocks, kBlock, 0, st>>>(d_prim_bounds, n, items, seg, n > (uint32_t)kSmall ? 0 : -1);
    k_init_root<<<1, 1, 0, st>>>(nodes, ctl, small_list, n);
    launches += 2;

    std::vector<uint32_t> level_first;  // first ANode of each level; level_first.back() = end
    Ctl h{};
    uint32_t first = 0, end = n > (uint32_t)kSmall ? 1u : 0u;
    level_first.push_back(0);
    while (end > first) {
        const uint32_t count = end - first;
        if (count > level_cap) { b200pt_set_error("b200pt_bvh_build_sah_device: level scratch overflow"); return B200PT_ERR_INVALID; }
        k_level_reduce<<<n_blocks, kBlock, 0, st>>>(items, seg, nodes, n);
        k_level_decide1<<<blocks(count, 128), 128, 0, st>>>(nodes, first, count, bins, ctl);
        k_level_bin<<<n_blocks, kBlock, 0, st>>>(items, seg, nodes, first, bins, n);
        k_level_decide2<<<blocks(count, 64), 64, 0, st>>>(nodes, first, count, bins, ctl, small_list, pool_cap, max_prims);
        k_level_pred<<<n_blocks, kBlock, 0, st>>>(items, seg, nodes, pred, block_sum, n);
        if (n_blocks > 4096) {  // chunked scan (see k_scan_chunks): the one-block scan costs 38 us per level at 10 M primitives
            k_scan_chunk_sums<<<n_chunks, 1024, 0, st>>>(block_sum, n_blocks, chunk_sums);
            k_scan_blocks<<<1, 1024, 0, st>>>(chunk_sums, n_chunks);
            k_scan_chunks<<<n_chunks, 1024, 0, st>>>(block_sum, n_blocks, chunk_sums);
        } else k_scan_blocks<<<1, 1024, 0, st>>>(block_sum, n_blocks);
        k_scan_apply<<<n_blocks, kBlock, 0, st>>>(pred, block_sum, T, n);
        k_level_pairs<<<n_blocks, kBlock, 0, st>>>(seg, nodes, pred, T, list_f, list_t, n);
        k_level_swap<<<n_blocks, kBlock, 0, st>>>(items, seg, nodes, T, list_f, list_t, n);
        launches += 9;
        B2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (h.error) break;
        first = end;
        end = h.n_anodes;
        level_first.push_back(first);
    }
    if (!h.error) {
        B2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
    }
    if (!h.error && h.n_small > 0) {
        k_small_build<<<blocks(h.n_small, 128), 128, 0, st>>>(items, small_list, h.n_small, max_prims, tmp, small_size, ctl);
        launches += 1;
        B2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
    }
    g_launches.fetch_add(launches);
    if (h.error) {
        b200pt_set_error(h.error == kErrPool        ? "b200pt_bvh_build_sah_device: node pool overflow (tree too unbalanced for the device builder; use the host builder)"
                         : h.error == kErrEmptySide ? "b200pt_bvh_build_sah: SAH partition produced an empty side (reference panics here)"
                                                    : "b200pt_bvh_build_sah: leaf with >= 65536 primitives (reference asserts)");
        return h.error == kErrPool ? B200PT_ERR_UNSUPPORTED : B200PT_ERR_INVALID;
    }
    launches = 0;
    const int n_levels = (int)level_first.size() - 1;  // levels [level_first[l], level_first[l+1])
    level_first.push_back(h.n_anodes);
    for (int l = n_levels; l >= 0; --l) {
        const uint32_t f = level_first[l], c = level_first[l + 1] - f;
        if (c == 0) continue;
        k_finish_up<<<blocks(c, 128), 128, 0, st>>>(nodes, f, c, items, tmp, small_size);
        ++launches;
    }
    uint32_t total = 0;
    if (h.n_anodes > 0) {
        for (int l = 0; l <= n_levels; ++l) {
            const uint32_t f = level_first[l], c = level_first[l + 1] - f;
            if (c == 0) continue;
            k_finish_down<<<blocks(c, 128), 128, 0, st>>>(nodes, f, c, small_size, small_dfs, d_nodes);
            ++launches;
        }
        ANode root;
        B2_CUDA(cudaMemcpyAsync(&root, nodes, sizeof(ANode), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        total = root.size;
    } else {
        B2_CUDA(cudaMemsetAsync(small_dfs, 0, sizeof(uint32_t), st));
        B2_CUDA(cudaMemcpyAsync(&total, small_size, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (h.n_small > 0) {
        k_emit_small<<<blocks((uint64_t)h.n_small * 32, 256), 256, 0, st>>>(small_list, h.n_small, tmp, small_size, small_dfs, d_nodes);
        ++launches;
    }
    k_ordered<<<n_blocks, kBlock, 0, st>>>(items, n, d_ordered);
    ++launches;
    g_launches.fetch_add(launches);
    B2_CUDA(cudaStreamSynchronize(st));
    B2_CUDA(cudaGetLastError());
    *n_nodes_out = total;
    return B200PT_OK;
}

int bvh_build_sah_device(const float* d_prim_bounds, int64_t n, int max_prims, b200pt_bvh_node* d_nodes, int64_t* n_nodes_out, uint32_t* d_ordered,
                         cudaStream_t st) {
    *n_nodes_out = 0;
    if (n == 0) return B200PT_OK;
    if (n >= (1LL << 31)) { b200pt_set_error("b200pt_bvh_build_sah_device: more than 2^31-1 primitives"); return B200PT_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(ws().mu);
    if (int rc = workspace_reserve(build_scratch_bytes((uint32_t)n))) return rc;
    Arena A{ws().base, ws().cap};
    return build_in_arena(A, d_prim_bounds, n, max_prims, d_nodes, n_nodes_out, d_ordered, st);
}

}  // namespace b2

namespace b2 {

static size_t hlbvh_scratch_bytes(uint32_t n) {
    const size_t nb = ((size_t)n + kBlock - 1) / kBlock;
    return 4 * padded(4 * (size_t)n) + padded(4 * 64 * nb) + padded(n) + padded(4 * nb) + padded(4 * (size_t)n) + padded(4 * (size_t)n) + padded(64 * (size_t)n) +
           padded(4 * 8192) + padded(32 * 8192) * 2 + padded(8 * 8192) * 2 + padded(sizeof(Ctl)) + padded(64) +
           7 * padded(4 * (2 * (size_t)n + 8)) + padded(24 * (2 * (size_t)n + 8)) + 2 * padded(4 * 32) +  // level-parallel treelet state (HlLevels)
           padded(4 * (64 * nb / 1024 + 2));                                                              // chunk sums of the histogram scan
}

int bvh_build_hlbvh_device(const float* d_prim_bounds, int64_t n64, int max_prims, b200pt_bvh_node* d_nodes, int64_t* n_nodes_out, uint32_t* d_ordered,
                           cudaStream_t st) {
    *n_nodes_out = 0;
    if (n64 == 0) return B200PT_OK;
    if (n64 >= (1LL << 31)) { b200pt_set_error("b200pt_bvh_build_hlbvh_device: more than 2^31-1 primitives"); return B200PT_ERR_INVALID; }
    const uint32_t n = (uint32_t)n64;
    max_prims &= 0xff;
    std::lock_guard<std::mutex> lock(ws().mu);
    // B200PT_HLBVH_TIMING=1: host-side wall clock of the build's phases on stderr (where a build's time goes besides its kernels)
    static const bool timing = [] { const char* e = std::getenv("B200PT_HLBVH_TIMING"); return e && e[0] == '1'; }();
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        std::fprintf(stderr, "[hlbvh] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    if (int rc = workspace_reserve(hlbvh_scratch_bytes(n))) return rc;
    lap("workspace");
    Arena A{ws().base, ws().cap};
    const uint32_t n_blocks = blocks(n, kBlock);
    uint32_t* code[2] = {A.take<uint32_t>(n), A.take<uint32_t>(n)};
    uint32_t* val[2] = {A.take<uint32_t>(n), A.take<uint32_t>(n)};
    uint32_t* hist = A.take<uint32_t>(64 * (size_t)n_blocks);
    const uint32_t n_chunks = blocks(64 * (uint64_t)n_blocks, 1024);
    uint32_t* chunk_sums = A.take<uint32_t>(n_chunks);
    uint8_t* flag = A.take<uint8_t>(n);
    uint32_t* block_sum = A.take<uint32_t>(n_blocks);
    uint32_t* T = A.take<uint32_t>(n);
    uint32_t* starts = A.take<uint32_t>(n);
    b200pt_bvh_node* tmp = A.take<b200pt_bvh_node>(2 * (size_t)n);
    uint32_t* sizes = A.take<uint32_t>(8192);
    b200pt_bvh_node* d_roots = A.take<b200pt_bvh_node>(8192);
    b200pt_bvh_node* d_upper = A.take<b200pt_bvh_node>(8192);
    long long* d_upper_index = A.take<long long>(8192);
    long long* d_base = A.take<long long>(8192);
    Ctl* ctl = A.take<Ctl>(1);
    uint32_t* acc = A.take<uint32_t>(6);
    if (!acc) { b200pt_set_error("b200pt_bvh_build_hlbvh_device: internal error: scratch arena too small"); return B200PT_ERR_INVALID; }

    int64_t launches = 0;
    const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    B2_CUDA(cudaMemcpyAsync(acc, init, sizeof(init), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemsetAsync(ctl, 0, sizeof(Ctl), st));
    k_hl_bounds<<<n_blocks, kBlock, 0, st>>>(d_prim_bounds, n, acc);           // hlbvh.rs:42
    k_hl_morton<<<n_blocks, kBlock, 0, st>>>(d_prim_bounds, n, acc, code[0], val[0]);  // hlbvh.rs:104-141
    launches += 2;
    int cur = 0;
    for (int pass = 0; pass < 5; ++pass) {  // morton.rs:60-100
        k_hl_hist<<<n_blocks, kBlock, 0, st>>>(code[cur], n, 6 * pass, hist, n_blocks);
        k_scan_chunk_sums<<<n_chunks, 1024, 0, st>>>(hist, 64 * n_blocks, chunk_sums);
        k_scan_blocks<<<1, 1024, 0, st>>>(chunk_sums, n_chunks);
        k_scan_chunks<<<n_chunks, 1024, 0, st>>>(hist, 64 * n_blocks, chunk_sums);
        k_hl_scatter<<<n_blocks, kBlock, 0, st>>>(code[cur], val[cur], n, 6 * pass, hist, n_blocks, code[cur ^ 1], val[cur ^ 1]);
        launches += 5;
        cur ^= 1;
    }
    k_hl_flags<<<n_blocks, kBlock, 0, st>>>(code[cur], n, flag, block_sum);
    k_scan_blocks<<<1, 1024, 0, st>>>(block_sum, n_blocks);
    k_scan_apply<<<n_blocks, kBlock, 0, st>>>(flag, block_sum, T, n);
    k_hl_starts<<<n_blocks, kBlock, 0, st>>>(flag, T, n, starts);
    launches += 4;
    uint32_t last[2];  // T[n-1] + flag[n-1] = number of treelets
    uint8_t last_flag;
    B2_CUDA(cudaMemcpyAsync(&last[0], T + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&last_flag, flag + (n - 1), 1, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const uint32_t n_treelets = last[0] + last_flag;
    lap("morton + sort + starts");
    if (n_treelets == 0 || n_treelets > 4096) { b200pt_set_error("b200pt_bvh_build_hlbvh_device: internal error: treelet count"); return B200PT_ERR_INVALID; }
    static const bool serial_treelets = [] { const char* e = std::getenv("B200PT_HLBVH_TREELETS"); return e && std::strcmp(e, "serial") == 0; }();  // A/B: one thread per treelet
    if (serial_treelets) k_hl_treelets<<<blocks(n_treelets, 64), 64, 0, st>>>(d_prim_bounds, code[cur], val[cur], n, starts, n_treelets, max_prims, tmp, sizes, ctl);
    else {
        const size_t cap = 2 * (size_t)n + 8;
        HlLevels L;
        L.first = A.take<uint32_t>(cap); L.count = A.take<uint32_t>(cap); L.c0 = A.take<uint32_t>(cap); L.size = A.take<uint32_t>(cap);
        L.pos = A.take<uint32_t>(cap); L.tre = A.take<uint32_t>(cap); L.bit = A.take<int32_t>(cap); L.box = A.take<float>(6 * cap);
        L.counter = A.take<uint32_t>(32); L.lvl = A.take<uint32_t>(32);
        if (!L.lvl) { b200pt_set_error("b200pt_bvh_build_hlbvh_device: internal error: scratch arena too small"); return B200PT_ERR_INVALID; }
        const DevCtx* dc = dev_ctx(current_device());
        const int grid = (dc ? dc->sm_count : 148) * 8;
        const int kDepths = 19;  // the split bit starts at 17 and falls by at least one per depth: depth 18 holds leaves only
        k_hlp_init<<<blocks(n_treelets, 128), 128, 0, st>>>(L, starts, n_treelets, n);
        for (int d = 0; d < kDepths; ++d) {
            k_hlp_level<<<grid, 256, 0, st>>>(L, d, d_prim_bounds, code[cur], val[cur], max_prims, ctl);
            k_hlp_mark<<<1, 1, 0, st>>>(L, d);
        }
        for (int d = kDepths - 1; d >= 0; --d) k_hlp_up<<<grid, 256, 0, st>>>(L, d);
        for (int d = 0; d < kDepths; ++d) k_hlp_down<<<grid, 256, 0, st>>>(L, d);
        k_hlp_emit<<<grid, 256, 0, st>>>(L, starts, n_treelets, tmp, sizes);
        launches += 2 + 4 * kDepths;
    }
    k_hl_gather_roots<<<blocks(n_treelets, 128), 128, 0, st>>>(tmp, starts, n_treelets, d_roots);
    launches += 2;
    lap("treelets");
    std::vector<b200pt_bvh_node> roots(n_treelets), upper(n_treelets);
    std::vector<uint32_t> h_sizes(n_treelets);
    std::vector<int64_t> upper_index(n_treelets), base(n_treelets);
    Ctl h{};
    B2_CUDA(cudaMemcpyAsync(roots.data(), d_roots, n_treelets * sizeof(b200pt_bvh_node), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(h_sizes.data(), sizes, n_treelets * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    g_launches.fetch_add(launches);
    if (h.error) { b200pt_set_error("b200pt_bvh_build_hlbvh: leaf with >= 65536 primitives (reference asserts)"); return B200PT_ERR_INVALID; }
    lap("read-back");
    int64_t n_upper = 0, total = 0;
    if (int rc = b200pt_hlbvh_upper_layout(roots.data(), h_sizes.data(), n_treelets, upper.data(), upper_index.data(), &n_upper, base.data(), &total)) return rc;
    lap("upper SAH tree (host)");
    B2_CUDA(cudaMemcpyAsync(d_base, base.data(), n_treelets * 8, cudaMemcpyHostToDevice, st));
    if (n_upper > 0) {
        B2_CUDA(cudaMemcpyAsync(d_upper, upper.data(), (size_t)n_upper * sizeof(b200pt_bvh_node), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_upper_index, upper_index.data(), (size_t)n_upper * 8, cudaMemcpyHostToDevice, st));
        k_hl_emit_upper<<<blocks((uint64_t)n_upper, 128), 128, 0, st>>>(d_upper, d_upper_index, (uint32_t)n_upper, d_nodes);
    }
    k_hl_emit<<<blocks((uint64_t)n_treelets * 32, 256), 256, 0, st>>>(tmp, starts, sizes, d_base, n_treelets, d_nodes);
    B2_CUDA(cudaMemcpyAsync(d_ordered, val[cur], (size_t)n * 4, cudaMemcpyDeviceToDevice, st));  // leaves take their primitives in sorted order
    g_launches.fetch_add(n_upper > 0 ? 2 : 1);
    B2_CUDA(cudaStreamSynchronize(st));  // the host vectors above go out of scope
    B2_CUDA(cudaGetLastError());
    lap("emit");
    *n_nodes_out = total;
    return B200PT_OK;
}

}  // namespace b2

extern "C" int b200pt_bvh_build_hlbvh_device(const float* d_prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* d_nodes_out,
                                             int64_t* n_nodes_out, uint32_t* d_ordered_out, void* stream) {
    if (int rc = b2::require_device()) return rc;
    if (n < 0 || !n_nodes_out || (n > 0 && (!d_prim_bounds || !d_nodes_out || !d_ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_hlbvh_device: invalid argument");
        return B200PT_ERR_INVALID;
    }
    return b2::bvh_build_hlbvh_device(d_prim_bounds, n, max_prims_in_node, d_nodes_out, n_nodes_out, d_ordered_out, (cudaStream_t)stream);
}

// Same signature and results as b200pt_bvh_build_hlbvh, built on the GPU: host buffers in, host buffers out.
extern "C" int b200pt_bvh_build_hlbvh_gpu(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                                          int64_t* n_nodes_out, uint32_t* ordered_out) {
    if (int rc = b2::require_device()) return rc;
    if (n < 0 || !n_nodes_out || (n > 0 && (!prim_bounds || !nodes_out || !ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_hlbvh_gpu: invalid argument");
        return B200PT_ERR_INVALID;
    }
    *n_nodes_out = 0;
    if (n == 0) return B200PT_OK;
    float* d_bounds = nullptr; b200pt_bvh_node* d_nodes = nullptr; uint32_t* d_ordered = nullptr;
    auto free_all = [&]() { cudaFree(d_bounds); cudaFree(d_nodes); cudaFree(d_ordered); };
    cudaError_t e;
    if ((e = cudaMalloc(&d_bounds, 6 * (size_t)n * sizeof(float))) != cudaSuccess || (e = cudaMalloc(&d_nodes, 2 * (size_t)n * sizeof(b200pt_bvh_node))) != cudaSuccess ||
        (e = cudaMalloc(&d_ordered, (size_t)n * 4)) != cudaSuccess || (e = cudaMemcpy(d_bounds, prim_bounds, 6 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
        free_all();
        return b2::cuda_fail(e, "b200pt_bvh_build_hlbvh_gpu");
    }
    int rc = b2::bvh_build_hlbvh_device(d_bounds, n, max_prims_in_node, d_nodes, n_nodes_out, d_ordered, 0);
    if (!rc && ((e = cudaMemcpy(nodes_out, d_nodes, (size_t)*n_nodes_out * sizeof(b200pt_bvh_node), cudaMemcpyDeviceToHost)) != cudaSuccess ||
                (e = cudaMemcpy(ordered_out, d_ordered, (size_t)n * 4, cudaMemcpyDeviceToHost)) != cudaSuccess))
        rc = b2::cuda_fail(e, "b200pt_bvh_build_hlbvh_gpu");
    free_all();
    return rc;
}

// ================================================================ BVHAccel::new without leaving the device
// Triangles already in HBM -> bounds -> GPU SAH build -> the traversal records (64-byte two-box nodes, 64-byte leaf-order
// triangles): what accel_build_device (b200pt_api.cu) does on the host, as three kernels.  The pre-order index of an
// interior node among the interior nodes ("wide index") is a prefix sum over the node array.
namespace {

__global__ void __launch_bounds__(kBlock) k_acc_flag_interior(const b200pt_bvh_node* __restrict__ nodes, uint32_t n, uint8_t* __restrict__ flag,
                                                               uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t warp_sum[kBlock / 32];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t p = 0;
    if (i < n) { p = nodes[i].n_primitives == 0 ? 1u : 0u; flag[i] = (uint8_t)p; }
    const uint32_t c = __popc(__ballot_sync(kFull, p));
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kBlock / 32; ++w) t += warp_sum[w]; block_sum[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(kBlock) k_acc_wide(const b200pt_bvh_node* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ T, float4* __restrict__ wide,
                                                      float4* __restrict__ tris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const b200pt_bvh_node nd = nodes[i];
    if (nd.n_primitives != 0) {  // leaf: its first triangle record carries the leaf's primitive count
        tris[(size_t)nd.offset * 4 + 2].w = __uint_as_float((uint32_t)nd.n_primitives);
        return;
    }
    const b200pt_bvh_node c0 = nodes[i + 1], c1 = nodes[nd.offset];
    const int k0 = c0.n_primitives == 0 ? (int)T[i + 1] : ~(int)c0.offset, k1 = c1.n_primitives == 0 ? (int)T[nd.offset] : ~(int)c1.offset;
    float4* q = wide + 4 * (size_t)T[i];
    q[0] = make_float4(c0.bounds[0], c0.bounds[1], c0.bounds[2], c0.bounds[3]);
    q[1] = make_float4(c0.bounds[4], c0.bounds[5], c1.bounds[0], c1.bounds[1]);
    q[2] = make_float4(c1.bounds[2], c1.bounds[3], c1.bounds[4], c1.bounds[5]);
    q[3] = make_float4(__int_as_float(k0), __int_as_float(k1), __int_as_float((int)nd.axis), 0.0f);
}
__global__ void __launch_bounds__(kBlock) k_acc_tris(const float* __restrict__ tri_verts, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ ordered,
                                                      uint32_t n, float4* __restrict__ tris) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t p = ordered[j];
    const float* v = tri_verts + 9 * (size_t)p;
    const uint32_t fl = flags ? flags[p] : 0u;
    tris[(size_t)j * 4 + 0] = make_float4(v[0], v[1], v[2], v[3]);
    tris[(size_t)j * 4 + 1] = make_float4(v[4], v[5], v[6], v[7]);
    tris[(size_t)j * 4 + 2] = make_float4(v[8], __uint_as_float(p), __uint_as_float(fl), __uint_as_float(0u));
    tris[(size_t)j * 4 + 3] = make_float4(0.0f - 1.0f, 0.0f - 1.0f, 1.0f - 1.0f, 0.0f - 1.0f);  // default uvs (0,0) (1,0) (1,1): uv0 - uv2, uv1 - uv2
}

}  // namespace

extern "C" int b200pt_accel_create_device(const float* d_tri_verts, int64_t n_prims, const uint32_t* d_prim_flags, int max_prims_in_node, void* stream,
                                          b200pt_accel** out) {
    if (!out) { b200pt_set_error("b200pt_accel_create_device: out is null"); return B200PT_ERR_INVALID; }
    *out = nullptr;
    if (int rc = b2::require_device()) return rc;
    if (n_prims <= 0 || !d_tri_verts || n_prims >= (1LL << 30)) { b200pt_set_error("b200pt_accel_create_device: invalid argument"); return B200PT_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n = (uint32_t)n_prims;
    {   // stream-ordered allocations from a pool that keeps its memory: a rebuild then costs no cudaMalloc
        static std::once_flag once[64];
        const int dev = b2::current_device();
        std::call_once(once[dev & 63], [dev] {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        });
    }
    b200pt_accel* a = new b200pt_accel();
    b2::AccelImpl& A = a->impl;
    std::memset(&A.dev, 0, sizeof(A.dev));
    A.device = A.dev.device = b2::current_device();
    auto fail = [&](int rc) { b2::accel_free_device(&A); delete a; return rc; };
    float* d_bounds = nullptr; b200pt_bvh_node* d_nodes = nullptr; uint32_t* d_ordered = nullptr;
    uint8_t* flag = nullptr; uint32_t *block_sum = nullptr, *T = nullptr;
    auto free_tmp = [&]() { for (void* q : {(void*)d_bounds, (void*)d_nodes, (void*)d_ordered, (void*)flag, (void*)block_sum, (void*)T}) if (q) cudaFreeAsync(q, st); };
    auto check = [&](cudaError_t e, const char* what) { return e == cudaSuccess ? 0 : b2::cuda_fail(e, what); };
    int rc = 0;
    if ((rc = check(cudaMallocAsync(&d_bounds, 6 * (size_t)n * sizeof(float), st), "cudaMalloc(bounds)")) || (rc = check(cudaMallocAsync(&d_nodes, 2 * (size_t)n * sizeof(b200pt_bvh_node), st), "cudaMalloc(nodes)")) ||
        (rc = check(cudaMallocAsync(&d_ordered, (size_t)n * 4, st), "cudaMalloc(ordered)"))) { free_tmp(); return fail(rc); }
    k_tri_bounds<<<blocks(n, 256), 256, 0, st>>>(d_tri_verts, n, d_bounds);
    b2::g_launches.fetch_add(1);
    int64_t n_nodes = 0;
    if ((rc = b2::bvh_build_sah_device(d_bounds, n, max_prims_in_node, d_nodes, &n_nodes, d_ordered, st))) { free_tmp(); return fail(rc); }
    const uint32_t nn = (uint32_t)n_nodes, nb = blocks(nn, kBlock);
    if ((rc = check(cudaMallocAsync(&flag, nn, st), "cudaMalloc")) || (rc = check(cudaMallocAsync(&block_sum, (size_t)nb * 4, st), "cudaMalloc")) || (rc = check(cudaMallocAsync(&T, (size_t)nn * 4, st), "cudaMalloc")) ||
        (rc = check(cudaMallocAsync(&A.d_tris, (size_t)n * 4 * sizeof(float4), st), "cudaMalloc(tris)")) || (rc = check(cudaMallocAsync(&A.d_ref, (size_t)nn * sizeof(b200pt_bvh_node), st), "cudaMalloc(ref nodes)"))) { free_tmp(); return fail(rc); }
    k_acc_flag_interior<<<nb, kBlock, 0, st>>>(d_nodes, nn, flag, block_sum);
    k_scan_blocks<<<1, 1024, 0, st>>>(block_sum, nb);
    k_scan_apply<<<nb, kBlock, 0, st>>>(flag, block_sum, T, nn);
    uint32_t last_T = 0; uint8_t last_flag = 0;
    b200pt_bvh_node root;
    cudaMemcpyAsync(&last_T, T + (nn - 1), 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last_flag, flag + (nn - 1), 1, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&root, d_nodes, sizeof(root), cudaMemcpyDeviceToHost, st);
    if ((rc = check(cudaStreamSynchronize(st), "accel_create_device"))) { free_tmp(); return fail(rc); }
    const uint32_t n_wide = last_T + last_flag;
    if ((rc = check(cudaMallocAsync(&A.d_wide, (size_t)std::max<uint32_t>(n_wide, 1) * 4 * sizeof(float4), st), "cudaMalloc(wide)"))) { free_tmp(); return fail(rc); }
    k_acc_tris<<<blocks(n, kBlock), kBlock, 0, st>>>(d_tri_verts, d_prim_flags, d_ordered, n, A.d_tris);
    k_acc_wide<<<nb, kBlock, 0, st>>>(d_nodes, nn, T, A.d_wide, A.d_tris);
    cudaMemcpyAsync(A.d_ref, d_nodes, (size_t)nn * sizeof(b200pt_bvh_node), cudaMemcpyDeviceToDevice, st);
    b2::g_launches.fetch_add(5);
    if ((rc = check(cudaStreamSynchronize(st), "accel_create_device")) || (rc = check(cudaGetLastError(), "accel_create_device"))) { free_tmp(); return fail(rc); }
    free_tmp();
    A.n_nodes = n_nodes; A.n_prims = n_prims; A.n_wide = n_wide;
    std::memcpy(A.world_bound, root.bounds, 24);
    std::memcpy(A.dev.root_bounds, root.bounds, 24);
    A.dev.root_code = root.n_primitives == 0 ? 0 : ~(int)root.offset;
    A.dev.n_nodes = (int)n_nodes;
    A.dev.n_prims = n_prims;
    A.dev.wide = A.d_wide; A.dev.tris = A.d_tris; A.dev.ref_nodes = A.d_ref;
    *out = a;
    return B200PT_OK;
}

// The LinearBVHNode array / ordered_prims of a device-resident accelerator (parity checks; pass NULL to skip one).
extern "C" int b200pt_accel_download(const b200pt_accel* a, b200pt_bvh_node* nodes_out, int64_t* n_nodes_out, uint32_t* ordered_out) {
    if (!a) { b200pt_set_error("b200pt_accel_download: null accelerator"); return B200PT_ERR_INVALID; }
    if (int rc = b2::use_device(a->impl.device)) return rc;
    if (n_nodes_out) *n_nodes_out = a->impl.n_nodes;
    if (nodes_out && a->impl.n_nodes > 0) B2_CUDA(cudaMemcpy(nodes_out, a->impl.d_ref, (size_t)a->impl.n_nodes * sizeof(b200pt_bvh_node), cudaMemcpyDeviceToHost));
    if (ordered_out && a->impl.n_prims > 0) {  // original index = second word of the third float4 of every leaf-order triangle record
        std::vector<float4> rec((size_t)a->impl.n_prims * 4);
        B2_CUDA(cudaMemcpy(rec.data(), a->impl.d_tris, rec.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < a->impl.n_prims; ++j) std::memcpy(&ordered_out[j], &rec[(size_t)j * 4 + 2].y, 4);
    }
    return B200PT_OK;
}

extern "C" int b200pt_triangle_bounds_device(const float* d_tri_verts, int64_t n, float* d_bounds_out, void* stream) {
    if (int rc = b2::require_device()) return rc;
    if (n <= 0) return B200PT_OK;
    k_tri_bounds<<<blocks((uint64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(d_tri_verts, n, d_bounds_out);
    b2::g_launches.fetch_add(1);
    B2_CUDA(cudaGetLastError());
    return B200PT_OK;
}

extern "C" int b200pt_bvh_build_sah_device(const float* d_prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* d_nodes_out,
                                           int64_t* n_nodes_out, uint32_t* d_ordered_out, void* stream) {
    if (int rc = b2::require_device()) return rc;
    if (n < 0 || !n_nodes_out || (n > 0 && (!d_prim_bounds || !d_nodes_out || !d_ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_sah_device: invalid argument");
        return B200PT_ERR_INVALID;
    }
    return b2::bvh_build_sah_device(d_prim_bounds, n, max_prims_in_node, d_nodes_out, n_nodes_out, d_ordered_out, (cudaStream_t)stream);
}

// Same signature and results as b200pt_bvh_build_sah, built on the GPU: host buffers in, host buffers out.
extern "C" int b200pt_bvh_build_sah_gpu(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                                        int64_t* n_nodes_out, uint32_t* ordered_out) {
    if (int rc = b2::require_device()) return rc;
    if (n < 0 || !n_nodes_out || (n > 0 && (!prim_bounds || !nodes_out || !ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_sah_gpu: invalid argument");
        return B200PT_ERR_INVALID;
    }
    *n_nodes_out = 0;
    if (n == 0) return B200PT_OK;
    if (n >= (1LL << 31)) { b200pt_set_error("b200pt_bvh_build_sah_gpu: more than 2^31-1 primitives"); return B200PT_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(ws().mu);
    if (int rc = workspace_reserve(build_scratch_bytes((uint32_t)n) + padded(24 * (size_t)n) + padded(64 * (size_t)n) + padded(4 * (size_t)n))) return rc;
    Arena A{ws().base, ws().cap};
    float* d_bounds = A.take<float>(6 * (size_t)n);
    b200pt_bvh_node* d_nodes = A.take<b200pt_bvh_node>(2 * (size_t)n);
    uint32_t* d_ordered = A.take<uint32_t>((size_t)n);
    B2_CUDA(cudaMemcpy(d_bounds, prim_bounds, 6 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    int rc = b2::build_in_arena(A, d_bounds, n, max_prims_in_node, d_nodes, n_nodes_out, d_ordered, 0);
    // the device builder sizes its node pool for n / 8 large nodes; a tree more lopsided than that is the same tree on the host
    if (rc == B200PT_ERR_UNSUPPORTED) return b200pt_bvh_build_sah(prim_bounds, n, max_prims_in_node, nodes_out, n_nodes_out, ordered_out);
    if (rc) return rc;
    B2_CUDA(cudaMemcpy(nodes_out, d_nodes, (size_t)*n_nodes_out * sizeof(b200pt_bvh_node), cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(ordered_out, d_ordered, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return B200PT_OK;
}

extern "C" int b200pt_bvh_build_release(void) {
    std::lock_guard<std::mutex> lock(ws().mu);
    if (ws().base) cudaFree(ws().base);
    ws().base = nullptr;
    ws().cap = 0;
    return B200PT_OK;
}
