// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY UNPINNED for this file: no render of the reference
// uses the HLBVH split method (see oracle_math.h for what is pinned by execution).
//
// BVHAccel::new with SplitMethod::HLBVH, restated from accelerators/src/bvh/hlbvh.rs:33-449 and morton.rs:37-120.
//
// Two properties of the reference that this restatement keeps because they change the result:
//  * encode_morton_3 (morton.rs:43-49) feeds `float_to_bits(v.x)` — the IEEE-754 bit pattern of the scaled centroid
//    offset (core/src/pbrt/common.rs:179-187 is a transmute), not its integer value — into left_shift_3.  A release
//    build (debug_assert off) therefore interleaves mantissa bits 0..7 and (bits 8,9 | bits 24,25) of each float: the
//    "Morton" codes are spatially incoherent, the tree is still a valid BVH.  A debug build panics at morton.rs:105.
//  * treelets are emitted by worker threads that bump one shared `ordered_prims_offset` (hlbvh.rs:259), so with more
//    than one thread the position of a treelet's primitives in `ordered_prims` depends on scheduling.  The tree
//    structure and every box do not.  This restatement emits the treelets in order, which is what `--nthreads 1`
//    produces; leaf k's primitives then sit at its range of the Morton-sorted array.
#pragma once
#include <cstring>
#include <vector>
#include "oracle_bvh.h"

namespace orc {

struct MortonPrimitive {  // morton.rs:9-16
    size_t primitive_index;
    uint32_t morton_code;
};

// morton.rs:102-120 (release semantics: the debug_assert is compiled out)
inline uint32_t left_shift_3(uint32_t x) {
    uint32_t x1 = (x == (1u << 10)) ? x - 1 : x;
    x1 = (x1 | (x1 << 16)) & 0x030000FFu;
    x1 = (x1 | (x1 << 8)) & 0x0300F00Fu;
    x1 = (x1 | (x1 << 4)) & 0x030C30C3u;
    x1 = (x1 | (x1 << 2)) & 0x09249249u;
    return x1;
}
inline uint32_t float_to_bits(Float f) {  // core/src/pbrt/common.rs:179-187
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
// morton.rs:43-49
inline uint32_t encode_morton_3(V3 v) {
    return (left_shift_3(float_to_bits(v.z)) << 2) | (left_shift_3(float_to_bits(v.y)) << 1) | left_shift_3(float_to_bits(v.x));
}

// morton.rs:60-100: LSD radix sort, 5 passes of 6 bits (stable)
inline void radix_sort(std::vector<MortonPrimitive>& v) {
    const int kBitsPerPass = 6, kBits = 30, kPasses = kBits / kBitsPerPass, kBuckets = 1 << kBitsPerPass;
    std::vector<MortonPrimitive> temp(v.size());
    for (int pass = 0; pass < kPasses; ++pass) {
        const int low_bit = pass * kBitsPerPass;
        std::vector<MortonPrimitive>& in = (pass & 1) ? temp : v;
        std::vector<MortonPrimitive>& out = (pass & 1) ? v : temp;
        size_t count[kBuckets] = {0}, out_index[kBuckets];
        for (const MortonPrimitive& mp : in) count[(mp.morton_code >> low_bit) & (kBuckets - 1)] += 1;
        out_index[0] = 0;
        for (int i = 1; i < kBuckets; ++i) out_index[i] = out_index[i - 1] + count[i - 1];
        for (const MortonPrimitive& mp : in) out[out_index[(mp.morton_code >> low_bit) & (kBuckets - 1)]++] = mp;
    }
    if (kPasses & 1) v.swap(temp);
}

struct HLBVHBuilder : BVHBuilder {
    std::vector<MortonPrimitive> mp;
    size_t ordered_offset = 0;

    // hlbvh.rs:243-345
    int emit_lbvh(size_t first, size_t n_prims, int bit_index) {
        if (bit_index == -1 || n_prims < (size_t)max_prims_in_node) {
            BVHBuildNode nd;
            nd.bounds = Bounds3();
            nd.children[0] = nd.children[1] = -1;
            nd.split_axis = 0;
            nd.first_prim_offset = ordered_offset;
            ordered_offset += n_prims;
            nd.n_primitives = n_prims;
            for (size_t i = 0; i < n_prims; ++i) {
                size_t pi = mp[first + i].primitive_index;
                ordered[nd.first_prim_offset + i] = (uint32_t)pi;
                nd.bounds = bunion(nd.bounds, info[pi].bounds);
            }
            total_nodes += 1;
            pool.push_back(nd);
            return (int)pool.size() - 1;
        }
        const uint32_t mask = 1u << bit_index;
        if ((mp[first].morton_code & mask) == (mp[first + n_prims - 1].morton_code & mask)) return emit_lbvh(first, n_prims, bit_index - 1);
        size_t search_start = 0, search_end = n_prims - 1;
        while (search_start + 1 != search_end) {
            size_t mid = (search_start + search_end) / 2;
            if ((mp[first + search_start].morton_code & mask) == (mp[first + mid].morton_code & mask)) search_start = mid;
            else search_end = mid;
        }
        const size_t split = search_end;
        int c0 = emit_lbvh(first, split, bit_index - 1);
        int c1 = emit_lbvh(first + split, n_prims - split, bit_index - 1);
        total_nodes += 1;
        BVHBuildNode nd;
        nd.bounds = bunion(pool[c0].bounds, pool[c1].bounds);  // common.rs:150-159
        nd.children[0] = c0;
        nd.children[1] = c1;
        nd.split_axis = bit_index % 3;
        nd.first_prim_offset = 0;
        nd.n_primitives = 0;
        pool.push_back(nd);
        return (int)pool.size() - 1;
    }

    static int upper_bucket(Float centroid, Float lo, Float hi) {  // hlbvh.rs:382-389, 430-436
        Float v = (Float)kBuckets * ((centroid - lo) / (hi - lo));
        int b;
        if (!(v == v) || v <= 0.0f) b = 0;
        else if (v >= 2147483648.0f) b = 0x7fffffff;
        else b = (int)v;
        if (b == kBuckets) b = kBuckets - 1;
        return b;
    }

    // hlbvh.rs:353-449.  Returns -1 when the reference would panic (zero centroid extent or an empty side).
    int build_upper_sah(std::vector<int>& roots, size_t start, size_t end) {
        if (end - start == 1) return roots[start];
        Bounds3 bounds, cb;
        for (size_t i = start; i < end; ++i) bounds = bunion(bounds, pool[roots[i]].bounds);
        for (size_t i = start; i < end; ++i) cb = bunion(cb, (pool[roots[i]].bounds.pmin + pool[roots[i]].bounds.pmax) * 0.5f);
        const int dim = maximum_extent(cb);
        if (cb.pmax[dim] == cb.pmin[dim]) return -1;  // hlbvh.rs:376 assert_ne
        struct Bucket { size_t count = 0; Bounds3 bounds; } buckets[kBuckets];
        for (size_t i = start; i < end; ++i) {
            const Bounds3& rb = pool[roots[i]].bounds;
            int b = upper_bucket((rb.pmin[dim] + rb.pmax[dim]) * 0.5f, cb.pmin[dim], cb.pmax[dim]);
            if (b < 0 || b >= kBuckets) return -1;
            buckets[b].count += 1;
            buckets[b].bounds = bunion(buckets[b].bounds, rb);
        }
        Float cost[kBuckets - 1];
        for (int i = 0; i < kBuckets - 1; ++i) {
            Bounds3 b0, b1;
            size_t c0 = 0, c1 = 0;
            for (int j = 0; j <= i; ++j) { b0 = bunion(b0, buckets[j].bounds); c0 += buckets[j].count; }
            for (int j = i + 1; j < kBuckets; ++j) { b1 = bunion(b1, buckets[j].bounds); c1 += buckets[j].count; }
            cost[i] = 0.125f + ((Float)c0 * surface_area(b0) + (Float)c1 * surface_area(b1)) / surface_area(bounds);
        }
        Float min_cost = cost[0];
        int min_bucket = 0;
        for (int i = 1; i < kBuckets - 1; ++i)
            if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
        // itertools::partition over treelet_roots[start..end] (hlbvh.rs:427-437)
        auto pred = [&](int r) {
            const Bounds3& rb = pool[r].bounds;
            return upper_bucket((rb.pmin[dim] + rb.pmax[dim]) * 0.5f, cb.pmin[dim], cb.pmax[dim]) <= min_bucket;
        };
        size_t split = 0, f = start, b = end;
        while (f < b) {
            size_t front = f++;
            if (!pred(roots[front])) {
                bool found = false;
                while (b > f) {
                    --b;
                    if (pred(roots[b])) { found = true; break; }
                }
                if (!found) break;
                int t = roots[front]; roots[front] = roots[b]; roots[b] = t;
            }
            split += 1;
        }
        const size_t mid = start + split;
        if (!(mid > start && mid < end)) return -1;  // hlbvh.rs:440-441 asserts
        total_nodes += 1;
        int c0 = build_upper_sah(roots, start, mid);
        if (c0 < 0) return -1;
        int c1 = build_upper_sah(roots, mid, end);
        if (c1 < 0) return -1;
        BVHBuildNode nd;
        nd.bounds = bunion(pool[c0].bounds, pool[c1].bounds);
        nd.children[0] = c0;
        nd.children[1] = c1;
        nd.split_axis = dim;
        nd.first_prim_offset = 0;
        nd.n_primitives = 0;
        pool.push_back(nd);
        return (int)pool.size() - 1;
    }
};

// hlbvh.rs:33-98 + mod.rs:43-124.  Returns false where the reference panics.
inline bool bvh_build_hlbvh(const Float* prim_bounds, size_t n, int max_prims_in_node, std::vector<LinearBVHNode>& nodes,
                            std::vector<uint32_t>& ordered, std::vector<uint32_t>* morton_sorted = nullptr) {
    nodes.clear();
    ordered.clear();
    if (n == 0) return true;
    HLBVHBuilder b;
    b.max_prims_in_node = max_prims_in_node;
    b.info.resize(n);
    Bounds3 bounds;
    for (size_t i = 0; i < n; ++i) {
        const Float* pb = prim_bounds + 6 * i;
        b.info[i].primitive_number = i;
        b.info[i].bounds = Bounds3(V3(pb[0], pb[1], pb[2]), V3(pb[3], pb[4], pb[5]));
        b.info[i].centroid = 0.5f * (b.info[i].bounds.pmin + b.info[i].bounds.pmax);
        bounds = bunion(bounds, b.info[i].bounds);  // hlbvh.rs:42: bounds of the primitive BOUNDS, not of the centroids
    }
    b.mp.resize(n);
    const Float morton_scale = (Float)(1 << 10);
    for (size_t i = 0; i < n; ++i) {  // hlbvh.rs:104-141
        V3 v = boffset(bounds, b.info[i].centroid) * morton_scale;
        b.mp[i].primitive_index = b.info[i].primitive_number;
        b.mp[i].morton_code = encode_morton_3(v);
    }
    radix_sort(b.mp);
    if (morton_sorted) {
        morton_sorted->resize(n);
        for (size_t i = 0; i < n; ++i) (*morton_sorted)[i] = b.mp[i].morton_code;
    }
    // treelets: runs of equal top-12 bits (hlbvh.rs:53-69)
    const uint32_t mask = 0x3FFC0000u;
    std::vector<std::pair<size_t, size_t>> treelets;
    for (size_t start = 0, end = 1; end <= n; ++end) {
        if (end == n || ((b.mp[start].morton_code & mask) != (b.mp[end].morton_code & mask))) {
            treelets.emplace_back(start, end - start);
            start = end;
        }
    }
    b.ordered.assign(n, 0);
    b.pool.reserve(2 * n);
    std::vector<int> roots;
    const int first_bit_index = 30 - 1 - 12;  // hlbvh.rs:20
    for (auto& t : treelets) roots.push_back(b.emit_lbvh(t.first, t.second, first_bit_index));
    int root = b.build_upper_sah(roots, 0, roots.size());
    if (root < 0) return false;
    nodes.assign(b.total_nodes, LinearBVHNode());
    uint32_t off = 0;
    b.flatten(root, nodes, &off);
    ordered.swap(b.ordered);
    return true;
}

}  // namespace orc
