// Host-side BVH construction for the B200 path: BVHAccel::new with
// SplitMethod::SAH (reference: accelerators/src/bvh/mod.rs:43-153,
// sah.rs:26-367, common.rs:66-224).
//
// Design (not a port): the reference recurses, allocates one arena node per
// build() call and flattens afterwards.  Here the tree is emitted directly in
// its final depth-first order by an explicit work stack, so the LinearBVHNode
// array is written exactly once and no intermediate tree exists:
//   * a node's index is the number of nodes emitted before it (pre-order), which
//     is what flatten_bvh_tree assigns (mod.rs:126-153);
//   * the second child's index is patched into its parent when that child is
//     popped;
//   * leaves append their primitives to `ordered` in pop order == DFS order.
// Bounds unions are min/max only and therefore exact and order-independent, so
// the 11 split costs are evaluated with one prefix and one suffix sweep over the
// 12 buckets instead of the reference's O(12^2) re-accumulation; the float
// operations that do round (centroid, bucket index, cost, surface area) keep
// the reference's operand order so the tree is structurally identical.
//
// Compiled with -ffp-contract=off (see build.py): no FMA contraction on host.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/b200pt.h"

namespace {

struct Box {
    float lo[3], hi[3];
};
inline float fmin_ref(float a, float b) { return a < b ? a : b; }  // core/src/pbrt/common.rs:83-94
inline float fmax_ref(float a, float b) { return a > b ? a : b; }  // :97-108
inline Box empty_box() {
    const float m = std::numeric_limits<float>::max();  // bounds3.rs:26-29
    return Box{{m, m, m}, {-m, -m, -m}};
}
inline void grow(Box& b, const Box& o) {
    for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_ref(b.lo[k], o.lo[k]); b.hi[k] = fmax_ref(b.hi[k], o.hi[k]); }
}
inline void grow_pt(Box& b, const float* p) {
    for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_ref(b.lo[k], p[k]); b.hi[k] = fmax_ref(b.hi[k], p[k]); }
}
inline float area(const Box& b) {  // bounds3.rs:94-105
    if (b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2]) return 0.0f;
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    float h = dx * dy + dx * dz + dy * dz;
    return h + h;
}
inline int widest_axis(const Box& b) {  // bounds3.rs:122-134
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (dx > dy && dx > dz) return 0;
    return dy > dz ? 1 : 2;
}

struct Item {
    uint32_t prim;
    Box box;
    float c[3];
};

constexpr int kBins = 12;  // sah.rs:11

// (12 * Bounds3::offset(c)[dim]) as usize, 12 -> 11  (sah.rs:305-313, bounds3.rs:153-168)
inline int bin_of(const Box& cb, const Item& it, int dim) {
    float o = it.c[dim] - cb.lo[dim];
    if (cb.hi[dim] > cb.lo[dim]) o /= cb.hi[dim] - cb.lo[dim];
    float v = (float)kBins * o;
    int b = (!(v == v) || v <= 0.0f) ? 0 : (v >= 2147483648.0f ? 0x7fffffff : (int)v);  // Rust saturating cast
    return b == kBins ? kBins - 1 : b;
}

struct Job {
    size_t begin, end;
    int64_t parent;  // node to patch with this node's index as second child, or -1
};

}  // namespace

extern "C" int b200pt_set_error(const char* msg);  // defined in b200pt_api.cu

extern "C" int b200pt_triangle_bounds(const float* tri_verts, int64_t n, float* bounds_out) {
    if ((!tri_verts || !bounds_out) && n > 0) { b200pt_set_error("b200pt_triangle_bounds: null pointer"); return B200PT_ERR_INVALID; }
    for (int64_t i = 0; i < n; ++i) {  // shapes/src/triangle.rs:427-431
        const float* v = tri_verts + 9 * i;
        Box b{{v[0], v[1], v[2]}, {v[0], v[1], v[2]}};
        grow_pt(b, v + 3);
        grow_pt(b, v + 6);
        float* o = bounds_out + 6 * i;
        o[0] = b.lo[0]; o[1] = b.lo[1]; o[2] = b.lo[2]; o[3] = b.hi[0]; o[4] = b.hi[1]; o[5] = b.hi[2];
    }
    return B200PT_OK;
}

extern "C" int b200pt_bvh_build_sah(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                                    int64_t* n_nodes_out, uint32_t* ordered_out) {
    if (n < 0 || !n_nodes_out || (n > 0 && (!prim_bounds || !nodes_out || !ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_sah: invalid argument");
        return B200PT_ERR_INVALID;
    }
    *n_nodes_out = 0;
    if (n == 0) return B200PT_OK;  // mod.rs:47-53: empty accelerator, no nodes
    if (n > 0xfffffff0LL) { b200pt_set_error("b200pt_bvh_build_sah: too many primitives"); return B200PT_ERR_INVALID; }
    max_prims_in_node &= 0xff;  // reference stores it as u8 (mod.rs:357)

    std::vector<Item> items((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* pb = prim_bounds + 6 * i;
        Item& it = items[(size_t)i];
        it.prim = (uint32_t)i;
        for (int k = 0; k < 3; ++k) { it.box.lo[k] = pb[k]; it.box.hi[k] = pb[3 + k]; }
        for (int k = 0; k < 3; ++k) it.c[k] = 0.5f * (it.box.lo[k] + it.box.hi[k]);  // common.rs:86
    }

    int64_t n_nodes = 0, n_ordered = 0;
    std::vector<Job> work;
    work.push_back(Job{0, (size_t)n, -1});
    while (!work.empty()) {
        Job job = work.back();
        work.pop_back();
        const int64_t me = n_nodes++;
        if (job.parent >= 0) nodes_out[job.parent].offset = (uint32_t)me;
        b200pt_bvh_node& node = nodes_out[me];

        Box bb = empty_box();
        for (size_t i = job.begin; i < job.end; ++i) grow(bb, items[i].box);
        for (int k = 0; k < 3; ++k) { node.bounds[k] = bb.lo[k]; node.bounds[3 + k] = bb.hi[k]; }
        node.pad = 0;
        const size_t count = job.end - job.begin;

        auto make_leaf = [&]() {  // sah.rs:187-211, mod.rs:133-141
            node.offset = (uint32_t)n_ordered;
            node.n_primitives = (uint16_t)count;
            node.axis = 0;
            for (size_t i = job.begin; i < job.end; ++i) ordered_out[n_ordered++] = items[i].prim;
        };
        if (count == 1) { make_leaf(); continue; }
        Box cb = empty_box();
        for (size_t i = job.begin; i < job.end; ++i) grow_pt(cb, items[i].c);
        const int dim = widest_axis(cb);
        if (cb.hi[dim] == cb.lo[dim]) {  // sah.rs:61
            if (count >= 65536) { b200pt_set_error("b200pt_bvh_build_sah: leaf with >= 65536 primitives (reference asserts)"); return B200PT_ERR_INVALID; }
            make_leaf();
            continue;
        }

        size_t mid;
        if (count <= 2) {
            // sah.rs:81-83: equal counts; with two primitives whose centroids
            // differ along dim the smaller one comes first.
            mid = (job.begin + job.end) / 2;
            if (items[job.end - 1].c[dim] < items[job.begin].c[dim]) std::swap(items[job.begin], items[job.end - 1]);
        } else {
            size_t bin_count[kBins] = {0};
            Box bin_box[kBins];
            for (int b = 0; b < kBins; ++b) bin_box[b] = empty_box();
            for (size_t i = job.begin; i < job.end; ++i) {
                int b = bin_of(cb, items[i], dim);
                bin_count[b] += 1;
                grow(bin_box[b], items[i].box);
            }
            // prefix / suffix sweeps (exact: unions are min/max, counts are integers)
            Box left[kBins - 1], right[kBins - 1];
            size_t nl[kBins - 1], nr[kBins - 1];
            Box acc = empty_box();
            size_t cnt = 0;
            for (int b = 0; b < kBins - 1; ++b) { grow(acc, bin_box[b]); cnt += bin_count[b]; left[b] = acc; nl[b] = cnt; }
            acc = empty_box();
            cnt = 0;
            for (int b = kBins - 1; b >= 1; --b) { grow(acc, bin_box[b]); cnt += bin_count[b]; right[b - 1] = acc; nr[b - 1] = cnt; }
            const float total_area = area(bb);
            float best = 0.0f;
            int best_bin = 0;
            for (int b = 0; b < kBins - 1; ++b) {  // sah.rs:321-347: first minimum wins
                float cost = 1.0f + ((float)nl[b] * area(left[b]) + (float)nr[b] * area(right[b])) / total_area;
                if (b == 0 || cost < best) { best = cost; best_bin = b; }
            }
            if (count > (size_t)max_prims_in_node || best < (float)count) {  // sah.rs:351
                // itertools::partition: front/back swap partition
                size_t f = job.begin, bk = job.end, split = 0;
                while (f < bk) {
                    size_t front = f++;
                    if (!(bin_of(cb, items[front], dim) <= best_bin)) {
                        bool found = false;
                        while (bk > f) {
                            --bk;
                            if (bin_of(cb, items[bk], dim) <= best_bin) { found = true; break; }
                        }
                        if (!found) break;
                        std::swap(items[front], items[bk]);
                    }
                    ++split;
                }
                mid = job.begin + split;
            } else {
                if (count >= 65536) { b200pt_set_error("b200pt_bvh_build_sah: leaf with >= 65536 primitives (reference asserts)"); return B200PT_ERR_INVALID; }
                make_leaf();
                continue;
            }
        }
        if (mid == job.begin || mid == job.end) {
            // The reference would recurse on an empty range and hit assert_ne!(start, end) (sah.rs:37).
            b200pt_set_error("b200pt_bvh_build_sah: SAH partition produced an empty side (reference panics here)");
            return B200PT_ERR_INVALID;
        }
        node.n_primitives = 0;
        node.axis = (uint8_t)dim;
        node.offset = 0;  // patched when the second child is emitted
        work.push_back(Job{mid, job.end, me});       // second child: popped after the whole first subtree
        work.push_back(Job{job.begin, mid, -1});     // first child: next node, index me + 1
    }
    // An interior node's bounds are union(child0, child1) in the reference (common.rs:150-159), not the fold over
    // its range: same box, but min(a, b) = a < b ? a : b makes the sign of a zero order-dependent.  Children follow
    // their parent in pre-order, so one reverse sweep reproduces the bottom-up construction.
    for (int64_t i = n_nodes - 1; i >= 0; --i) {
        b200pt_bvh_node& nd = nodes_out[i];
        if (nd.n_primitives != 0) continue;
        const b200pt_bvh_node &c0 = nodes_out[i + 1], &c1 = nodes_out[nd.offset];
        for (int k = 0; k < 3; ++k) {
            nd.bounds[k] = fmin_ref(c0.bounds[k], c1.bounds[k]);
            nd.bounds[3 + k] = fmax_ref(c0.bounds[3 + k], c1.bounds[3 + k]);
        }
    }
    *n_nodes_out = n_nodes;
    return B200PT_OK;
}
