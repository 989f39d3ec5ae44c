// Host-side BVHAccel::new with SplitMethod::HLBVH (reference: accelerators/src/bvh/hlbvh.rs:33-449, morton.rs:37-120,
// mod.rs:43-153).
//
// Design (not a port): the reference builds arena trees per treelet, an arena SAH tree above them and flattens the
// whole thing afterwards.  Here each treelet is emitted once, already in depth-first order, into its own slice of a
// scratch array (a treelet over k primitives has at most 2k-1 nodes, so slice [2*first, 2*(first+k)) never overlaps);
// the upper SAH recursion then writes the final array front to back and block-copies each treelet where its root
// lands, rebasing the second-child indices.  Sorting uses one 64-bit key (code << 32 | input position): identical to
// the reference's stable 5 x 6-bit LSD radix sort because the input positions are ascending.
//
// Reference behaviour that is kept because it shows in the result (see oracle/oracle_hlbvh.h for the long form):
//  * encode_morton_3 interleaves bits of the IEEE-754 *bit pattern* of the scaled offset (morton.rs:43-49 with
//    float_to_bits = transmute, core/src/pbrt/common.rs:179-187), release-build semantics;
//  * treelets are emitted in order (the reference's `--nthreads 1` order of `ordered_prims`).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/b200pt.h"
#include "host_hlbvh.h"

extern "C" int b200pt_set_error(const char* msg);

namespace {

inline float fmin_ref(float a, float b) { return a < b ? a : b; }  // core/src/pbrt/common.rs:83-108
inline float fmax_ref(float a, float b) { return a > b ? a : b; }
struct Box {
    float lo[3], hi[3];
};
inline Box empty_box() {
    const float m = std::numeric_limits<float>::max();
    return Box{{m, m, m}, {-m, -m, -m}};
}
inline void grow(Box& b, const float* o) {  // o = lo.xyz, hi.xyz
    for (int k = 0; k < 3; ++k) { b.lo[k] = fmin_ref(b.lo[k], o[k]); b.hi[k] = fmax_ref(b.hi[k], o[3 + k]); }
}
inline float area(const Box& b) {  // bounds3.rs:94-105
    if (b.hi[0] < b.lo[0] || b.hi[1] < b.lo[1] || b.hi[2] < b.lo[2]) return 0.0f;
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    float h = dx * dy + dx * dz + dy * dz;
    return h + h;
}
inline uint32_t spread3(uint32_t x) {  // left_shift_3, morton.rs:102-120, debug_assert compiled out
    if (x == (1u << 10)) x -= 1;
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
inline uint32_t bits_of(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline void put_box(b200pt_bvh_node& nd, const Box& b) {
    for (int k = 0; k < 3; ++k) { nd.bounds[k] = b.lo[k]; nd.bounds[3 + k] = b.hi[k]; }
}
inline void unite_children(b200pt_bvh_node& nd, const b200pt_bvh_node& c0, const b200pt_bvh_node& c1) {  // common.rs:150-159
    for (int k = 0; k < 3; ++k) {
        nd.bounds[k] = fmin_ref(c0.bounds[k], c1.bounds[k]);
        nd.bounds[3 + k] = fmax_ref(c0.bounds[3 + k], c1.bounds[3 + k]);
    }
}

constexpr int kBins = 12;             // hlbvh.rs:19
constexpr uint32_t kTreeletMask = 0x3FFC0000u;  // hlbvh.rs:55: top 12 of 30 bits
constexpr int kFirstBit = 30 - 1 - 12;          // hlbvh.rs:20

struct Treelet {
    size_t first, count;  // range of the sorted array
    uint32_t n_nodes;     // nodes in scratch[2 * first ...]
};

struct Builder {
    const float* prim_bounds;
    int max_prims;
    std::vector<uint64_t> keys;          // code << 32 | primitive
    std::vector<b200pt_bvh_node> scratch;  // treelet slices
    std::vector<Treelet> treelets;
    b200pt_bvh_node* out;
    uint32_t* ordered;
    int64_t n_out = 0;
    const char* error = nullptr;

    uint32_t code(size_t i) const { return (uint32_t)(keys[i] >> 32); }
    uint32_t prim(size_t i) const { return (uint32_t)keys[i]; }

    // emit_lbvh (hlbvh.rs:243-345) for one treelet, iteratively, straight into pre-order
    void emit_treelet(Treelet& t) {
        struct Job { size_t first, count; int bit; int64_t parent; };
        b200pt_bvh_node* dst = scratch.data() + 2 * t.first;
        std::vector<Job> work;
        work.push_back(Job{t.first, t.count, kFirstBit, -1});
        uint32_t n_nodes = 0;
        while (!work.empty()) {
            Job j = work.back();
            work.pop_back();
            // skip bits that do not split the range (hlbvh.rs:278-291)
            while (j.bit >= 0 && j.count >= (size_t)max_prims &&
                   ((code(j.first) ^ code(j.first + j.count - 1)) & (1u << j.bit)) == 0)
                --j.bit;
            const uint32_t me = n_nodes++;
            if (j.parent >= 0) dst[j.parent].offset = me;
            b200pt_bvh_node& nd = dst[me];
            nd.pad = 0;
            if (j.bit == -1 || j.count < (size_t)max_prims) {  // hlbvh.rs:257-273
                Box b = empty_box();
                for (size_t i = 0; i < j.count; ++i) {
                    ordered[j.first + i] = prim(j.first + i);
                    grow(b, prim_bounds + 6 * (size_t)prim(j.first + i));
                }
                put_box(nd, b);
                if (j.count >= 65536) error = "b200pt_bvh_build_hlbvh: leaf with >= 65536 primitives (reference asserts)";
                nd.offset = (uint32_t)j.first;
                nd.n_primitives = (uint16_t)j.count;
                nd.axis = 0;
                continue;
            }
            // first position whose bit differs from the range's first (hlbvh.rs:293-316: the codes are sorted, so the
            // bit is 0...01...1 over the range and the binary search finds the boundary)
            const uint32_t mask = 1u << j.bit;
            size_t lo = 0, hi = j.count - 1;
            while (lo + 1 != hi) {
                size_t mid = (lo + hi) / 2;
                if (((code(j.first + lo) ^ code(j.first + mid)) & mask) == 0) lo = mid; else hi = mid;
            }
            nd.n_primitives = 0;
            nd.axis = (uint8_t)(j.bit % 3);
            nd.offset = 0;
            work.push_back(Job{j.first + hi, j.count - hi, j.bit - 1, (int64_t)me});
            work.push_back(Job{j.first, hi, j.bit - 1, -1});
        }
        for (int64_t i = (int64_t)n_nodes - 1; i >= 0; --i)
            if (dst[i].n_primitives == 0) unite_children(dst[i], dst[i + 1], dst[dst[i].offset]);
        t.n_nodes = n_nodes;
    }

    static int bin_of(float centroid, float lo, float hi) {  // hlbvh.rs:382-389
        float v = (float)kBins * ((centroid - lo) / (hi - lo));
        int b = (!(v == v) || v <= 0.0f) ? 0 : (v >= 2147483648.0f ? 0x7fffffff : (int)v);
        return b == kBins ? kBins - 1 : b;
    }

    // build_upper_sah (hlbvh.rs:353-449) fused with flatten_bvh_tree (mod.rs:126-153): lays the final array out.
    // Needs only the treelets' root boxes and node counts; records where each treelet's block starts (treelet_base)
    // and the interior nodes above the treelets (upper / upper_index).  Returns the node's index.
    std::vector<b200pt_bvh_node> root_box;   // per treelet: its root node (bounds)
    std::vector<int64_t> treelet_base;       // per treelet: index of its root in the final array
    std::vector<b200pt_bvh_node> upper;      // interior nodes of the upper SAH tree
    std::vector<int64_t> upper_index;        // their indices in the final array
    const b200pt_bvh_node& root_of(uint32_t t) const { return root_box[t]; }

    struct Placed { int64_t index; b200pt_bvh_node node; };
    Placed layout(std::vector<uint32_t>& order, size_t start, size_t end) {
        const int64_t me = n_out;
        if (end - start == 1) {
            const uint32_t t = order[start];
            treelet_base[t] = me;
            n_out += treelets[t].n_nodes;
            return Placed{me, root_box[t]};
        }
        n_out += 1;
        b200pt_bvh_node nd{};
        Box bounds = empty_box(), cb = empty_box();
        for (size_t i = start; i < end; ++i) grow(bounds, root_of(order[i]).bounds);
        for (size_t i = start; i < end; ++i) {
            const float* rb = root_of(order[i]).bounds;
            float c[6];
            for (int k = 0; k < 3; ++k) c[k] = c[3 + k] = (rb[k] + rb[3 + k]) * 0.5f;
            grow(cb, c);
        }
        float dx = cb.hi[0] - cb.lo[0], dy = cb.hi[1] - cb.lo[1], dz = cb.hi[2] - cb.lo[2];
        const int dim = (dx > dy && dx > dz) ? 0 : (dy > dz ? 1 : 2);
        if (cb.hi[dim] == cb.lo[dim]) { error = "b200pt_bvh_build_hlbvh: treelet centroids coincide (reference asserts, hlbvh.rs:376)"; return Placed{me, nd}; }
        size_t count[kBins] = {0};
        Box box[kBins];
        for (int b = 0; b < kBins; ++b) box[b] = empty_box();
        auto bin = [&](uint32_t t) {
            const float* rb = root_of(t).bounds;
            return bin_of((rb[dim] + rb[3 + dim]) * 0.5f, cb.lo[dim], cb.hi[dim]);
        };
        for (size_t i = start; i < end; ++i) {
            int b = bin(order[i]);
            if (b < 0 || b >= kBins) { error = "b200pt_bvh_build_hlbvh: bucket out of range (reference asserts)"; return Placed{me, nd}; }
            count[b] += 1;
            grow(box[b], root_of(order[i]).bounds);
        }
        Box right[kBins - 1];
        size_t nr[kBins - 1];
        Box acc = empty_box();
        size_t cnt = 0;
        for (int b = kBins - 1; b >= 1; --b) { grow(acc, box[b].lo); cnt += count[b]; right[b - 1] = acc; nr[b - 1] = cnt; }
        acc = empty_box();
        cnt = 0;
        const float total_area = area(bounds);
        float best = 0.0f;
        int best_bin = 0;
        for (int b = 0; b < kBins - 1; ++b) {
            grow(acc, box[b].lo);
            cnt += count[b];
            float cost = 0.125f + ((float)cnt * area(acc) + (float)nr[b] * area(right[b])) / total_area;
            if (b == 0 || cost < best) { best = cost; best_bin = b; }
        }
        // itertools::partition over the treelet roots (hlbvh.rs:427-437)
        size_t f = start, bk = end, split = 0;
        while (f < bk) {
            size_t front = f++;
            if (!(bin(order[front]) <= best_bin)) {
                bool found = false;
                while (bk > f) {
                    --bk;
                    if (bin(order[bk]) <= best_bin) { found = true; break; }
                }
                if (!found) break;
                std::swap(order[front], order[bk]);
            }
            ++split;
        }
        const size_t mid = start + split;
        if (!(mid > start && mid < end)) { error = "b200pt_bvh_build_hlbvh: upper SAH partition produced an empty side (reference asserts)"; return Placed{me, nd}; }
        const size_t slot = upper.size();
        upper.push_back(nd);
        upper_index.push_back(me);
        Placed c0 = layout(order, start, mid);
        if (error) return Placed{me, nd};
        Placed c1 = layout(order, mid, end);
        if (error) return Placed{me, nd};
        unite_children(nd, c0.node, c1.node);
        nd.offset = (uint32_t)c1.index;
        nd.n_primitives = 0;
        nd.axis = (uint8_t)dim;
        nd.pad = 0;
        upper[slot] = nd;
        return Placed{me, nd};
    }
    void layout_all() {
        std::vector<uint32_t> order(treelets.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
        treelet_base.assign(treelets.size(), 0);
        n_out = 0;
        if (!error && !order.empty()) layout(order, 0, order.size());
    }
};

}  // namespace

// Morton codes as the reference computes them (exposed for the tests): code[i] for primitive i.
extern "C" int b200pt_hlbvh_morton_codes(const float* prim_bounds, int64_t n, uint32_t* codes_out) {
    if (n < 0 || (n > 0 && (!prim_bounds || !codes_out))) { b200pt_set_error("b200pt_hlbvh_morton_codes: invalid argument"); return B200PT_ERR_INVALID; }
    Box all = empty_box();
    for (int64_t i = 0; i < n; ++i) grow(all, prim_bounds + 6 * i);  // hlbvh.rs:42
    for (int64_t i = 0; i < n; ++i) {
        const float* pb = prim_bounds + 6 * i;
        uint32_t c = 0;
        for (int k = 2; k >= 0; --k) {
            float cen = 0.5f * (pb[k] + pb[3 + k]);  // common.rs:86
            float o = cen - all.lo[k];               // Bounds3::offset, bounds3.rs:153-168
            if (all.hi[k] > all.lo[k]) o /= all.hi[k] - all.lo[k];
            c |= spread3(bits_of(o * 1024.0f)) << k;  // morton.rs:43-49
        }
        codes_out[i] = c;
    }
    return B200PT_OK;
}

extern "C" int b200pt_bvh_build_hlbvh(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                                      int64_t* n_nodes_out, uint32_t* ordered_out) {
    if (n < 0 || !n_nodes_out || (n > 0 && (!prim_bounds || !nodes_out || !ordered_out))) {
        b200pt_set_error("b200pt_bvh_build_hlbvh: invalid argument");
        return B200PT_ERR_INVALID;
    }
    *n_nodes_out = 0;
    if (n == 0) return B200PT_OK;
    if (n > 0x7fffffffLL) { b200pt_set_error("b200pt_bvh_build_hlbvh: too many primitives"); return B200PT_ERR_INVALID; }
    Builder B;
    B.prim_bounds = prim_bounds;
    B.max_prims = max_prims_in_node & 0xff;
    B.out = nodes_out;
    B.ordered = ordered_out;
    std::vector<uint32_t> codes((size_t)n);
    b200pt_hlbvh_morton_codes(prim_bounds, n, codes.data());
    B.keys.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) B.keys[(size_t)i] = ((uint64_t)codes[(size_t)i] << 32) | (uint64_t)i;
    std::sort(B.keys.begin(), B.keys.end());
    for (size_t start = 0, end = 1; end <= (size_t)n; ++end) {  // hlbvh.rs:53-69
        if (end == (size_t)n || ((B.code(start) ^ B.code(end)) & kTreeletMask) != 0) {
            B.treelets.push_back(Treelet{start, end - start, 0});
            start = end;
        }
    }
    B.scratch.resize(2 * (size_t)n);
    for (Treelet& t : B.treelets) B.emit_treelet(t);
    for (const Treelet& t : B.treelets) B.root_box.push_back(B.scratch[2 * t.first]);
    B.layout_all();
    if (B.error) { b200pt_set_error(B.error); return B200PT_ERR_INVALID; }
    for (size_t k = 0; k < B.upper.size(); ++k) nodes_out[B.upper_index[k]] = B.upper[k];
    for (size_t t = 0; t < B.treelets.size(); ++t) {  // block-copy every treelet to where its root landed
        const Treelet& tr = B.treelets[t];
        const b200pt_bvh_node* src = B.scratch.data() + 2 * tr.first;
        const int64_t base = B.treelet_base[t];
        for (uint32_t i = 0; i < tr.n_nodes; ++i) {
            nodes_out[base + i] = src[i];
            if (src[i].n_primitives == 0) nodes_out[base + i].offset += (uint32_t)base;
        }
    }
    *n_nodes_out = B.n_out;
    return B200PT_OK;
}

// Upper-tree layout for the GPU builder (csrc/bvh_build.cu): treelet root nodes + node counts in, final positions out.
// upper_nodes_out / upper_index_out need room for n_treelets - 1 entries.
extern "C" int b200pt_hlbvh_upper_layout(const b200pt_bvh_node* treelet_roots, const uint32_t* treelet_n_nodes, int64_t n_treelets,
                                         b200pt_bvh_node* upper_nodes_out, int64_t* upper_index_out, int64_t* n_upper_out, int64_t* treelet_base_out,
                                         int64_t* n_nodes_out) {
    if (n_treelets <= 0 || !treelet_roots || !treelet_n_nodes || !upper_nodes_out || !upper_index_out || !n_upper_out || !treelet_base_out || !n_nodes_out) {
        b200pt_set_error("b200pt_hlbvh_upper_layout: invalid argument");
        return B200PT_ERR_INVALID;
    }
    Builder B;
    B.prim_bounds = nullptr; B.max_prims = 0; B.out = nullptr; B.ordered = nullptr;
    for (int64_t t = 0; t < n_treelets; ++t) {
        B.treelets.push_back(Treelet{0, 0, treelet_n_nodes[t]});
        B.root_box.push_back(treelet_roots[t]);
    }
    B.layout_all();
    if (B.error) { b200pt_set_error(B.error); return B200PT_ERR_INVALID; }
    for (size_t k = 0; k < B.upper.size(); ++k) { upper_nodes_out[k] = B.upper[k]; upper_index_out[k] = B.upper_index[k]; }
    for (int64_t t = 0; t < n_treelets; ++t) treelet_base_out[t] = B.treelet_base[(size_t)t];
    *n_upper_out = (int64_t)B.upper.size();
    *n_nodes_out = B.n_out;
    return B200PT_OK;
}
