// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY: pinned by execution for Whitted-class renders, see oracle_math.h.
//
// BVHAccel (SAH build, flatten, closest-hit / any-hit traversal) and the
// watertight Triangle test, restated from accelerators/src/bvh/{mod,common,sah}.rs,
// core/src/geometry/bounds3.rs and shapes/src/triangle.rs.
#pragma once
#include <vector>
#include "oracle_math.h"
#include "oracle_texture.h"

namespace orc {

// accelerators/src/bvh/common.rs:163-179.  The reference struct has no
// #[repr(C)]; this 32-byte layout is the interchange form used by the C ABI.
struct LinearBVHNode {
    Float bounds[6];  // pmin.xyz, pmax.xyz
    uint32_t offset;  // leaf: first primitive; interior: second child
    uint16_t n_primitives;
    uint8_t axis;
    uint8_t pad;
};
static_assert(sizeof(LinearBVHNode) == 32, "LinearBVHNode must be 32 bytes");

// accelerators/src/bvh/common.rs:66-88
struct BVHPrimitiveInfo {
    size_t primitive_number;
    Bounds3 bounds;
    V3 centroid;
};

struct BVHBuildNode {
    Bounds3 bounds;
    int children[2];  // indices into the build-node pool, -1 = none
    int split_axis;
    size_t first_prim_offset, n_primitives;
};

struct BVHBuilder {
    std::vector<BVHPrimitiveInfo> info;
    std::vector<BVHBuildNode> pool;
    std::vector<uint32_t> ordered;  // ordered_prims: ordered position -> original primitive number
    int max_prims_in_node;
    size_t total_nodes = 0;

    static const int kBuckets = 12;  // sah.rs:11

    // sah.rs:293-313 / 354-360: bucket of a centroid (saturating `as usize`).
    static int bucket_of(const Bounds3& cb, V3 c, int dim) {
        Float v = (Float)kBuckets * boffset(cb, c)[dim];
        int b;
        if (!(v == v) || v <= 0.0f) b = 0;                // NaN / negative saturate to 0
        else if (v >= 2147483648.0f) b = 0x7fffffff;
        else b = (int)v;
        if (b == kBuckets) b = kBuckets - 1;
        return b;
    }

    int new_leaf(size_t start, size_t end, const Bounds3& bounds) {
        // sah.rs:187-211
        BVHBuildNode n;
        n.bounds = bounds;
        n.children[0] = n.children[1] = -1;
        n.split_axis = 0;
        n.first_prim_offset = ordered.size();
        n.n_primitives = end - start;
        for (size_t i = start; i < end; ++i) ordered.push_back((uint32_t)info[i].primitive_number);
        pool.push_back(n);
        return (int)pool.size() - 1;
    }

    // itertools 0.13 `partition` (call sites sah.rs:229,354): scan from the
    // front for an element failing the predicate, rfind from the back for one
    // satisfying it, swap; stop when the back search fails.
    template <class Pred> size_t partition(size_t start, size_t end, Pred pred) {
        size_t split = 0;
        size_t f = start, b = end;  // remaining range [f, b)
        while (f < b) {
            size_t front = f++;
            if (!pred(info[front])) {
                bool found = false;
                while (b > f) {
                    --b;
                    if (pred(info[b])) { found = true; break; }
                }
                if (!found) break;
                BVHPrimitiveInfo t = info[front]; info[front] = info[b]; info[b] = t;
            }
            split += 1;
        }
        return split;
    }

    // sah.rs:26-125
    int build(size_t start, size_t end) {
        total_nodes += 1;
        Bounds3 bounds;
        for (size_t i = start; i < end; ++i) bounds = bunion(bounds, info[i].bounds);
        size_t n = end - start;
        if (n == 1) return new_leaf(start, end, bounds);

        Bounds3 cb;
        for (size_t i = start; i < end; ++i) cb = bunion(cb, info[i].centroid);
        int dim = maximum_extent(cb);
        if (cb.pmax[dim] == cb.pmin[dim]) return new_leaf(start, end, bounds);

        size_t mid;
        if (n <= 2) {
            // sah.rs:81-83, 240-254: equal counts; for n == 2 the smaller
            // centroid goes first (centroids differ along dim, checked above).
            mid = (start + end) / 2;
            if (info[end - 1].centroid[dim] < info[start].centroid[dim]) {
                BVHPrimitiveInfo t = info[start]; info[start] = info[end - 1]; info[end - 1] = t;
            }
        } else {
            // sah.rs:293-367
            struct Bucket { size_t count = 0; Bounds3 bounds; } buckets[kBuckets];
            for (size_t i = start; i < end; ++i) {
                int b = bucket_of(cb, info[i].centroid, dim);
                buckets[b].count += 1;
                buckets[b].bounds = bunion(buckets[b].bounds, info[i].bounds);
            }
            Float cost[kBuckets - 1];
            for (int i = 0; i < kBuckets - 1; ++i) {
                Bounds3 b0, b1;
                size_t c0 = 0, c1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = bunion(b0, buckets[j].bounds); c0 += buckets[j].count; }
                for (int j = i + 1; j < kBuckets; ++j) { b1 = bunion(b1, buckets[j].bounds); c1 += buckets[j].count; }
                cost[i] = 1.0f + ((Float)c0 * surface_area(b0) + (Float)c1 * surface_area(b1)) / surface_area(bounds);
            }
            Float min_cost = cost[0];
            int min_bucket = 0;
            for (int i = 1; i < kBuckets - 1; ++i)
                if (cost[i] < min_cost) { min_cost = cost[i]; min_bucket = i; }
            Float leaf_cost = (Float)n;
            if (n > (size_t)max_prims_in_node || min_cost < leaf_cost) {
                size_t split = partition(start, end, [&](const BVHPrimitiveInfo& pi) {
                    return bucket_of(cb, pi.centroid, dim) <= min_bucket;
                });
                mid = start + split;
            } else {
                return new_leaf(start, end, bounds);
            }
        }
        // sah.rs:144-176: children first, then the interior node.
        int c0 = build(start, mid);
        int c1 = build(mid, end);
        BVHBuildNode nd;
        nd.bounds = bunion(pool[c0].bounds, pool[c1].bounds);  // common.rs:150-159
        nd.children[0] = c0;
        nd.children[1] = c1;
        nd.split_axis = dim;
        nd.first_prim_offset = 0;
        nd.n_primitives = 0;
        pool.push_back(nd);
        return (int)pool.size() - 1;
    }

    // mod.rs:126-153
    uint32_t flatten(int node, std::vector<LinearBVHNode>& nodes, uint32_t* offset) {
        const BVHBuildNode nd = pool[node];
        uint32_t my = (*offset)++;
        LinearBVHNode& ln = nodes[my];
        auto set_bounds = [&](LinearBVHNode& l) {
            l.bounds[0] = nd.bounds.pmin.x; l.bounds[1] = nd.bounds.pmin.y; l.bounds[2] = nd.bounds.pmin.z;
            l.bounds[3] = nd.bounds.pmax.x; l.bounds[4] = nd.bounds.pmax.y; l.bounds[5] = nd.bounds.pmax.z;
        };
        if (nd.n_primitives > 0) {
            set_bounds(ln);
            ln.offset = (uint32_t)nd.first_prim_offset;
            ln.n_primitives = (uint16_t)nd.n_primitives;
            ln.axis = 0; ln.pad = 0;
        } else {
            flatten(nd.children[0], nodes, offset);
            uint32_t second = flatten(nd.children[1], nodes, offset);
            LinearBVHNode& l2 = nodes[my];
            set_bounds(l2);
            l2.offset = second;
            l2.n_primitives = 0;
            l2.axis = (uint8_t)nd.split_axis; l2.pad = 0;
        }
        return my;
    }
};

// BVHAccel::new (mod.rs:43-124), SAH only.  prim_bounds: 6 floats per primitive
// (world_bound of each primitive in RenderOptions.primitives order).
inline void bvh_build_sah(const Float* prim_bounds, size_t n, int max_prims_in_node, std::vector<LinearBVHNode>& nodes,
                          std::vector<uint32_t>& ordered) {
    nodes.clear();
    ordered.clear();
    if (n == 0) return;
    BVHBuilder b;
    b.max_prims_in_node = max_prims_in_node;
    b.info.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const Float* pb = prim_bounds + 6 * i;
        b.info[i].primitive_number = i;
        b.info[i].bounds = Bounds3(V3(pb[0], pb[1], pb[2]), V3(pb[3], pb[4], pb[5]));
        b.info[i].centroid = 0.5f * (b.info[i].bounds.pmin + b.info[i].bounds.pmax);  // common.rs:86
    }
    b.ordered.reserve(n);
    b.pool.reserve(2 * n);
    int root = b.build(0, n);
    nodes.assign(b.total_nodes, LinearBVHNode());
    uint32_t off = 0;
    b.flatten(root, nodes, &off);
    ordered.swap(b.ordered);
}

// core/src/geometry/bounds3.rs:292-325 — note t_z_max is NOT inflated.
inline bool bounds_intersect_p_inv(const Float* b, const Ray& ray, V3 inv_dir, const int neg[3]) {
    // b = pmin.xyz, pmax.xyz ; bounds[i] selects pmin (0) or pmax (1)
    Float t_min = (b[3 * neg[0] + 0] - ray.o.x) * inv_dir.x;
    Float t_max = (b[3 * (1 - neg[0]) + 0] - ray.o.x) * inv_dir.x;
    Float t_y_min = (b[3 * neg[1] + 1] - ray.o.y) * inv_dir.y;
    Float t_y_max = (b[3 * (1 - neg[1]) + 1] - ray.o.y) * inv_dir.y;
    Float g3 = gamma(3);
    t_max *= 1.0f + 2.0f * g3;
    t_y_max *= 1.0f + 2.0f * g3;
    if (t_min > t_y_max || t_y_min > t_max) return false;
    if (t_y_min > t_min) t_min = t_y_min;
    if (t_y_max < t_max) t_max = t_y_max;
    Float t_z_min = (b[3 * neg[2] + 2] - ray.o.z) * inv_dir.z;
    Float t_z_max = (b[3 * (1 - neg[2]) + 2] - ray.o.z) * inv_dir.z;
    if (t_min > t_z_max || t_z_min > t_max) return false;
    if (t_z_min > t_min) t_min = t_z_min;
    if (t_z_max < t_max) t_max = t_z_max;
    return t_min < ray.t_max && t_max > 0.0f;
}

// Result of the ray/triangle test proper (triangle.rs:438-545).
struct TriHit {
    Float t, b0, b1, b2;
    // diagnostics for the parity exemption band
    Float det, min_e_abs;
};

// shapes/src/triangle.rs:438-545 (== :731-838 for intersect_p).  Returns true
// when the candidate passes every test up to `t <= delta_t`.
inline bool triangle_test(const Ray& r, V3 p0, V3 p1, V3 p2, TriHit* h) {
    V3 p0t = p0 - r.o, p1t = p1 - r.o, p2t = p2 - r.o;
    int kz = max_dimension(vabs(r.d));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    V3 d = permute(r.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz);
    p1t = permute(p1t, kx, ky, kz);
    p2t = permute(p2t, kx, ky, kz);
    Float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    Float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    Float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    Float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (Float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (Float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (Float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < r.t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > r.t_max * det)) return false;
    Float inv_det = 1.0f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_z_t = max_component(vabs(V3(p0t.z, p1t.z, p2t.z)));
    Float delta_z = gamma(3) * max_z_t;
    Float max_x_t = max_component(vabs(V3(p0t.x, p1t.x, p2t.x)));
    Float max_y_t = max_component(vabs(V3(p0t.y, p1t.y, p2t.y)));
    Float delta_x = gamma(5) * (max_x_t + max_z_t);
    Float delta_y = gamma(5) * (max_y_t + max_z_t);
    Float delta_e = 2.0f * (gamma(2) * max_x_t * max_y_t + delta_y * max_x_t + delta_x * max_y_t);
    Float max_e = max_component(vabs(V3(e0, e1, e2)));
    Float delta_t = 3.0f * (gamma(3) * max_e * max_z_t + delta_e * max_z_t + delta_z * max_e) * pabs(inv_det);
    if (t <= delta_t) return false;
    h->t = t; h->b0 = b0; h->b1 = b1; h->b2 = b2;
    h->det = det;
    h->min_e_abs = pmin(pmin(pabs(e0), pabs(e1)), pabs(e2));
    return true;
}

// Hit geometry of an accepted candidate (triangle.rs:547-725).  Optional per-triangle vertex attributes: uv (get_uvs,
// triangle.rs:384-394: defaults (0,0),(1,0),(1,1)), normals N and tangents S (already in world space:
// TriangleMesh::new transforms them and does NOT renormalise, triangle.rs:92-99).
struct TriAttr {
    const Float* uv = nullptr;   // 6 floats: uv0 uv1 uv2
    const Float* n = nullptr;    // 9 floats: n0 n1 n2
    const Float* s = nullptr;    // 9 floats: s0 s1 s2
    bool flip = false;           // reverse_orientation ^ transform_swaps_handedness
    bool reverse = false;        // reverse_orientation
};
struct TriGeom {
    V3 p, p_error, n, dpdu, dpdv;
    P2 uv;
    V3 shading_n, shading_dpdu;  // Shading::{n, dpdu} after set_shading_geometry (== n, dpdu without N / S)
    V3 dndu, dndv;               // Shading::{dndu, dndv}: zero without vertex normals (they only feed the differentials of specular children)
};
// Returns false when the reference rejects the hit as degenerate (triangle.rs:567-572).
inline bool triangle_geometry(V3 p0, V3 p1, V3 p2, Float b0, Float b1, Float b2, const TriAttr& at, TriGeom* g) {
    P2 uv0(0.0f, 0.0f), uv1(1.0f, 0.0f), uv2(1.0f, 1.0f);
    if (at.uv) { uv0 = P2(at.uv[0], at.uv[1]); uv1 = P2(at.uv[2], at.uv[3]); uv2 = P2(at.uv[4], at.uv[5]); }
    Float duv02x = uv0.x - uv2.x, duv02y = uv0.y - uv2.y;
    Float duv12x = uv1.x - uv2.x, duv12y = uv1.y - uv2.y;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    Float determinant = duv02x * duv12y - duv02y * duv12x;
    bool degenerate_uv = pabs(determinant) < 1e-8f;
    V3 dpdu, dpdv;
    if (!degenerate_uv) {
        Float invdet = 1.0f / determinant;
        dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
    }
    if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
        V3 ng = cross(p2 - p0, p1 - p0);
        if (length_squared(ng) == 0.0f) return false;
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    Float xs = pabs(b0 * p0.x) + pabs(b1 * p1.x) + pabs(b2 * p2.x);
    Float ys = pabs(b0 * p0.y) + pabs(b1 * p1.y) + pabs(b2 * p2.y);
    Float zs = pabs(b0 * p0.z) + pabs(b1 * p1.z) + pabs(b2 * p2.z);
    g->p_error = gamma(7) * V3(xs, ys, zs);
    g->p = b0 * p0 + b1 * p1 + b2 * p2;
    g->uv = P2(b0 * uv0.x + b1 * uv1.x + b2 * uv2.x, b0 * uv0.y + b1 * uv1.y + b2 * uv2.y);
    g->dpdu = dpdu;
    g->dpdv = dpdv;
    V3 n = normalize(cross(dp02, dp12));  // triangle.rs:625
    if (at.flip) n = -n;
    g->n = n;
    g->shading_n = n;
    g->shading_dpdu = dpdu;
    if (at.n || at.s) {  // triangle.rs:631-721
        V3 ns = n;
        if (at.n) {
            V3 ns2 = b0 * V3(at.n[0], at.n[1], at.n[2]) + b1 * V3(at.n[3], at.n[4], at.n[5]) + b2 * V3(at.n[6], at.n[7], at.n[8]);
            if (length_squared(ns2) > 0.0f) ns = normalize(ns2);
        }
        V3 ss = normalize(dpdu);
        if (at.s) {
            V3 ss2 = b0 * V3(at.s[0], at.s[1], at.s[2]) + b1 * V3(at.s[3], at.s[4], at.s[5]) + b2 * V3(at.s[6], at.s[7], at.s[8]);
            if (length_squared(ss2) > 0.0f) ss = normalize(ss2);
        }
        V3 ts = cross(ss, ns);
        if (length_squared(ts) > 0.0f) {
            ts = normalize(ts);
            ss = cross(ts, ns);
        } else {
            coordinate_system(ns, &ss, &ts);
        }
        if (at.n) {  // triangle.rs:679-713
            V3 n0(at.n[0], at.n[1], at.n[2]), n1(at.n[3], at.n[4], at.n[5]), n2(at.n[6], at.n[7], at.n[8]);
            V3 dn1 = n0 - n2, dn2 = n1 - n2;
            if (degenerate_uv) {
                V3 dn = cross(n2 - n0, n1 - n0);
                if (length_squared(dn) != 0.0f) coordinate_system(dn, &g->dndu, &g->dndv);
            } else {
                Float invdet = 1.0f / determinant;
                g->dndu = (duv12y * dn1 - duv02y * dn2) * invdet;
                g->dndv = (-duv12x * dn1 + duv02x * dn2) * invdet;
            }
        }
        if (at.reverse) ts = -ts;
        // SurfaceInteraction::set_shading_geometry(ss, ts, .., orientation_is_authoritative = true), surface_interaction.rs:152-173
        g->shading_n = normalize(cross(ss, ts));
        g->n = face_forward(g->n, g->shading_n);
        g->shading_dpdu = ss;
    }
    return true;
}

// Primitive flags carried per triangle (constant alpha textures,
// triangle.rs:278-312: every mesh owns an alpha and a shadow-alpha texture).
enum : uint32_t {
    PRIM_FLIP_NORMAL = 1u,        // reverse_orientation ^ transform_swaps_handedness
    PRIM_ALPHA_ZERO = 2u,         // constant alpha texture evaluates to exactly 0
    PRIM_SHADOW_ALPHA_ZERO = 4u,  // constant shadowalpha texture evaluates to exactly 0
    PRIM_REVERSE_ORIENTATION = 8u,  // reverse_orientation alone (flips the shading bitangent, triangle.rs:714-716)
    PRIM_HAS_UV = 16u,            // the primitive's mesh has "uv"/"st" (tri_uvs holds its three uvs)
    PRIM_HAS_NORMALS = 32u,       // ... has "N" (tri_normals)
    PRIM_HAS_TANGENTS = 64u,      // ... has "S" (tri_tangents)
    PRIM_ALPHA_TEXTURE = 128u,    // the mesh's alpha / shadowalpha is a non-constant texture (Accel::alpha_tex)
};

// Flattened accelerator: nodes + triangles in ordered_prims order.
struct Accel {
    std::vector<LinearBVHNode> nodes;
    std::vector<uint32_t> ordered;     // ordered position -> original primitive index
    std::vector<Float> verts;          // 9 floats per ORIGINAL primitive
    std::vector<uint32_t> flags;       // per ORIGINAL primitive
    std::vector<Float> uvs, normals, tangents;  // optional: 6 / 9 / 9 floats per ORIGINAL primitive (empty = the mesh has none)
    // alpha masks (triangle.rs:278-312): per ORIGINAL primitive the index of its mesh's alpha / shadowalpha texture or -1
    std::vector<int32_t> alpha_tex;
    std::vector<FloatTexture> textures;
    std::vector<uint8_t> noise_perm;
    // mask.evaluate(..) == 0.0 -> the hit is rejected (triangle.rs:601-603, 886-899)
    bool alpha_rejects(size_t prim, P2 uv, bool shadow) const {
        if (!(flag(prim) & PRIM_ALPHA_TEXTURE) || alpha_tex.empty()) return false;
        int ta = alpha_tex[2 * prim], ts = alpha_tex[2 * prim + 1];
        if (ta >= 0 && float_texture_evaluate(textures[(size_t)ta], noise_perm.data(), uv.x, uv.y) == 0.0f) return true;
        if (shadow && ts >= 0 && float_texture_evaluate(textures[(size_t)ts], noise_perm.data(), uv.x, uv.y) == 0.0f) return true;
        return false;
    }
    V3 vert(size_t prim, int k) const { const Float* v = &verts[9 * prim + 3 * k]; return V3(v[0], v[1], v[2]); }
    uint32_t flag(size_t prim) const { return flags.empty() ? 0u : flags[prim]; }
    TriAttr attr(size_t prim) const {
        TriAttr at;
        const uint32_t fl = flag(prim);
        if (!uvs.empty() && (fl & PRIM_HAS_UV)) at.uv = &uvs[6 * prim];
        if (!normals.empty() && (fl & PRIM_HAS_NORMALS)) at.n = &normals[9 * prim];
        if (!tangents.empty() && (fl & PRIM_HAS_TANGENTS)) at.s = &tangents[9 * prim];
        at.flip = (fl & PRIM_FLIP_NORMAL) != 0;
        at.reverse = (fl & PRIM_REVERSE_ORIENTATION) != 0;
        return at;
    }
};

struct HitRecord {
    Float t;
    uint32_t prim;  // original primitive index, 0xffffffff = miss
    Float b0, b1;
    // diagnostics (not part of the device format)
    Float b2, det, min_e_abs;
    Float second_t;  // closest other accepted candidate t seen (for the tie band)
};

struct TraversalCounters {
    uint64_t nodes = 0, tris = 0;
};

// Triangle::intersect as seen through GeometricPrimitive::intersect
// (geometric_primitive.rs:67-88): test, degenerate rejection, alpha test.
inline bool prim_intersect(const Accel& a, uint32_t prim, const Ray& r, TriHit* th) {
    V3 p0 = a.vert(prim, 0), p1 = a.vert(prim, 1), p2 = a.vert(prim, 2);
    if (!triangle_test(r, p0, p1, p2, th)) return false;
    TriGeom g;
    uint32_t fl = a.flag(prim);
    if (!triangle_geometry(p0, p1, p2, th->b0, th->b1, th->b2, a.attr(prim), &g)) return false;
    if (fl & PRIM_ALPHA_ZERO) return false;  // triangle.rs:587-607
    if (a.alpha_rejects(prim, g.uv, false)) return false;
    return true;
}
// Triangle::intersect_p (triangle.rs:731-903).
inline bool prim_intersect_p(const Accel& a, uint32_t prim, const Ray& r) {
    V3 p0 = a.vert(prim, 0), p1 = a.vert(prim, 1), p2 = a.vert(prim, 2);
    TriHit th;
    if (!triangle_test(r, p0, p1, p2, &th)) return false;
    TriGeom g;
    uint32_t fl = a.flag(prim);
    if (!triangle_geometry(p0, p1, p2, th.b0, th.b1, th.b2, a.attr(prim), &g)) return false;
    if (fl & (PRIM_ALPHA_ZERO | PRIM_SHADOW_ALPHA_ZERO)) return false;
    if (a.alpha_rejects(prim, g.uv, true)) return false;  // triangle.rs:886-899
    return true;
}

// BVHAccel::intersect, mod.rs:173-226.  Lowers r.t_max on every accepted hit.
inline bool bvh_intersect(const Accel& a, Ray& r, HitRecord* out, TraversalCounters* ctr = nullptr) {
    bool hit = false;
    out->prim = 0xffffffffu;
    out->t = kInfinity;
    out->second_t = kInfinity;
    if (a.nodes.empty()) return false;
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int neg[3] = {inv_dir.x < 0.0f ? 1 : 0, inv_dir.y < 0.0f ? 1 : 0, inv_dir.z < 0.0f ? 1 : 0};
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    for (;;) {
        const LinearBVHNode& node = a.nodes[cur];
        if (ctr) ctr->nodes += 1;
        if (bounds_intersect_p_inv(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    uint32_t prim = a.ordered[node.offset + i];
                    if (ctr) ctr->tris += 1;
                    TriHit th;
                    if (prim_intersect(a, prim, r, &th)) {
                        if (hit) out->second_t = out->t;
                        hit = true;
                        r.t_max = th.t;
                        out->t = th.t; out->prim = prim; out->b0 = th.b0; out->b1 = th.b1; out->b2 = th.b2;
                        out->det = th.det; out->min_e_abs = th.min_e_abs;
                    }
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            } else {
                if (neg[node.axis] == 1) { stack[to_visit++] = cur + 1; cur = node.offset; }
                else { stack[to_visit++] = node.offset; cur = cur + 1; }
            }
        } else {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return hit;
}

// BVHAccel::intersect_p, mod.rs:231-283.
inline bool bvh_intersect_p(const Accel& a, const Ray& r, TraversalCounters* ctr = nullptr) {
    if (a.nodes.empty()) return false;
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int neg[3] = {inv_dir.x < 0.0f ? 1 : 0, inv_dir.y < 0.0f ? 1 : 0, inv_dir.z < 0.0f ? 1 : 0};
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    for (;;) {
        const LinearBVHNode& node = a.nodes[cur];
        if (ctr) ctr->nodes += 1;
        if (bounds_intersect_p_inv(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    if (ctr) ctr->tris += 1;
                    if (prim_intersect_p(a, a.ordered[node.offset + i], r)) return true;
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            } else {
                if (neg[node.axis] == 1) { stack[to_visit++] = cur + 1; cur = node.offset; }
                else { stack[to_visit++] = node.offset; cur = cur + 1; }
            }
        } else {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return false;
}

// ---------------------------------------------------------------------------
// Two-level scenes: TransformedPrimitive (core/src/primitives/transformed_primitive.rs:43-73) with a static
// transform, as created by Api::pbrt_object_instance (api/src/lib.rs:937-987): the object's primitives get their own
// BVHAccel, the scene aggregate holds one TransformedPrimitive per ObjectInstance among its primitives.
struct Instance {
    int object;
    M4 i2w, w2i;  // primitive_to_world.m and its inverse
};
struct TopLevel {
    Accel top;                 // nodes / ordered over the TOP-LEVEL primitives; verts / flags of ALL triangles (global ids)
    int64_t n_top_tris = 0;    // top-level primitive p < n_top_tris is triangle p, otherwise instance p - n_top_tris
    std::vector<Accel> objects;            // per object: own nodes / ordered (local ids) and a copy of its vertices
    std::vector<int64_t> object_first_prim;  // global id of the object's first triangle
    std::vector<Instance> instances;
};

// Transform::transform_bounds, transform.rs:552-561 (TransformedPrimitive::world_bound for a static transform)
inline Bounds3 xf_bounds(const M4& m, const Bounds3& b) {
    Bounds3 r(xf_point(m, V3(b.pmin.x, b.pmin.y, b.pmin.z)), xf_point(m, V3(b.pmin.x, b.pmin.y, b.pmin.z)));
    r = bunion(r, xf_point(m, V3(b.pmax.x, b.pmin.y, b.pmin.z)));
    r = bunion(r, xf_point(m, V3(b.pmin.x, b.pmax.y, b.pmin.z)));
    r = bunion(r, xf_point(m, V3(b.pmin.x, b.pmin.y, b.pmax.z)));
    r = bunion(r, xf_point(m, V3(b.pmin.x, b.pmax.y, b.pmax.z)));
    r = bunion(r, xf_point(m, V3(b.pmax.x, b.pmax.y, b.pmin.z)));
    r = bunion(r, xf_point(m, V3(b.pmax.x, b.pmin.y, b.pmax.z)));
    r = bunion(r, xf_point(m, V3(b.pmax.x, b.pmax.y, b.pmax.z)));
    return r;
}

// TransformedPrimitive::intersect, transformed_primitive.rs:51-64: the ray goes to instance space through
// Transform::transform_ray (origin nudged by its error bound, t_max shortened), the nested aggregate is intersected and
// the INSTANCE-space t_max is written back to the world ray.  The hit's geometry is transformed by the caller.
inline bool instance_intersect(const TopLevel& s, int inst, Ray& r, HitRecord* out, TraversalCounters* ctr) {
    const Instance& I = s.instances[(size_t)inst];
    Ray ray = xf_ray(I.w2i, r);
    HitRecord h;
    if (!bvh_intersect(s.objects[(size_t)I.object], ray, &h, ctr)) return false;
    r.t_max = ray.t_max;
    *out = h;
    out->prim = (uint32_t)(s.object_first_prim[(size_t)I.object] + h.prim);
    return true;
}
inline bool instance_intersect_p(const TopLevel& s, int inst, const Ray& r, TraversalCounters* ctr) {
    const Instance& I = s.instances[(size_t)inst];
    Ray ray = xf_ray(I.w2i, r);
    return bvh_intersect_p(s.objects[(size_t)I.object], ray, ctr);
}

// BVHAccel::intersect over the scene aggregate whose primitives are triangles and TransformedPrimitives.
// *inst_out = instance index of the accepted hit, -1 for a top-level triangle.
inline bool top_intersect(const TopLevel& s, Ray& r, HitRecord* out, int* inst_out, TraversalCounters* ctr = nullptr) {
    const Accel& a = s.top;
    bool hit = false;
    out->prim = 0xffffffffu; out->t = kInfinity; out->second_t = kInfinity;
    *inst_out = -1;
    if (a.nodes.empty()) return false;
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int neg[3] = {inv_dir.x < 0.0f ? 1 : 0, inv_dir.y < 0.0f ? 1 : 0, inv_dir.z < 0.0f ? 1 : 0};
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    for (;;) {
        const LinearBVHNode& node = a.nodes[cur];
        if (ctr) ctr->nodes += 1;
        if (bounds_intersect_p_inv(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    uint32_t prim = a.ordered[node.offset + i];
                    if ((int64_t)prim < s.n_top_tris) {
                        if (ctr) ctr->tris += 1;
                        TriHit th;
                        if (prim_intersect(a, prim, r, &th)) {
                            hit = true; r.t_max = th.t; *inst_out = -1;
                            out->t = th.t; out->prim = prim; out->b0 = th.b0; out->b1 = th.b1; out->b2 = th.b2;
                            out->det = th.det; out->min_e_abs = th.min_e_abs;
                        }
                    } else {
                        int inst = (int)((int64_t)prim - s.n_top_tris);
                        HitRecord h;
                        if (instance_intersect(s, inst, r, &h, ctr)) { hit = true; *out = h; *inst_out = inst; }
                    }
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            } else {
                if (neg[node.axis] == 1) { stack[to_visit++] = cur + 1; cur = node.offset; }
                else { stack[to_visit++] = node.offset; cur = cur + 1; }
            }
        } else {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return hit;
}
inline bool top_intersect_p(const TopLevel& s, const Ray& r, TraversalCounters* ctr = nullptr) {
    const Accel& a = s.top;
    if (a.nodes.empty()) return false;
    V3 inv_dir(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int neg[3] = {inv_dir.x < 0.0f ? 1 : 0, inv_dir.y < 0.0f ? 1 : 0, inv_dir.z < 0.0f ? 1 : 0};
    size_t to_visit = 0, cur = 0;
    size_t stack[64];
    for (;;) {
        const LinearBVHNode& node = a.nodes[cur];
        if (ctr) ctr->nodes += 1;
        if (bounds_intersect_p_inv(node.bounds, r, inv_dir, neg)) {
            if (node.n_primitives > 0) {
                for (uint32_t i = 0; i < node.n_primitives; ++i) {
                    uint32_t prim = a.ordered[node.offset + i];
                    if ((int64_t)prim < s.n_top_tris) {
                        if (ctr) ctr->tris += 1;
                        if (prim_intersect_p(a, prim, r)) return true;
                    } else if (instance_intersect_p(s, (int)((int64_t)prim - s.n_top_tris), r, ctr)) return true;
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            } else {
                if (neg[node.axis] == 1) { stack[to_visit++] = cur + 1; cur = node.offset; }
                else { stack[to_visit++] = node.offset; cur = cur + 1; }
            }
        } else {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return false;
}

// Triangle::world_bound, triangle.rs:427-431.
inline void triangle_world_bound(V3 p0, V3 p1, V3 p2, Float out[6]) {
    Bounds3 b(p0, p0);
    b = bunion(b, p1);
    b = bunion(b, p2);
    out[0] = b.pmin.x; out[1] = b.pmin.y; out[2] = b.pmin.z;
    out[3] = b.pmax.x; out[4] = b.pmax.y; out[5] = b.pmax.z;
}

}  // namespace orc
