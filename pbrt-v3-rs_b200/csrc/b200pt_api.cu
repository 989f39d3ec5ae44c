// C ABI: library init/errors and the accelerator (impl Primitive for BVHAccel,
// accelerators/src/bvh/mod.rs:156-283) entry points.  See include/b200pt.h.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_envmap.h"

namespace b2 {
std::atomic<int64_t> g_launches{0};
static const int kMaxDevices = 64;
static DevCtx* g_ctx[kMaxDevices] = {};
static std::mutex g_ctx_mu;
static std::atomic<int> g_default_device{-1};  // the first device b200pt_init bound
static thread_local int t_device = -1;         // b200pt_init / b200pt_set_device of this thread
static thread_local std::string t_error;

int current_device() { return t_device >= 0 ? t_device : g_default_device.load(); }
DevCtx* dev_ctx(int device) { return device >= 0 && device < kMaxDevices ? g_ctx[device] : nullptr; }
int use_device(int device) {
    if (!dev_ctx(device)) {
        b200pt_set_error("b200pt: device was not initialised (call b200pt_init(device) first; there is no CPU fallback)");
        return B200PT_ERR_NO_DEVICE;
    }
    B2_CUDA(cudaSetDevice(device));
    return B200PT_OK;
}

int cuda_fail(cudaError_t e, const char* what) {
    std::string m = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
    b200pt_set_error(m.c_str());
    cudaGetLastError();  // clear sticky-less errors
    return e == cudaErrorMemoryAllocation ? B200PT_ERR_OOM : B200PT_ERR_CUDA;
}

// Host scratch for the host-buffer batch calls: three chunk slots so the H2D
// copy, the kernel and the D2H copy of consecutive chunks overlap on their own
// streams (copies use the two DMA engines, PCIe is full duplex).
struct BatchScratch {
    static const int kSlots = 3;
    static const int64_t kChunk = 1 << 20;  // rays per chunk (32 MB in, 16 MB out)
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    void* d_rays[kSlots] = {nullptr, nullptr, nullptr};
    void* d_out[kSlots] = {nullptr, nullptr, nullptr};
    bool ready = false;
    std::mutex mu;
    int ensure() {
        if (ready) return B200PT_OK;
        for (int i = 0; i < kSlots; ++i) {
            B2_CUDA(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
            B2_CUDA(cudaMalloc(&d_rays[i], kChunk * sizeof(b200pt_ray)));
            B2_CUDA(cudaMalloc(&d_out[i], kChunk * sizeof(b200pt_hit)));
        }
        ready = true;
        return B200PT_OK;
    }
};
static BatchScratch g_scratch[kMaxDevices];

// Fourth float4 of a triangle record: uv0 - uv2, uv1 - uv2 (get_uvs, triangle.rs:384-394, 551-552).
float4 record_duv(const float* uv6) {
    if (!uv6) return make_float4(0.0f - 1.0f, 0.0f - 1.0f, 1.0f - 1.0f, 0.0f - 1.0f);
    return make_float4(uv6[0] - uv6[4], uv6[1] - uv6[5], uv6[2] - uv6[4], uv6[3] - uv6[5]);
}

int bvh_max_depth(const b200pt_bvh_node* nodes, int64_t n_nodes) {
    if (n_nodes <= 0) return 0;
    std::vector<int32_t> depth((size_t)n_nodes, -1);
    depth[0] = 0;
    int best = 0;
    for (int64_t i = 0; i < n_nodes; ++i) {  // pre-order: a node's depth is known before its children are reached
        const int32_t d = depth[(size_t)i];
        if (d < 0) return -1;
        best = std::max(best, (int)d);
        if (nodes[i].n_primitives != 0) continue;
        const int64_t c0 = i + 1, c1 = nodes[i].offset;
        if (c0 >= n_nodes || c1 <= i || c1 >= n_nodes) return -1;
        depth[(size_t)c0] = d + 1; depth[(size_t)c1] = d + 1;
    }
    return best;
}

int accel_build_device(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered, const float* tri_verts,
                       const uint32_t* flags, int64_t n_prims, AccelImpl* a, const float* tri_uvs) {
    if (n_nodes > 0) {
        const int depth = bvh_max_depth(nodes, n_nodes);
        if (depth < 0) { b200pt_set_error("b200pt_accel_create: malformed node array"); return B200PT_ERR_INVALID; }
        if (depth > B2_STACK) {
            b200pt_set_error("b200pt_accel_create: the BVH is more than 64 levels deep; BVHAccel::intersect's fixed traversal stack (accelerators/src/bvh/mod.rs:185) overflows on such a tree too");
            return B200PT_ERR_UNSUPPORTED;
        }
    }
    a->n_nodes = n_nodes;
    a->n_prims = n_prims;
    std::memset(&a->dev, 0, sizeof(a->dev));
    a->device = a->dev.device = current_device();
    a->dev.root_code = B2_EMPTY_ROOT;
    a->dev.n_nodes = (int)n_nodes;
    a->dev.n_prims = n_prims;
    const float m = 3.402823466e+38f;  // Bounds3f::EMPTY, bounds3.rs:26-29
    float empty[6] = {m, m, m, -m, -m, -m};
    std::memcpy(a->world_bound, empty, sizeof(empty));
    if (n_nodes == 0) return B200PT_OK;
    std::memcpy(a->world_bound, nodes[0].bounds, sizeof(empty));
    std::memcpy(a->dev.root_bounds, nodes[0].bounds, sizeof(empty));

    // wide index of every interior reference node, in pre-order
    std::vector<int32_t> wide_of((size_t)n_nodes, -1);
    int64_t n_wide = 0;
    for (int64_t i = 0; i < n_nodes; ++i)
        if (nodes[i].n_primitives == 0) wide_of[(size_t)i] = (int32_t)n_wide++;
    a->n_wide = n_wide;
    auto code_of = [&](int64_t i) -> int32_t {
        return nodes[i].n_primitives == 0 ? wide_of[(size_t)i] : ~(int32_t)nodes[i].offset;
    };
    std::vector<float4> wide((size_t)std::max<int64_t>(n_wide, 1) * 4);
    std::atomic<int> bad{0};
    parallel_for(n_nodes, [&](int64_t i_begin, int64_t i_end) {
    for (int64_t i = i_begin; i < i_end; ++i) {
        if (nodes[i].n_primitives != 0) continue;
        int64_t c0 = i + 1, c1 = nodes[i].offset;
        if (c1 <= i || c1 >= n_nodes || c0 >= n_nodes) { bad = 1; continue; }
        const float* a0 = nodes[c0].bounds;
        const float* a1 = nodes[c1].bounds;
        float4* q = &wide[(size_t)wide_of[(size_t)i] * 4];
        q[0] = make_float4(a0[0], a0[1], a0[2], a0[3]);
        q[1] = make_float4(a0[4], a0[5], a1[0], a1[1]);
        q[2] = make_float4(a1[2], a1[3], a1[4], a1[5]);
        int32_t k0 = code_of(c0), k1 = code_of(c1), ax = nodes[i].axis;
        float f0, f1, f2;
        std::memcpy(&f0, &k0, 4); std::memcpy(&f1, &k1, 4); std::memcpy(&f2, &ax, 4);
        q[3] = make_float4(f0, f1, f2, 0.0f);
    }
    });
    if (bad) { b200pt_set_error("b200pt_accel_create: malformed node array"); return B200PT_ERR_INVALID; }
    a->dev.root_code = code_of(0);

    // triangles in BVHAccel.primitives order
    std::vector<float4> tris((size_t)std::max<int64_t>(n_prims, 1) * 4);
    parallel_for(n_prims, [&](int64_t j_begin, int64_t j_end) {
    for (int64_t j = j_begin; j < j_end; ++j) {
        uint32_t p = ordered[j];
        if ((int64_t)p >= n_prims) { bad = 2; continue; }
        const float* v = tri_verts + 9 * (size_t)p;
        uint32_t fl = flags ? flags[p] : 0u, zero = 0u;
        float fp, ff, fz;
        std::memcpy(&fp, &p, 4); std::memcpy(&ff, &fl, 4); std::memcpy(&fz, &zero, 4);
        tris[(size_t)j * 4 + 0] = make_float4(v[0], v[1], v[2], v[3]);
        tris[(size_t)j * 4 + 1] = make_float4(v[4], v[5], v[6], v[7]);
        tris[(size_t)j * 4 + 2] = make_float4(v[8], fp, ff, fz);
        tris[(size_t)j * 4 + 3] = record_duv(tri_uvs && (fl & B200PT_PRIM_HAS_UV) ? tri_uvs + 6 * (size_t)p : nullptr);
    }
    });
    if (bad) { b200pt_set_error("b200pt_accel_create: ordered_prims index out of range"); return B200PT_ERR_INVALID; }
    parallel_for(n_nodes, [&](int64_t i_begin, int64_t i_end) {
    for (int64_t i = i_begin; i < i_end; ++i) {
        if (nodes[i].n_primitives == 0) continue;
        uint32_t cnt = nodes[i].n_primitives;
        if ((int64_t)nodes[i].offset + cnt > n_prims) { bad = 3; continue; }
        float fc;
        std::memcpy(&fc, &cnt, 4);
        tris[(size_t)nodes[i].offset * 4 + 2].w = fc;  // leaves own disjoint triangle ranges
    }
    });
    if (bad) { b200pt_set_error("b200pt_accel_create: leaf range out of bounds"); return B200PT_ERR_INVALID; }
    B2_CUDA(cudaMalloc(&a->d_wide, wide.size() * sizeof(float4)));
    B2_CUDA(cudaMalloc(&a->d_tris, tris.size() * sizeof(float4)));
    B2_CUDA(cudaMalloc(&a->d_ref, (size_t)n_nodes * 32));
    B2_CUDA(cudaMemcpy(a->d_wide, wide.data(), wide.size() * sizeof(float4), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(a->d_tris, tris.data(), tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(a->d_ref, nodes, (size_t)n_nodes * 32, cudaMemcpyHostToDevice));
    a->dev.wide = a->d_wide;
    a->dev.tris = a->d_tris;
    a->dev.ref_nodes = a->d_ref;
    return B200PT_OK;
}

void accel_free_device(AccelImpl* a) {
    if (a->d_wide) cudaFree(a->d_wide);
    if (a->d_tris) cudaFree(a->d_tris);
    if (a->d_ref) cudaFree(a->d_ref);
    a->d_wide = a->d_tris = a->d_ref = nullptr;
}

void alpha_free_device(AlphaImpl* a) {
    for (void* p : a->allocs) cudaFree(p);
    a->allocs.clear();
    a->dev = nullptr;
}

int alpha_build_device(const b200pt_float_texture* tex, int32_t n_tex, const int32_t* prim_alpha_tex, const float* tri_uvs, const uint32_t* prim_flags,
                       int64_t n_prims, const uint8_t* noise_perm, AlphaImpl* out) {
    alpha_free_device(out);
    bool any = false;
    if (prim_flags)
        for (int64_t i = 0; i < n_prims && !any; ++i) any = (prim_flags[i] & B200PT_PRIM_ALPHA_TEXTURE) != 0;
    if (!any) return B200PT_OK;
    if (!tex || n_tex <= 0 || !prim_alpha_tex) {
        b200pt_set_error("alpha textures: a primitive carries B200PT_PRIM_ALPHA_TEXTURE but float_textures / prim_alpha_tex are missing");
        return B200PT_ERR_INVALID;
    }
    bool dots = false;
    std::vector<DFloatTex> dt((size_t)n_tex);
    auto upload = [&](const void* src, size_t bytes, void** dst) -> int {
        B2_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 16)));
        out->allocs.push_back(*dst);
        B2_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return B200PT_OK;
    };
    int rc;
    for (int32_t k = 0; k < n_tex; ++k) {
        const b200pt_float_texture& t = tex[k];
        DFloatTex& d = dt[(size_t)k];
        if (t.type < B200PT_TEX_CONSTANT || t.type > B200PT_TEX_IMAGEMAP) { b200pt_set_error("alpha textures: unknown float texture type (constant, checkerboard, dots, imagemap are on this path)"); return B200PT_ERR_UNSUPPORTED; }
        d.type = t.type; d.su = t.su; d.sv = t.sv; d.du = t.du; d.dv = t.dv; d.v0 = t.value[0]; d.v1 = t.value[1];
        d.wrap = t.wrap; d.width = t.width; d.height = t.height; d.texels = nullptr;
        if (t.type == B200PT_TEX_DOTS) dots = true;
        if (t.type == B200PT_TEX_IMAGEMAP) {
            if (!t.texels || t.width <= 0 || t.height <= 0 || t.wrap < 0 || t.wrap > 2) { b200pt_set_error("alpha textures: imagemap without texels / with an unknown wrap mode"); return B200PT_ERR_INVALID; }
            void* p = nullptr;
            if ((rc = upload(t.texels, (size_t)t.width * t.height * sizeof(float), &p))) return rc;
            d.texels = (const float*)p;
        }
    }
    if (dots && !noise_perm) { b200pt_set_error("alpha textures: \"dots\" needs noise_perm (NOISE_PERM[0..256) of core/src/texture/common.rs)"); return B200PT_ERR_INVALID; }
    std::vector<float> uv((size_t)n_prims * 6);
    std::vector<int32_t> pt((size_t)n_prims * 2, -1);
    for (int64_t i = 0; i < n_prims; ++i) {
        const uint32_t fl = prim_flags[i];
        float* o = &uv[(size_t)i * 6];
        if (tri_uvs && (fl & B200PT_PRIM_HAS_UV)) std::memcpy(o, tri_uvs + 6 * (size_t)i, 6 * sizeof(float));
        else { o[0] = 0.0f; o[1] = 0.0f; o[2] = 1.0f; o[3] = 0.0f; o[4] = 1.0f; o[5] = 1.0f; }  // Triangle::get_uvs, triangle.rs:384-394
        if (!(fl & B200PT_PRIM_ALPHA_TEXTURE)) continue;
        for (int c = 0; c < 2; ++c) {
            const int32_t v = prim_alpha_tex[2 * i + c];
            if (v < -1 || v >= n_tex) { b200pt_set_error("alpha textures: prim_alpha_tex index out of range"); return B200PT_ERR_INVALID; }
            pt[(size_t)i * 2 + c] = v;
        }
    }
    void *d_tex = nullptr, *d_uv = nullptr, *d_pt = nullptr, *d_perm = nullptr;
    if ((rc = upload(dt.data(), dt.size() * sizeof(DFloatTex), &d_tex))) return rc;
    if ((rc = upload(uv.data(), uv.size() * sizeof(float), &d_uv))) return rc;
    if ((rc = upload(pt.data(), pt.size() * sizeof(int32_t), &d_pt))) return rc;
    if (noise_perm && (rc = upload(noise_perm, 256, &d_perm))) return rc;
    DeviceAlpha h;
    h.tex = (const DFloatTex*)d_tex; h.uv = (const float*)d_uv; h.prim_tex = (const int*)d_pt; h.perm = (const unsigned char*)d_perm;
    void* d_h = nullptr;
    if ((rc = upload(&h, sizeof(h), &d_h))) return rc;
    out->dev = (const DeviceAlpha*)d_h;
    return B200PT_OK;
}

// Pipelined host-buffer batch (the e2e path): chunk i uses slot i % 3.
template <class LaunchFn>
static int run_host_batch(int device, const void* rays, int64_t n, void* out, size_t out_elem, LaunchFn launch) {
    BatchScratch& g_scratch = b2::g_scratch[device];
    std::lock_guard<std::mutex> g(g_scratch.mu);
    int rc = g_scratch.ensure();
    if (rc) return rc;
    const int64_t C = BatchScratch::kChunk;
    int64_t n_chunks = (n + C - 1) / C;
    for (int64_t c = 0; c < n_chunks; ++c) {
        int slot = (int)(c % BatchScratch::kSlots);
        int64_t b = c * C, m = std::min<int64_t>(C, n - b);
        cudaStream_t s = g_scratch.stream[slot];
        B2_CUDA(cudaMemcpyAsync(g_scratch.d_rays[slot], (const char*)rays + b * sizeof(b200pt_ray), m * sizeof(b200pt_ray),
                                cudaMemcpyHostToDevice, s));
        rc = launch(g_scratch.d_rays[slot], m, g_scratch.d_out[slot], s);
        if (rc) return rc;
        B2_CUDA(cudaMemcpyAsync((char*)out + b * out_elem, g_scratch.d_out[slot], m * out_elem, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < BatchScratch::kSlots; ++i) B2_CUDA(cudaStreamSynchronize(g_scratch.stream[i]));
    return B200PT_OK;
}

}  // namespace b2

using namespace b2;

extern "C" {

int b200pt_set_error(const char* msg) {
    t_error = msg ? msg : "";
    return 0;
}
const char* b200pt_last_error(void) { return t_error.c_str(); }
int b200pt_version(void) { return 100; }
int b200pt_device_sm_count(void) { DevCtx* c = dev_ctx(current_device()); return c ? c->sm_count : 0; }
int64_t b200pt_device_l2_bytes(void) { DevCtx* c = dev_ctx(current_device()); return c ? c->l2_bytes : 0; }
int b200pt_current_device(void) { return current_device(); }
int64_t b200pt_launch_count(void) { return g_launches.load(); }

int b200pt_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        b200pt_set_error("b200pt_init: no CUDA device visible (this library has no CPU fallback)");
        return B200PT_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { b200pt_set_error("b200pt_init: device index out of range"); return B200PT_ERR_INVALID; }
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        b200pt_set_error("b200pt_init: device is not sm_100 (kernels are built for sm_100a only)");
        return B200PT_ERR_NO_DEVICE;
    }
    if (device >= kMaxDevices) { b200pt_set_error("b200pt_init: device index above the library's table"); return B200PT_ERR_INVALID; }
    B2_CUDA(cudaSetDevice(device));
    {
        std::lock_guard<std::mutex> g(g_ctx_mu);
        if (!g_ctx[device]) {
            DevCtx* c = new DevCtx();
            c->device = device;
            c->sm_count = prop.multiProcessorCount;
            c->l2_bytes = prop.l2CacheSize;
            g_ctx[device] = c;
        }
    }
    int expected = -1;
    g_default_device.compare_exchange_strong(expected, device);
    t_device = device;
    return B200PT_OK;
}

int b200pt_set_device(int device) {
    if (!dev_ctx(device)) { b200pt_set_error("b200pt_set_device: device was not initialised with b200pt_init"); return B200PT_ERR_NO_DEVICE; }
    t_device = device;
    return use_device(device);
}

int b200pt_envmap_prepare(const float* map_rgb, int32_t map_width, int32_t map_height, const float L[3], int32_t size4[4],
                          float* level0_rgb_out, float* importance_out, float power_lookup_out[3]) {
    if (!L || !size4 || map_width < 0 || map_height < 0) { b200pt_set_error("b200pt_envmap_prepare: invalid argument"); return B200PT_ERR_INVALID; }
    b2host::EnvMapTables em;
    b2host::build_envmap(map_rgb, map_width, map_height, L, &em);
    size4[0] = em.width; size4[1] = em.height; size4[2] = em.nu; size4[3] = em.nv;
    if (level0_rgb_out)
        for (size_t k = 0; k < (size_t)em.width * em.height; ++k)
            for (int c = 0; c < 3; ++c) level0_rgb_out[3 * k + c] = em.texels[4 * k + c];
    if (importance_out) std::memcpy(importance_out, em.cond_func.data(), em.cond_func.size() * sizeof(float));
    if (power_lookup_out) std::memcpy(power_lookup_out, em.power_lookup, 12);
    return B200PT_OK;
}

int b200pt_accel_create(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered_prims, const float* tri_verts,
                        const uint32_t* prim_flags, int64_t n_prims, b200pt_accel** out) {
    return b200pt_accel_create_uv(nodes, n_nodes, ordered_prims, tri_verts, nullptr, prim_flags, n_prims, out);
}

int b200pt_accel_create_uv(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered_prims, const float* tri_verts,
                           const float* tri_uvs, const uint32_t* prim_flags, int64_t n_prims, b200pt_accel** out) {
    if (!out) { b200pt_set_error("b200pt_accel_create: out is null"); return B200PT_ERR_INVALID; }
    *out = nullptr;
    int rc = require_device();
    if (rc) return rc;
    if (n_nodes < 0 || n_prims < 0 || (n_nodes > 0 && (!nodes || !ordered_prims || !tri_verts)) || n_nodes > 0x7ffffff0LL) {
        b200pt_set_error("b200pt_accel_create: invalid argument");
        return B200PT_ERR_INVALID;
    }
    b200pt_accel* a = new b200pt_accel();
    rc = accel_build_device(nodes, n_nodes, ordered_prims, tri_verts, prim_flags, n_prims, &a->impl, tri_uvs);
    if (rc) { accel_free_device(&a->impl); delete a; return rc; }
    *out = a;
    return B200PT_OK;
}

int b200pt_accel_set_alpha_textures(b200pt_accel* a, const b200pt_float_texture* float_textures, int32_t n_float_textures,
                                    const int32_t* prim_alpha_tex, const float* tri_uvs, const uint32_t* prim_flags, const uint8_t* noise_perm) {
    if (!a) { b200pt_set_error("b200pt_accel_set_alpha_textures: null accelerator"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    rc = alpha_build_device(float_textures, n_float_textures, prim_alpha_tex, tri_uvs, prim_flags, a->impl.n_prims, noise_perm, &a->alpha);
    a->impl.dev.alpha = a->alpha.dev;
    return rc;
}

void b200pt_accel_destroy(b200pt_accel* a) {
    if (!a) return;
    alpha_free_device(&a->alpha);
    accel_free_device(&a->impl);
    delete a;
}

int b200pt_accel_world_bound(const b200pt_accel* a, float* bounds6) {
    if (!a || !bounds6) { b200pt_set_error("b200pt_accel_world_bound: null argument"); return B200PT_ERR_INVALID; }
    std::memcpy(bounds6, a->impl.world_bound, 6 * sizeof(float));
    return B200PT_OK;
}

int b200pt_intersect_batch_device(const b200pt_accel* a, const void* d_rays, int64_t n, void* d_hits, void* stream, int variant) {
    if (!a || n < 0 || (n > 0 && (!d_rays || !d_hits))) { b200pt_set_error("b200pt_intersect_batch_device: invalid argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    return launch_intersect(a->impl.dev, d_rays, n, d_hits, (cudaStream_t)stream, variant);
}

int b200pt_occluded_batch_device(const b200pt_accel* a, const void* d_rays, int64_t n, void* d_occluded, void* stream, int variant) {
    if (!a || n < 0 || (n > 0 && (!d_rays || !d_occluded))) { b200pt_set_error("b200pt_occluded_batch_device: invalid argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    return launch_occluded(a->impl.dev, d_rays, n, d_occluded, (cudaStream_t)stream, variant);
}

int b200pt_count_work_device(const b200pt_accel* a, const void* d_rays, int64_t n, int any_hit, uint64_t totals[2], void* d_per_ray) {
    if (!a || n < 0 || !totals || (n > 0 && !d_rays)) { b200pt_set_error("b200pt_count_work_device: invalid argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    unsigned long long* d_tot = nullptr;
    B2_CUDA(cudaMalloc(&d_tot, 16));
    B2_CUDA(cudaMemset(d_tot, 0, 16));
    rc = launch_count_work(a->impl.dev, d_rays, n, any_hit, d_tot, d_per_ray, 0);
    if (rc) { cudaFree(d_tot); return rc; }
    cudaError_t e = cudaMemcpy(totals, d_tot, 16, cudaMemcpyDeviceToHost);
    cudaFree(d_tot);
    if (e != cudaSuccess) return cuda_fail(e, "count_work copy");
    return B200PT_OK;
}

int b200pt_intersect_batch(const b200pt_accel* a, const b200pt_ray* rays, int64_t n, b200pt_hit* hits) {
    if (!a || n < 0 || (n > 0 && (!rays || !hits))) { b200pt_set_error("b200pt_intersect_batch: invalid argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    const DeviceAccel& A = a->impl.dev;
    return run_host_batch(a->impl.device, rays, n, hits, sizeof(b200pt_hit),
                          [&](void* dr, int64_t m, void* dout, cudaStream_t s) { return launch_intersect(A, dr, m, dout, s, 0); });
}

int b200pt_occluded_batch(const b200pt_accel* a, const b200pt_ray* rays, int64_t n, uint8_t* occluded) {
    if (!a || n < 0 || (n > 0 && (!rays || !occluded))) { b200pt_set_error("b200pt_occluded_batch: invalid argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(a->impl.device);
    if (rc) return rc;
    const DeviceAccel& A = a->impl.dev;
    return run_host_batch(a->impl.device, rays, n, occluded, 1,
                          [&](void* dr, int64_t m, void* dout, cudaStream_t s) { return launch_occluded(A, dr, m, dout, s, 0); });
}

int b200pt_accel_intersect1(const b200pt_accel* a, b200pt_ray* ray, b200pt_hit* hit) {
    if (!ray || !hit) { b200pt_set_error("b200pt_accel_intersect1: null argument"); return B200PT_ERR_INVALID; }
    int rc = b200pt_intersect_batch(a, ray, 1, hit);
    if (rc) return rc;
    if (hit->prim != B200PT_MISS) ray->tmax = hit->t;  // Primitive::intersect lowers r.t_max (geometric_primitive.rs:72)
    return B200PT_OK;
}

int b200pt_accel_occluded1(const b200pt_accel* a, const b200pt_ray* ray, uint8_t* occluded) {
    if (!ray || !occluded) { b200pt_set_error("b200pt_accel_occluded1: null argument"); return B200PT_ERR_INVALID; }
    return b200pt_occluded_batch(a, ray, 1, occluded);
}

}  // extern "C"
