// Shared host-side plumbing of libb200pt: error reporting, launch counting,
// device buffers.  Internal header (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200pt.h"
#include "traverse.cuh"

extern "C" int b200pt_set_error(const char* msg);

namespace b2 {

extern std::atomic<int64_t> g_launches;

// Per-device library state.  A handle (accelerator, scene) remembers the device it was created on and every entry point
// switches to it, so one process can drive several GPUs (b200pt_render_multi); nothing below is shared across devices.
struct DevCtx {
    int device = -1;
    int sm_count = 0;
    int64_t l2_bytes = 0;
    // 8-byte work counters of the persistent traversal kernels: one per launch in flight, handed out round-robin
    // (the wave loop passes its own counters instead; the ring only serves the *_batch_device entry points)
    static const int kRing = 4096;
    unsigned long long* ring = nullptr;
    std::atomic<unsigned> ring_next{0};
    int persist_grid[4] = {0, 0, 0, 0};  // resident CTAs of the persistent kernels (traverse_kernels.cu)
    std::mutex mu;                        // first-use setup of the fields above
};
int current_device();            // the calling thread's device: b200pt_set_device, else the first b200pt_init; -1 = none
DevCtx* dev_ctx(int device);     // nullptr unless b200pt_init(device) succeeded
int use_device(int device);      // cudaSetDevice for a handle's device; B200PT_ERR_NO_DEVICE if it was never initialised

int cuda_fail(cudaError_t e, const char* what);  // records message, returns B200PT_ERR_CUDA / OOM

#define B2_CUDA(call)                                      \
    do {                                                   \
        cudaError_t _e = (call);                           \
        if (_e != cudaSuccess) return b2::cuda_fail(_e, #call); \
    } while (0)

// Host-side loop over [0, n) cut into one contiguous range per hardware thread (scene upload preparation: tens of
// millions of independent records).  f(begin, end) must only write its own range.
template <class F> inline void parallel_for(int64_t n, F f, int64_t grain = 1 << 16) {
    int64_t want = (n + grain - 1) / grain;
    unsigned hw = std::thread::hardware_concurrency();
    int64_t nt = std::min<int64_t>(want, hw ? hw : 1);
    if (nt <= 1) { if (n > 0) f((int64_t)0, n); return; }
    std::vector<std::thread> th;
    const int64_t chunk = (n + nt - 1) / nt;
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=] { f(b, e); });
    }
    for (auto& x : th) x.join();
}

// Binds the calling thread to its current device (creation entry points and the handle-less *_gpu / *_device calls).
inline int require_device() {
    const int d = current_device();
    if (d < 0) {
        b200pt_set_error("b200pt: no device bound (call b200pt_init first; there is no CPU fallback)");
        return B200PT_ERR_NO_DEVICE;
    }
    return use_device(d);
}

// Host mirror of the device accelerator: owns the device allocations.
struct AccelImpl {
    DeviceAccel dev;
    int device = -1;
    float4* d_wide = nullptr;
    float4* d_tris = nullptr;
    float4* d_ref = nullptr;
    int64_t n_nodes = 0, n_prims = 0, n_wide = 0;
    float world_bound[6];
};

int accel_build_device(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered, const float* tri_verts,
                       const uint32_t* flags, int64_t n_prims, AccelImpl* out, const float* tri_uvs = nullptr);
// Alpha-mask textures on the device (alpha_tex.cuh): uploads the texture table, the per-primitive texture indices and
// the absolute uvs the evaluation needs; the caller stores out->dev into DeviceAccel::alpha.
struct AlphaImpl {
    const DeviceAlpha* dev = nullptr;  // device copy of the table
    std::vector<void*> allocs;
};
int alpha_build_device(const b200pt_float_texture* tex, int32_t n_tex, const int32_t* prim_alpha_tex, const float* tri_uvs, const uint32_t* prim_flags,
                       int64_t n_prims, const uint8_t* noise_perm, AlphaImpl* out);
void alpha_free_device(AlphaImpl* a);
// Depth of a flattened BVH (root = 0; -1 for a malformed array).  The walks keep one pending far child per level: the
// reference's `nodes_to_visit: [usize; 64]` (mod.rs:185) overflows past 64 levels, the device stacks (B2_STACK, B2_STACK2) too.
int bvh_max_depth(const b200pt_bvh_node* nodes, int64_t n_nodes);
float4 record_duv(const float* uv6);  // uv0 - uv2, uv1 - uv2; uv6 == nullptr: the default uvs
void accel_free_device(AccelImpl* a);

// Kernel launchers (traverse_kernels.cu).  TraceLaunch lets the wavefront loop keep its queue sizes on the device:
// with n_dev set the kernel reads the ray count from device memory (n is then only an upper bound for the grid) and
// work_ctr is a zeroed 8-byte counter the caller owns for this launch.
struct TraceLaunch {
    const int* n_dev = nullptr;
    unsigned long long* work_ctr = nullptr;
};
int launch_intersect(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, int variant, float* d_b2 = nullptr,
                     const TraceLaunch* tl = nullptr);
int launch_occluded(const DeviceAccel& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant, const TraceLaunch* tl = nullptr);
int launch_count_work(const DeviceAccel& A, const void* d_rays, int64_t n, int any_hit, unsigned long long* d_totals, void* d_per_ray, cudaStream_t s);

}  // namespace b2

struct b200pt_accel {
    b2::AccelImpl impl;
    b2::AlphaImpl alpha;
};
