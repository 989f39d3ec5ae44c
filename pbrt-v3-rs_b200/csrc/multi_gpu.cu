// Integrator::render on several GPUs of ONE process (b200pt.h "multi-GPU").
//
// The reference is a single process that hands 16x16 tiles to a thread pool (core/src/integrator/
// sampler_integrator.rs:252-296).  Here the scene is replicated on every GPU, the pixel rows are cut into bands dealt
// in snake order to the devices (b200pt_band_owner, b200pt_render_shard_device), one host thread per device renders its bands into its own
// film, and the bands are gathered on the first device over NVLink:
//   * box-sized filters (radius <= 0.5 px in y): a device's samples only reach its own rows, so each device SENDS ITS
//     OWN BANDS (1 / n of the film) straight into their place in the first device's film - NCCL send / recv grouped
//     per device (ncclGroupStart .. ncclGroupEnd), no reduction, no full-film traffic;
//   * wider filters: neighbouring bands overlap by the filter apron, so the films are summed with one ncclReduce.
// NCCL is loaded at run time (libnccl.so.2; the process may already hold torch's copy); without it, or with
// B200PT_GATHER=peer, the same transfers run as cudaMemcpyPeerAsync over the same links.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b2 {

// ---- NCCL, resolved at first use ------------------------------------------------------------------------------
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
        }
        if (!api.so) return;
        auto sym = [&](const char* n) { return dlsym(api.so, n); };
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv && api.Reduce && api.GetErrorString;
    });
    return api;
}

// Sums `n_src` staged films into dst (peer-copy path with filters wider than a pixel).
__global__ void __launch_bounds__(256) k_film_sum(float4* __restrict__ dst, const float4* __restrict__ src, int n_src, long long n_pix) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float4 a = dst[i];
    for (int k = 0; k < n_src; ++k) {
        const float4 b = src[(long long)k * n_pix + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dst[i] = a;
}

// Box-sized filters: a sample whose film position has a zero fractional part in y also lands in the pixel row ABOVE its
// own (film_tile.rs:73-76: p0 = ceil(p - 0.5 - radius)), so the first sample row of band k contributes to the last pixel
// row of band k - 1, which another device owns.  spill[k] = that row as the owner of band k rendered it (it holds nothing
// but those contributions); it is added after the row's own samples, the order the single-device film kernel uses.
__global__ void __launch_bounds__(256) k_film_add_spill(float4* __restrict__ film, const float4* __restrict__ spill, int band_rows, int n_bands, int cw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)(n_bands - 1) * cw) return;
    const int k = 1 + (int)(i / cw), x = (int)(i % cw);
    const float4 b = spill[(long long)k * cw + x];
    float4& a = film[((long long)k * band_rows - 1) * cw + x];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

}  // namespace b2

struct b200pt_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<b200pt_scene*> scenes;
    std::vector<float4*> d_film;       // per device: the full cropped window
    std::vector<cudaStream_t> stream;  // per device
    float4* d_stage = nullptr;         // first device: (n - 1) staged films (peer path, wide filters)
    float4* d_spill = nullptr;         // first device: one film row per band (box-sized filters, see k_film_add_spill)
    int spill_bands = 0;
    std::vector<ncclComm_t> comms;
    bool nccl = false;
    bool disjoint = true;              // a device's samples only reach its own rows
    b200pt_film film;
    long long n_pix = 0;
    uint64_t rays[3] = {0, 0, 0};
    double gather_ms = 0.0;
    std::mutex mu;
};

using namespace b2;

static int nccl_fail(ncclResult_t r, const char* what) {
    std::string m = std::string("NCCL error in ") + what + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "?");
    b200pt_set_error(m.c_str());
    return B200PT_ERR_CUDA;
}
#define B2_NCCL(call)                                          \
    do {                                                       \
        ncclResult_t _r = (call);                              \
        if (_r != ncclSuccess) return nccl_fail(_r, #call);    \
    } while (0)

extern "C" {

void b200pt_multi_destroy(b200pt_multi* m) {
    if (!m) return;
    for (int i = 0; i < m->n; ++i) {
        if (m->devices[(size_t)i] >= 0) cudaSetDevice(m->devices[(size_t)i]);
        if (i < (int)m->scenes.size() && m->scenes[(size_t)i]) b200pt_scene_destroy(m->scenes[(size_t)i]);
        if (i < (int)m->d_film.size() && m->d_film[(size_t)i]) cudaFree(m->d_film[(size_t)i]);
        if (i < (int)m->stream.size() && m->stream[(size_t)i]) cudaStreamDestroy(m->stream[(size_t)i]);
        if (i < (int)m->comms.size() && m->comms[(size_t)i]) nccl_api().CommDestroy(m->comms[(size_t)i]);
        if (i == 0 && m->d_stage) cudaFree(m->d_stage);
        if (i == 0 && m->d_spill) cudaFree(m->d_spill);
    }
    delete m;
}

int b200pt_multi_create(const b200pt_scene_desc* desc, const int32_t* devices, int32_t n_devices, b200pt_multi** out) {
    if (!out) { b200pt_set_error("b200pt_multi_create: out is null"); return B200PT_ERR_INVALID; }
    *out = nullptr;
    if (!desc || !devices || n_devices < 1 || n_devices > 64) { b200pt_set_error("b200pt_multi_create: invalid argument"); return B200PT_ERR_INVALID; }
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) { b200pt_set_error("b200pt_multi_create: a device is listed twice"); return B200PT_ERR_INVALID; }
    const int before = current_device();
    b200pt_multi* m = new b200pt_multi();
    m->n = n_devices;
    m->devices.assign(devices, devices + n_devices);
    m->scenes.assign((size_t)n_devices, nullptr);
    m->d_film.assign((size_t)n_devices, nullptr);
    m->stream.assign((size_t)n_devices, nullptr);
    m->film = desc->film;
    m->n_pix = (long long)(desc->film.crop[2] - desc->film.crop[0]) * (desc->film.crop[3] - desc->film.crop[1]);
    m->disjoint = desc->film.filter_radius[1] <= 0.5f;
    auto fail = [&](int rc) { std::string keep = b200pt_last_error(); b200pt_multi_destroy(m); if (before >= 0) b200pt_set_device(before); b200pt_set_error(keep.c_str()); return rc; };
    for (int i = 0; i < n_devices; ++i) {
        int rc = b200pt_init(devices[i]);  // registers the device (and binds this thread to it for the moment)
        if (rc) return fail(rc);
    }
    // one host thread per device: scene upload and accelerator records are independent per GPU
    std::vector<int> rcs((size_t)n_devices, 0);
    std::vector<std::string> errs((size_t)n_devices);
    std::vector<std::thread> th;
    for (int i = 0; i < n_devices; ++i)
        th.emplace_back([&, i] {
            int rc = b200pt_set_device(m->devices[(size_t)i]);
            if (!rc) rc = b200pt_scene_create(desc, &m->scenes[(size_t)i]);
            if (!rc) {
                cudaError_t e = cudaMalloc(&m->d_film[(size_t)i], (size_t)std::max<long long>(m->n_pix, 1) * sizeof(float4));
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream[(size_t)i], cudaStreamNonBlocking);
                if (e != cudaSuccess) rc = cuda_fail(e, "b200pt_multi_create");
            }
            if (rc) errs[(size_t)i] = b200pt_last_error();
            rcs[(size_t)i] = rc;
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < n_devices; ++i)
        if (rcs[(size_t)i]) { b200pt_set_error(errs[(size_t)i].c_str()); return fail(rcs[(size_t)i]); }
    // the gather path: NCCL over NVLink / NVSwitch, else peer copies
    const char* g = std::getenv("B200PT_GATHER");
    const bool want_peer = g && std::strcmp(g, "peer") == 0;
    if (n_devices > 1 && !want_peer && nccl_api().ok) {
        m->comms.assign((size_t)n_devices, nullptr);
        ncclResult_t r = nccl_api().CommInitAll(m->comms.data(), n_devices, m->devices.data());
        if (r == ncclSuccess) m->nccl = true;
        else m->comms.clear();
    }
    if (n_devices > 1 && !m->nccl) {
        cudaSetDevice(m->devices[0]);
        for (int i = 1; i < n_devices; ++i) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->devices[0], m->devices[(size_t)i]);
            if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[(size_t)i], 0); if (e != cudaSuccess) cudaGetLastError(); }
        }
        if (!m->disjoint) {
            cudaError_t e = cudaMalloc(&m->d_stage, (size_t)(n_devices - 1) * (size_t)std::max<long long>(m->n_pix, 1) * sizeof(float4));
            if (e != cudaSuccess) return fail(cuda_fail(e, "b200pt_multi_create (staging)"));
        }
    }
    if (before >= 0) b200pt_set_device(before);
    else b200pt_set_device(devices[0]);
    *out = m;
    return B200PT_OK;
}

// Renders the image on all devices; the assembled film {X, Y, Z, weight} ends up in the first device's buffer and, if
// film_xyzw is not NULL, in host memory.
int b200pt_multi_render(b200pt_multi* m, int32_t band_rows, float* film_xyzw) {
    if (!m || band_rows < 1) { b200pt_set_error("b200pt_multi_render: invalid argument"); return B200PT_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(m->mu);
    const int n = m->n;
    const int cw = m->film.crop[2] - m->film.crop[0], ch = m->film.crop[3] - m->film.crop[1];
    const int before = current_device();
    std::vector<int> rcs((size_t)n, 0);
    std::vector<std::string> errs((size_t)n);
    std::vector<double> gather_ms((size_t)n, 0.0);
    std::vector<std::thread> th;
    NcclApi& N = nccl_api();
    std::promise<void> first_ready;                       // peer-copy path: the first device's render is complete
    std::shared_future<void> first_ready_f = first_ready.get_future().share();
    bool first_signalled = false;
    // the spill rows of box-sized filters (one row per band on the first device)
    if (n > 1 && m->disjoint) {
        const int n_bands = (ch + band_rows - 1) / band_rows;
        if (n_bands > m->spill_bands) {
            cudaSetDevice(m->devices[0]);
            if (m->d_spill) cudaFree(m->d_spill);
            m->d_spill = nullptr; m->spill_bands = 0;
            B2_CUDA(cudaMalloc(&m->d_spill, (size_t)n_bands * (size_t)std::max(cw, 1) * sizeof(float4)));
            m->spill_bands = n_bands;
        }
        // rows of bands whose upper neighbour has the same owner are never written: they must add nothing
        cudaSetDevice(m->devices[0]);
        B2_CUDA(cudaMemsetAsync(m->d_spill, 0, (size_t)n_bands * (size_t)std::max(cw, 1) * sizeof(float4), m->stream[0]));
        B2_CUDA(cudaStreamSynchronize(m->stream[0]));
    }
    for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            auto body = [&]() -> int {
                int rc = b200pt_set_device(m->devices[(size_t)i]);
                if (rc) return rc;
                cudaStream_t st = m->stream[(size_t)i];
                // n > 1: the shard's running sums (RGB + weight); they are combined on the first device and converted to XYZ there
                rc = n == 1 ? b200pt_render_shard_device(m->scenes[(size_t)i], i, n, band_rows, m->d_film[(size_t)i], st)
                            : b200pt_render_shard_device_raw(m->scenes[(size_t)i], i, n, band_rows, m->d_film[(size_t)i], st);  // returns with the shard's film complete
                if (rc || n == 1) return rc;
                cudaEvent_t e0, e1;
                B2_CUDA(cudaEventCreate(&e0)); B2_CUDA(cudaEventCreate(&e1));
                B2_CUDA(cudaEventRecord(e0, st));
                const int n_bands = (ch + band_rows - 1) / band_rows;
                if (m->disjoint && i == 0) {
                    // the first device keeps the spill rows of its own bands before other devices' bands land on them
                    for (int k = 1; k < n_bands; ++k)
                        if (b200pt_band_owner(k, n) == 0 && b200pt_band_owner(k - 1, n) != 0) B2_CUDA(cudaMemcpyAsync(m->d_spill + (size_t)k * cw, m->d_film[0] + ((size_t)k * band_rows - 1) * cw, (size_t)cw * sizeof(float4), cudaMemcpyDeviceToDevice, st));
                }
                if (m->nccl && m->disjoint) {
                    // every device ships the bands it owns into their place in the first device's film, and the row above
                    // each of them into the spill buffer (same order on both sides of a pair)
                    B2_NCCL(N.GroupStart());
                    for (int k = 0; k < n_bands; ++k) {
                        const int owner = b200pt_band_owner(k, n), r0 = k * band_rows;
                        const bool spill = k >= 1 && b200pt_band_owner(k - 1, n) != owner;  // the row above belongs to another device
                        const size_t off = (size_t)r0 * cw, cnt = (size_t)(std::min(ch, r0 + band_rows) - r0) * cw * 4;
                        if (i == 0 && owner != 0) {
                            if (spill) B2_NCCL(N.Recv(m->d_spill + (size_t)k * cw, (size_t)cw * 4, ncclFloat, owner, m->comms[0], st));
                            B2_NCCL(N.Recv(m->d_film[0] + off, cnt, ncclFloat, owner, m->comms[0], st));
                        } else if (i != 0 && owner == i) {
                            if (spill) B2_NCCL(N.Send(m->d_film[(size_t)i] + off - (size_t)cw, (size_t)cw * 4, ncclFloat, 0, m->comms[(size_t)i], st));
                            B2_NCCL(N.Send(m->d_film[(size_t)i] + off, cnt, ncclFloat, 0, m->comms[(size_t)i], st));
                        }
                    }
                    B2_NCCL(N.GroupEnd());
                } else if (m->nccl) {
                    // filter aprons overlap: sum the films (in place on the first device)
                    B2_NCCL(N.Reduce(m->d_film[(size_t)i], m->d_film[(size_t)i], (size_t)m->n_pix * 4, ncclFloat, ncclSum, 0, m->comms[(size_t)i], st));
                } else if (i == 0) {
                    // peer copies: the other devices write into this device's film once its own render (and the stash above) is complete
                    B2_CUDA(cudaStreamSynchronize(st));
                    first_signalled = true;
                    first_ready.set_value();
                } else {
                    first_ready_f.wait();
                    if (m->disjoint) {
                        for (int k = 0; k < n_bands; ++k) {
                            if (b200pt_band_owner(k, n) != i) continue;
                            const int r0 = k * band_rows;
                            const bool spill = k >= 1 && b200pt_band_owner(k - 1, n) != i;
                            const size_t off = (size_t)r0 * cw, bytes = (size_t)(std::min(ch, r0 + band_rows) - r0) * cw * sizeof(float4);
                            if (spill) B2_CUDA(cudaMemcpyPeerAsync(m->d_spill + (size_t)k * cw, m->devices[0], m->d_film[(size_t)i] + off - (size_t)cw, m->devices[(size_t)i], (size_t)cw * sizeof(float4), st));
                            B2_CUDA(cudaMemcpyPeerAsync(m->d_film[0] + off, m->devices[0], m->d_film[(size_t)i] + off, m->devices[(size_t)i], bytes, st));
                        }
                    } else {
                        B2_CUDA(cudaMemcpyPeerAsync(m->d_stage + (size_t)(i - 1) * (size_t)m->n_pix, m->devices[0], m->d_film[(size_t)i], m->devices[(size_t)i],
                                                    (size_t)m->n_pix * sizeof(float4), st));
                    }
                }
                B2_CUDA(cudaEventRecord(e1, st));
                B2_CUDA(cudaStreamSynchronize(st));
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, e0, e1);
                gather_ms[(size_t)i] = ms;
                cudaEventDestroy(e0); cudaEventDestroy(e1);
                return B200PT_OK;
            };
            int rc = body();
            if (rc) errs[(size_t)i] = b200pt_last_error();
            rcs[(size_t)i] = rc;
            if (i == 0 && !first_signalled) {  // never leave the other threads waiting (failure, NCCL path, one device)
                try { first_ready.set_value(); } catch (const std::future_error&) {}
            }
        });
    for (auto& t : th) t.join();
    int rc = B200PT_OK;
    for (int i = 0; i < n && !rc; ++i)
        if (rcs[(size_t)i]) { b200pt_set_error(errs[(size_t)i].c_str()); rc = rcs[(size_t)i]; }
    m->rays[0] = m->rays[1] = m->rays[2] = 0;
    m->gather_ms = 0.0;
    for (int i = 0; i < n && !rc; ++i) {
        uint64_t c[3];
        b200pt_scene_ray_counts(m->scenes[(size_t)i], c);
        for (int k = 0; k < 3; ++k) m->rays[k] += c[k];
        m->gather_ms = std::max(m->gather_ms, gather_ms[(size_t)i]);
    }
    if (!rc) {
        cudaSetDevice(m->devices[0]);
        if (n > 1 && m->disjoint) {
            const int n_bands = (ch + band_rows - 1) / band_rows;
            if (n_bands > 1) {
                k_film_add_spill<<<(unsigned)(((long long)(n_bands - 1) * cw + 255) / 256), 256, 0, m->stream[0]>>>(m->d_film[0], m->d_spill, band_rows, n_bands, cw);
                g_launches.fetch_add(1);
            }
        }
        if (n > 1 && !m->nccl && !m->disjoint) {
            k_film_sum<<<(unsigned)((m->n_pix + 255) / 256), 256, 0, m->stream[0]>>>(m->d_film[0], m->d_stage, n - 1, m->n_pix);
            g_launches.fetch_add(1);
        }
        if (n > 1) rc = b200pt_film_finish_device(m->d_film[0], m->n_pix, m->d_film[0], m->stream[0]);
        cudaError_t e = cudaSuccess;
        if (film_xyzw) e = cudaMemcpyAsync(film_xyzw, m->d_film[0], (size_t)m->n_pix * sizeof(float4), cudaMemcpyDeviceToHost, m->stream[0]);
        if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream[0]);
        if (e != cudaSuccess) rc = cuda_fail(e, "b200pt_multi_render (film download)");
    }
    if (before >= 0) b200pt_set_device(before);
    return rc;
}

int b200pt_multi_info(const b200pt_multi* m, uint64_t rays[3], double* gather_ms, int32_t* uses_nccl) {
    if (!m) { b200pt_set_error("b200pt_multi_info: null handle"); return B200PT_ERR_INVALID; }
    if (rays) { rays[0] = m->rays[0]; rays[1] = m->rays[1]; rays[2] = m->rays[2]; }
    if (gather_ms) *gather_ms = m->gather_ms;
    if (uses_nccl) *uses_nccl = m->nccl ? 1 : 0;
    return B200PT_OK;
}

/* the device-side film of the first device after b200pt_multi_render (full cropped window, 4 floats per pixel) */
void* b200pt_multi_film_device(const b200pt_multi* m) { return m && !m->d_film.empty() ? (void*)m->d_film[0] : nullptr; }

int b200pt_render_multi(const b200pt_scene_desc* desc, const int32_t* devices, int32_t n_devices, int32_t band_rows, float* film_xyzw) {
    if (!film_xyzw) { b200pt_set_error("b200pt_render_multi: film_xyzw is null"); return B200PT_ERR_INVALID; }
    b200pt_multi* m = nullptr;
    int rc = b200pt_multi_create(desc, devices, n_devices, &m);
    if (rc) return rc;
    rc = b200pt_multi_render(m, band_rows, film_xyzw);
    std::string keep = rc ? b200pt_last_error() : "";
    b200pt_multi_destroy(m);
    if (rc) b200pt_set_error(keep.c_str());
    return rc;
}

}  // extern "C"
