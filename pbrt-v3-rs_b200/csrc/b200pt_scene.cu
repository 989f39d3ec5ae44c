// Scene + PathIntegrator entry points: the wavefront path tracer.
//
// Reference call stack replaced:
//   SamplerIntegrator::render / render_tile   core/src/integrator/sampler_integrator.rs:243-415
//   PathIntegrator::li                        integrators/src/path.rs:103-284
//   uniform_sample_one_light, estimate_direct core/src/integrator/common.rs:89-299
//   FilmTile::add_sample, Film::merge/write   core/src/film/film_tile.rs:62-108, film/mod.rs:220-417
//
// Wavefront organisation (one wave = up to wave_cap_for() paths, a path = one (pixel, sample)):
//   K1 k_raygen     Halton dims 0-4 -> camera ray, path state init
//   K2a closest-hit over the compacted ray queue            (traverse_kernels.cu)
//   K4 k_shade      emission, BSDF frame, light pick + sample_li + BSDF MIS sample -> shadow / MIS ray
//                   queues with pending contributions; BSDF sample -> next ray queue; Russian roulette
//   K2b any-hit over the shadow queue, K2a closest-hit over the MIS queue
//   K4' k_resolve   L += beta * (ld_light [if unoccluded] + ld_mis [if the MIS ray reached the light]) / pick_pdf
//   ... next bounce on the compacted survivors ...
//   K5 k_film       per pixel, gathers the per-sample radiances in pixel-major / sample order
//                   (atomic-free, deterministic) and applies the filter table.
// Per path the order of floating-point accumulation into L is the reference's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_sampler.h"
#include "host_envmap.h"
#include "instancing.cuh"
#include "shade.cuh"
#include "wavefront.cuh"
#include "shade_tree.cuh"

namespace b2 {

// ---- K1: camera rays --------------------------------------------------------------------------
// Implicit mode (list == nullptr): path p of the wave covers sample (first_sample + p) in pixel-major
// order over the shard's sample rows: global sample g -> pixel g / spp, sample g % spp.
// Explicit mode: list holds (x, y, sample) triples.
__global__ void __launch_bounds__(256) k_raygen(DeviceScene S, Wave W, long long first_sample, int n, int spp, const int* __restrict__ rows, const int* __restrict__ list,
                                                 float2* __restrict__ p_film_out, float4* __restrict__ rays_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int px, py, s;
    long long g = first_sample + p;
    if (list) { px = list[3 * p]; py = list[3 * p + 1]; s = list[3 * p + 2]; }
    else {
        long long pix = g / spp;
        s = (int)(g - pix * spp);
        int w = S.sb[2] - S.sb[0];
        px = S.sb[0] + (int)(pix % w);
        py = rows[(int)(pix / w)];
    }
    unsigned long long idx;
    if (S.sampler_type == B200PT_SAMPLER_HALTON) idx = halton_index(S.halton, px, py, (unsigned long long)s);
    else if (S.sampler_type == B200PT_SAMPLER_SOBOL) idx = sobol_interval_to_index(S.sobol, (unsigned long long)s, px - S.sobol.sb_min[0], py - S.sobol.sb_min[1]);
    else {  // owned pixel index: position of the row among this shard's rows (explicit lists own every row)
        long long krow = list ? (long long)(py - S.sb[1]) : (g / spp) / (S.sb[2] - S.sb[0]);
        idx = ((unsigned long long)(krow * (S.sb[2] - S.sb[0]) + (px - S.sb[0])) << 16) | (unsigned long long)s;
    }
    int dim = 0;
    P2 fs;  // Sampler::get_camera_sample, sampler/mod.rs:43-51
    if (S.sampler_type == B200PT_SAMPLER_SOBOL) { fs = mk2(sobol_film_dim(S.sobol, idx, 0, px), sobol_film_dim(S.sobol, idx, 1, py)); dim = 2; }
    else fs = smp_2d(S, idx, dim);
    P2 pf = mk2((float)px + fs.x, (float)py + fs.y);
    float tu = smp_1d(S, idx, dim);
    P2 pl = smp_2d(S, idx, dim);
    Ray32 r = camera_ray(S.camera, pf, tu, pl);
    bool live = list || (px >= S.pb[0] && px < S.pb[2] && py >= S.pb[1] && py < S.pb[3]);  // sampler_integrator.rs:348
    W.ray[0][2 * p] = make_float4(r.ox, r.oy, r.oz, live ? r.tmax : -1.0f);  // tmax < 0: the root test fails, the path dies as a miss
    W.ray[0][2 * p + 1] = make_float4(r.dx, r.dy, r.dz, r.time);
    W.qpid[0][p] = p;
    W.L[p] = make_float4(0.0f, 0.0f, 0.0f, live ? 1.0f : 0.0f);
    W.beta[p] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    W.hidx[p] = idx;
    W.meta[p] = meta_pack(dim, live ? 0 : 255, 0);
    if (W.cam_diff) camera_differentials(S.camera, pf, pl, mk(r.ox, r.oy, r.oz), mk(r.dx, r.dy, r.dz), S.cam_diff_scale, W.cam_diff + 3ll * p);
    if (p_film_out) p_film_out[p] = make_float2(pf.x, pf.y);  // wave-local, like the radiance k_store_samples writes
    if (rays_out) { rays_out[2 * p] = make_float4(r.ox, r.oy, r.oz, r.tmax); rays_out[2 * p + 1] = make_float4(r.dx, r.dy, r.dz, r.time); }
}

// ---- K6: sort the ray queue by material so that a shading warp runs one BSDF model ---------------
// Counting sort with 5 bins (miss, matte, plastic, glass, metal): pass 1 classifies every slot and
// histograms per block (one global atomic per bin per block), pass 2 scatters slot ids to their bin with
// one warp-aggregated atomic per bin per warp.  Order inside a bin is arbitrary; results do not depend on it.
// Queue sizes stay on the device (no host round trip per bounce): n_ptr, when set, points at the count a previous kernel
// left in the wave's control blocks; the grid is sized for an upper bound and strides.
B2_D int dev_count(const int* n_ptr, int n_host) { return n_ptr ? *n_ptr : n_host; }

__global__ void __launch_bounds__(256) k_bin_count(DeviceScene S, Wave W, const int* __restrict__ n_ptr, int n_host) {
    __shared__ int hist[kBins];
    const int n_active = dev_count(n_ptr, n_host);
    if (threadIdx.x < kBins) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_active; slot += gridDim.x * blockDim.x) {
        uint32_t prim = __float_as_uint(W.hit[slot].y);
        int key = 0;
        if (prim != 0xffffffffu) {
            int mat = __float_as_int(ldg4(S.prim_verts + 3ll * prim).w);
            key = mat < 0 ? kBinNull : 1 + S.materials[mat].type;
        }
        W.key[slot] = (uint8_t)key;
        atomicAdd(&hist[key], 1);
    }
    __syncthreads();
    if (threadIdx.x < kBins && hist[threadIdx.x]) atomicAdd(&W.counters[8 + threadIdx.x], hist[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_bin_scatter(Wave W, const int* __restrict__ n_ptr, int n_host) {
    __shared__ int wcount[8][kBins];  // per warp: count, then start offset inside the bin
    const int n_active = dev_count(n_ptr, n_host);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int bin_base[kBins];
    { int run = 0; for (int k = 0; k < kBins; ++k) { bin_base[k] = run; run += W.counters[8 + k]; } }
    for (int first = blockIdx.x * blockDim.x; first < n_active; first += gridDim.x * blockDim.x) {  // block-uniform trip count
        const int slot = first + threadIdx.x;
        const int key = slot < n_active ? (int)W.key[slot] : -1;
        unsigned mine = 0;
        for (int k = 0; k < kBins; ++k) {
            unsigned m = __ballot_sync(0xffffffffu, key == k);
            if (key == k) mine = m;
            if (lane == 0) wcount[warp][k] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x < kBins) {  // one global atomic per bin per block, then an exclusive scan over the block's warps
            int total = 0;
            for (int w = 0; w < 8; ++w) total += wcount[w][threadIdx.x];
            int run = total ? atomicAdd(&W.counters[16 + threadIdx.x], total) : 0;
            for (int w = 0; w < 8; ++w) { int c = wcount[w][threadIdx.x]; wcount[w][threadIdx.x] = run; run += c; }
        }
        __syncthreads();
        if (key >= 0) W.sorted[bin_base[key] + wcount[warp][key] + __popc(mine & ((1u << lane) - 1u))] = slot;
        __syncthreads();
    }
}

// ---- SpatialLightDistribution (core/src/light_distrib/spatial.rs) -------------------------------------------------
// compute_distribution(), spatial.rs:91-137: one thread per (queued voxel, light) walks the 128 Halton points in order
__global__ void __launch_bounds__(128) k_voxel_contrib(DeviceScene S, const int* __restrict__ n_work_ptr) {
    const long long total = (long long)*n_work_ptr * S.n_lights;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(t / S.n_lights), j = (int)(t % S.n_lights);
    const int v = S.vox_work[w];
    const int pz = v % S.n_voxels[2], py = (v / S.n_voxels[2]) % S.n_voxels[1], px = v / (S.n_voxels[2] * S.n_voxels[1]);
    const int pi[3] = {px, py, pz};
    float lo[3], hi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t0 = (float)pi[i] / (float)S.n_voxels[i], t1 = (float)(pi[i] + 1) / (float)S.n_voxels[i];
        lo[i] = lerp_ref(t0, S.wb[i], S.wb[3 + i]);
        hi[i] = lerp_ref(t1, S.wb[i], S.wb[3 + i]);
    }
    const DLight& light = S.lights[j];
    float contrib = 0.0f;
    for (int i = 0; i < 128; ++i) {
        const unsigned long long a = (unsigned long long)i;
        SurfHit sh;
        sh.p = mk(lerp_ref(radical_inverse_base2(a), lo[0], hi[0]), lerp_ref(radical_inverse_specialized(3, a), lo[1], hi[1]),
                  lerp_ref(radical_inverse_specialized(5, a), lo[2], hi[2]));
        sh.p_error = mk(0, 0, 0); sh.n = mk(0, 0, 0); sh.ns = sh.n; sh.dpdu = sh.n;
        const P2 u = mk2(radical_inverse_specialized(7, a), radical_inverse_specialized(11, a));
        const LightSample ls = sample_light(S, light, sh, u);
        if (ls.valid && ls.pdf > 0.0f) contrib += lum_y(ls.Li) / ls.pdf;
    }
    S.vox_table[(long long)S.vox_row[v] * (2 * S.n_lights + 2) + j] = contrib;
    }
}

// spatial.rs:139-160 + Distribution1D::new (distribution_1d.rs:22-48): one thread per queued voxel, lights in order
__global__ void __launch_bounds__(128) k_voxel_finish(DeviceScene S, const int* __restrict__ n_work_ptr) {
    const int n_work = *n_work_ptr;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x) {
    const int v = S.vox_work[w], n = S.n_lights;
    float* func = S.vox_table + (long long)S.vox_row[v] * (2 * n + 2);
    float* cdf = func + n;
    float sum = 0.0f;
    for (int j = 0; j < n; ++j) sum += func[j];
    const float avg = sum / (float)(128 * n);
    const float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (int j = 0; j < n; ++j) func[j] = pmax(func[j], min_contrib);
    cdf[0] = 0.0f;
    for (int j = 1; j < n + 1; ++j) cdf[j] = cdf[j - 1] + func[j - 1] / (float)n;
    const float func_int = cdf[n];
    if (func_int == 0.0f) { for (int j = 1; j < n + 1; ++j) cdf[j] = (float)j / (float)n; }
    else { for (int j = 1; j < n + 1; ++j) cdf[j] /= func_int; }
    func[2 * n + 1] = func_int;
    __threadfence();
    S.vox_state[v] = 2;
    }
}

// ---- K4: shade --------------------------------------------------------------------------------
// One path vertex (path.rs:117-280).  KM is the compile-time lobe mask of the material class the launch covers (the
// queue is sorted by class first, so a launch only carries the BSDF code its class can reach); kMiss = the slot holds
// an escaped ray, kNull = the hit primitive has no material (path.rs:146-150: the ray is re-spawned, nothing is counted).
enum { kShadeHit = 0, kShadeMiss = 1, kShadeNull = 2 };
template <uint32_t KM, int kMode>
B2_D void shade_vertex(const DeviceScene& S, const Wave& W, int cur, int slot, bool may_park) {
    const int pid = W.qpid[cur][slot];
    const float4 r1 = W.ray[cur][2 * slot + 1];
    const float4 hit = W.hit[slot];
    const float hb2 = W.hit_b2[slot];
    const V3 ray_d = mk(r1.x, r1.y, r1.z);
    const float time = r1.w;
    int meta = W.meta[pid];
    int dim = meta & 0xffff, bounces = (meta >> 16) & 0xff;
    bool specular_bounce = ((meta >> 24) & 1) != 0;
    if (bounces == 255) return;  // pixel outside the integrator's pixel bounds: no sample is taken
    float4 Lw = W.L[pid], bw = W.beta[pid];
    RGB L = rgb(Lw.x, Lw.y, Lw.z), beta = rgb(bw.x, bw.y, bw.z);
    float eta_scale = bw.w;
    const unsigned long long hidx = W.hidx[pid];
    const uint32_t prim = __float_as_uint(hit.y);
    const bool found = kMode == kShadeMiss ? false : (KM == KM_ALL ? prim != 0xffffffffu : true);

    if (kMode == kShadeMiss || !found) {  // path.rs:123-137: light from the environment, then the path ends
        if (bounces == 0 || specular_bounce) {
            for (int i = 0; i < S.n_infinite; ++i) {
                const DLight& il = S.lights[S.infinite_lights[i]];
                L = L + beta * infinite_le(il, S.inf_distr[il.inf_slot], ray_d);
            }
            W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
        }
        return;
    }
    HitCtx hc;
    surface_at(S, W, slot, prim, hit, hb2, ray_d, true, &hc);
    const SurfHit& sh = hc.sh;
    const V3 hit_wo = hc.wo;
    const int mat = hc.mat, alight = hc.alight;
    // path.rs:123-134: emitted light at the vertex
    if ((bounces == 0 || specular_bounce) && alight >= 0) L = L + beta * area_l(S.lights[alight], sh.n, -ray_d);
    if (bounces >= S.max_depth) {  // path.rs:137
        W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
        return;
    }
    if (kMode == kShadeNull || (KM == KM_ALL && mat < 0)) {
        // path.rs:146-150: no BSDF (Material "" / "none"): isect.spawn_ray(ray.d), `continue` without counting a bounce
        W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
        W.meta[pid] = meta | (1 << 25);  // the re-spawned ray carries no differentials (Hit::spawn_ray)
        const int ns = atomicAdd(&W.counters[0], 1);
        store_ray(W.ray[cur ^ 1], ns, offset_ray_origin(sh.p, sh.p_error, sh.n, ray_d), ray_d, __int_as_float(0x7f800000), time);
        W.qpid[cur ^ 1][ns] = pid;
        return;
    }
    // BSDF::new frame (bsdf.rs:100-120)
    BSDF bsdf;
    bsdf.ns = sh.ns; bsdf.ng = sh.n;
    bsdf.ss = normalize(sh.dpdu);
    bsdf.ts = cross(bsdf.ns, bsdf.ss);
    bsdf.m = S.materials + mat;
    DMaterial tex_mat;  // KM_TEX kernels only: the material with this intersection's Kd
    if ((KM & KM_TEX) && S.mat_kd_tex) {
        const int tex = S.mat_kd_tex[mat];
        if (tex >= 0) {
            const bool camera_ray = bounces == 0 && !((meta >> 25) & 1);  // path.rs: every later ray comes from spawn_ray
            textured_material(S, W, slot, prim, hit, hb2, sh, tex, (camera_ray && W.cam_diff) ? W.cam_diff + 3ll * pid : nullptr, S.materials[mat], &tex_mat);
            bsdf.m = &tex_mat;
        }
    }
    const uint32_t kNoSpec = BSDF_ALL & ~BSDF_SPECULAR;

    // path.rs:162-173 -> uniform_sample_one_light (integrator/common.rs:89-133)
    if (bsdf_num_components(bsdf, kNoSpec) > 0 && S.n_lights > 0) {
        // path.rs:156-157: light_distribution.lookup(&isect.hit.p)
        const float* lfunc = S.light_func;
        const float* lcdf = S.light_cdf;
        float lfunc_int = S.light_func_int;
        if (S.spatial) {
            const int v = spatial_voxel(S, sh.p);
            if (ld_volatile_int(S.vox_state + v) != 2) {
                // First touch of this voxel: the thread that wins the CAS takes a table row and queues the voxel; the slot is
                // parked and shaded again, by a second launch of this kernel, once the voxel kernels have run (nothing of
                // this path has been written yet).  A full row pool is reported to the host at the end of the wave.
                if (atomicCAS(S.vox_state + v, 0, 1) == 0) {
                    const int row = atomicAdd(S.vox_pool_next, 1);
                    if (row < S.vox_pool_cap) { S.vox_row[v] = row; S.vox_work[atomicAdd(&W.counters[24], 1)] = v; }
                    else atomicExch(&W.counters[26], 1);
                }
                if (may_park) W.deferred[atomicAdd(&W.counters[25], 1)] = slot;  // a parked slot whose voxel is still missing in the second pass (row pool exhausted) is dropped
                return;
            }
            const float* row = S.vox_table + (long long)S.vox_row[v] * (2 * S.n_lights + 2);
            lfunc = row; lcdf = row + S.n_lights; lfunc_int = row[2 * S.n_lights + 1];
        }
        float u_pick = smp_1d(S, hidx, dim);
        // Distribution1D::sample_discrete, distribution_1d.rs:81-94
        int ln = find_interval_cdf(lcdf, S.n_lights + 1, u_pick);
        float pick_pdf = lfunc_int > 0.0f ? lfunc[ln] / (lfunc_int * (float)S.n_lights) : 0.0f;
        if (pick_pdf != 0.0f) {
            P2 u_light = smp_2d(S, hidx, dim);
            P2 u_scatter = smp_2d(S, hidx, dim);
            const DLight& light = S.lights[ln];
            // ---- estimate_direct (common.rs:146-299): both halves, rays still to be traced ----
            const DirectEst de = estimate_direct_rays<KM>(S, light, sh, hit_wo, bsdf, u_light, u_scatter);
            const RGB ld_light = de.ld_light, mis_f = de.mis_f;
            const float mis_w = de.mis_w, mis_pdf = de.mis_pdf;
            int shadow_slot = -1, mis_slot = -1;
            if (de.shadow) {
                shadow_slot = atomicAdd(&W.counters[1], 1);
                store_ray(W.sh_ray, shadow_slot, de.sh_o, de.sh_d, 1.0f - kShadowEps, time);
            }
            if (de.mis) {
                mis_slot = atomicAdd(&W.counters[2], 1);
                store_ray(W.mis_ray, mis_slot, de.mis_o, de.mis_d, __int_as_float(0x7f800000), time);
            }
            if (shadow_slot >= 0 || mis_slot >= 0) {
                int k = atomicAdd(&W.counters[3], 1);
                W.pend_q[k] = pid;
                W.pend_a[pid] = make_float4(ld_light.r, ld_light.g, ld_light.b, pick_pdf);
                W.pend_b[pid] = make_float4(mis_f.r, mis_f.g, mis_f.b, mis_w);
                W.pend_c[pid] = make_float4(beta.r, beta.g, beta.b, mis_pdf);
                W.pend_d[pid] = make_int4(ln, shadow_slot, mis_slot, 0);
            }
        }
    }

    // path.rs:175-206: sample the BSDF for the next direction
    P2 u = smp_2d(S, hidx, dim);
    V3 wo = -ray_d;
    BxDFSample bs = bsdf_sample_f<KM>(bsdf, wo, u, BSDF_ALL);
    if (is_black(bs.f) || bs.pdf == 0.0f) {
        W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
        return;
    }
    beta = beta * (bs.f * abs_dot(bs.wi, sh.ns) / bs.pdf);
    specular_bounce = (bs.type & BSDF_SPECULAR) != 0;
    if ((bs.type & BSDF_SPECULAR) && (bs.type & BSDF_TRANSMISSION)) {
        float eta = 1.0f;  // BSDF::new(.., None): every in-scope material leaves bsdf.eta at 1.0
        eta_scale *= dot(wo, sh.n) > 0.0f ? eta * eta : 1.0f / (eta * eta);
    }
    V3 next_o = offset_ray_origin(sh.p, sh.p_error, sh.n, bs.wi);
    // path.rs:264-277: Russian roulette
    RGB rr_beta = beta * eta_scale;
    bool alive = true;
    if (max_component_value(rr_beta) < S.rr_threshold && bounces > 3) {
        float q = pmax(0.05f, 1.0f - max_component_value(rr_beta));
        float ur = smp_1d(S, hidx, dim);
        if (ur < q) alive = false;
        else beta = beta / (1.0f - q);
    }
    W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
    if (!alive) return;
    bounces += 1;
    W.beta[pid] = make_float4(beta.r, beta.g, beta.b, eta_scale);
    W.meta[pid] = meta_pack(dim, bounces, specular_bounce ? 1 : 0);
    int ns = atomicAdd(&W.counters[0], 1);
    store_ray(W.ray[cur ^ 1], ns, next_o, bs.wi, __int_as_float(0x7f800000), time);
    W.qpid[cur ^ 1][ns] = pid;
}

// The slots of the sorted queue that belong to bins [bin_lo, bin_hi) (bin_lo >= 0), or the slots the first pass parked
// for the spatial light table (bin_lo < 0).  Counts are read from the wave's control block; the grid strides.
template <uint32_t KM, int kMode, int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) k_shade(DeviceScene S, Wave W, int cur, int bin_lo, int bin_hi) {
    const int* order;
    int count = 0;
    if (bin_lo >= 0) {
        int start = 0;
        for (int k = 0; k < bin_lo; ++k) start += W.counters[8 + k];
        for (int k = bin_lo; k < bin_hi; ++k) count += W.counters[8 + k];
        order = W.sorted + start;
    } else {
        order = W.deferred;
        count = W.counters[25];
    }
    const bool may_park = bin_lo >= 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) shade_vertex<KM, kMode>(S, W, cur, order[i], may_park);
}

// All material classes of a bounce in ONE launch: virtual blocks of 128 slots are dealt over the classes' ranges of the
// sorted queue (bins 1..4), every block runs one class's code (the switch is block-uniform).  Same registers as the
// per-class kernels (the largest class sets them); what it buys is latency: shading ONE vertex is a ~50 us dependent
// chain, and four nearly empty class launches in a row cost four of those per bounce (profiles/r2_launches_floor_c3_tiny.csv:
// 183 us of a 256 us bounce floor), which is what limits small renders and the per-GPU share of a multi-GPU render.
static const uint32_t kKmMatteD = (1u << BX_LAMBERT) | (1u << BX_OREN_NAYAR);
static const uint32_t kKmPlasticD = (1u << BX_LAMBERT) | (1u << BX_MF_REFL) | KM_DIEL;
static const uint32_t kKmGlassD = (1u << BX_FRESNEL_SPECULAR) | (1u << BX_MF_REFL) | (1u << BX_MF_TRANS) | KM_DIEL;
static const uint32_t kKmMetalD = (1u << BX_MF_REFL) | KM_COND;
template <int kMinBlocks>
__global__ void __launch_bounds__(128, kMinBlocks) k_shade_classes(DeviceScene S, Wave W, int cur) {
    int start[5], vb[5];
    int acc = W.counters[8], v_total = 0;  // bin 0 = escaped rays (their own small kernel)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = W.counters[9 + k];
        start[k] = acc; acc += c;
        vb[k] = v_total; v_total += (c + 127) >> 7;
    }
    start[4] = acc; vb[4] = v_total;
    for (int v = blockIdx.x; v < v_total; v += gridDim.x) {
        int b = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) b += (v >= vb[k]) ? 1 : 0;
        const int i = ((v - vb[b]) << 7) + (int)threadIdx.x;
        if (i >= start[b + 1] - start[b]) continue;
        const int slot = W.sorted[start[b] + i];
        switch (b) {  // block-uniform
            case 0: shade_vertex<kKmMatteD, kShadeHit>(S, W, cur, slot, true); break;
            case 1: shade_vertex<kKmPlasticD, kShadeHit>(S, W, cur, slot, true); break;
            case 2: shade_vertex<kKmGlassD, kShadeHit>(S, W, cur, slot, true); break;
            default: shade_vertex<kKmMetalD, kShadeHit>(S, W, cur, slot, true); break;
        }
    }
}

// ---- (0,2)-sequence prepass: one thread per reference tile replays the tile sampler's PCG32 stream ------------
// ZeroTwoSequenceSampler::start_pixel for every pixel of the tile in row-major order (zero_two_sequence.rs:65-111,
// sampler_integrator.rs:323-345): per slot one (two) scramble draw(s), spp one-element shuffles (one draw each, the
// rejection threshold of bounded_uniform_u32(0, 1) is 0) and one Fisher-Yates shuffle of the spp samples.
__global__ void __launch_bounds__(64) k_zerotwo_tiles(int sb0, int sb1, int sb2, int sb3, int ntx, int nty, int dims, int n1, int n2, int spp,
                                                      const int* __restrict__ row_index, uint32_t* __restrict__ scr1, uint16_t* __restrict__ perm1,
                                                      uint32_t* __restrict__ scr2, uint16_t* __restrict__ perm2, uint16_t* __restrict__ scratch) {
    int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= ntx * nty) return;
    const int tx = tile % ntx, ty = tile / ntx, sw = sb2 - sb0;
    DPcg32 rng;
    pcg_set_sequence(rng, (unsigned long long)tile);  // clone_sampler(tile_idx) -> RNG::new(seed)
    uint16_t* idx = scratch + (size_t)tile * spp;
    const int x0 = sb0 + tx * 16, x1 = min(x0 + 16, sb2), y0 = sb1 + ty * 16, y1 = min(y0 + 16, sb3);
    for (int y = y0; y < y1; ++y) {
        const int krow = row_index[y - sb1];
        for (int x = x0; x < x1; ++x) {
            const long long pix = (long long)krow * sw + (x - sb0);
            for (int pass = 0; pass < 2; ++pass) {       // pass 0: van_der_corput slots, pass 1: sobol_2d slots
                const int keep = pass == 0 ? n1 : n2;
                for (int d = 0; d < dims; ++d) {
                    uint32_t s0 = pcg_next(rng), s1 = pass ? pcg_next(rng) : 0u;
                    for (int i = 0; i < spp; ++i) (void)pcg_next(rng);
                    const bool store = krow >= 0 && d < keep;
                    if (store) for (int i = 0; i < spp; ++i) idx[i] = (uint16_t)i;
                    for (int i = 0; i < spp; ++i) {
                        int other = i + (int)pcg_bounded(rng, (uint32_t)(spp - i));
                        if (store) { uint16_t t = idx[i]; idx[i] = idx[other]; idx[other] = t; }
                    }
                    if (store) {
                        if (pass == 0) {
                            scr1[pix * n1 + d] = s0;
                            uint16_t* o = perm1 + (pix * n1 + d) * spp;
                            for (int i = 0; i < spp; ++i) o[i] = idx[i];
                        } else {
                            scr2[(pix * n2 + d) * 2] = s0; scr2[(pix * n2 + d) * 2 + 1] = s1;
                            uint16_t* o = perm2 + (pix * n2 + d) * spp;
                            for (int i = 0; i < spp; ++i) o[i] = idx[i];
                        }
                    }
                }
            }
        }
    }
}

// ---- (0,2)-sequence, tile-sequential mode -----------------------------------------------------------------------
// With fewer pre-generated "dimensions" than a path can ask for (the reference's default is 4,
// samplers/src/zero_two_sequence.rs:157) PixelSampler::get_1d / get_2d fall back to the TILE sampler's RNG inside li()
// (pixel_sampler.rs:88-110), so the stream position of every later sample of the tile depends on the lengths of the
// paths before it: a tile is a sequential stream.  Tiles are independent (clone_sampler(tile_idx),
// sampler_integrator.rs:323), so the wave holds ONE path per tile: every step each tile moves on to its next
// (pixel, sample) in the reference's order (start_pixel draws the pixel's tables from the tile RNG, also for pixels
// outside the integrator's pixel bounds, sampler_integrator.rs:343-350), the wave runs that path to its end, and its
// radiance goes to the (pixel, sample) place of the sample store.  256 x spp waves of (number of tiles) paths: a slow
// path by construction, bit-compatible with the reference's streams.
struct ZtSeq {
    DPcg32* rng;     // [slot] the tile's PCG32 stream (also read by smp_1d / smp_2d through DeviceScene::zt.rng)
    int* tile;       // [slot] reference tile index
    int* cursor;     // [slot] pixel position inside the tile, row-major
    int* samp;       // [slot] sample number of the path in flight
    int* dst;        // [slot] place of the path's sample in the sample store, -1 = nowhere (rows of another shard)
    uint32_t* scr1;  // tables of the slot's CURRENT pixel: [slot][dims], [slot][dims][2], [slot][dims][spp] x 2
    uint32_t* scr2;
    uint16_t* perm1;
    uint16_t* perm2;
    int dims, spp, ntx;
};
// ZeroTwoSequenceSampler::start_pixel (zero_two_sequence.rs:65-111): per 1-D slot one scramble, per 2-D slot two, then
// spp one-element shuffles (one draw each: bounded_uniform_u32(0, 1) has threshold 0) and one shuffle of the spp samples.
B2_D void zt_start_pixel(DPcg32& rng, int dims, int spp, bool store, uint32_t* scr1, uint16_t* perm1, uint32_t* scr2, uint16_t* perm2) {
    for (int pass = 0; pass < 2; ++pass)
        for (int d = 0; d < dims; ++d) {
            const uint32_t s0 = pcg_next(rng), s1 = pass ? pcg_next(rng) : 0u;
            for (int i = 0; i < spp; ++i) (void)pcg_next(rng);
            uint16_t* o = (pass ? perm2 : perm1) + (size_t)d * spp;
            if (store) for (int i = 0; i < spp; ++i) o[i] = (uint16_t)i;
            for (int i = 0; i < spp; ++i) {
                const int other = i + (int)pcg_bounded(rng, (uint32_t)(spp - i));
                if (store) { const uint16_t t = o[i]; o[i] = o[other]; o[other] = t; }
            }
            if (store) {
                if (pass == 0) scr1[d] = s0;
                else { scr2[2 * d] = s0; scr2[2 * d + 1] = s1; }
            }
        }
}
// Camera sample + ray of sample `smp` of pixel (px, py) for the path in slot p (Sampler::get_camera_sample, k_raygen).
B2_D Ray32 zt_seq_emit(const DeviceScene& S, const Wave& W, int p, int px, int py, int smp, bool live, P2* pf_out) {
    const unsigned long long key = ((unsigned long long)p << 16) | (unsigned long long)smp;
    int dim = 0;
    const P2 fs = smp_2d(S, key, dim);
    const P2 pf = mk2((float)px + fs.x, (float)py + fs.y);
    const float tu = smp_1d(S, key, dim);
    const P2 pl = smp_2d(S, key, dim);
    const Ray32 r = camera_ray(S.camera, pf, tu, pl);
    W.ray[0][2 * p] = make_float4(r.ox, r.oy, r.oz, live ? r.tmax : -1.0f);
    W.ray[0][2 * p + 1] = make_float4(r.dx, r.dy, r.dz, r.time);
    W.qpid[0][p] = p;
    W.L[p] = make_float4(0.0f, 0.0f, 0.0f, live ? 1.0f : 0.0f);
    W.beta[p] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    W.hidx[p] = key;
    W.meta[p] = meta_pack(dim, live ? 0 : 255, 0);
    if (W.cam_diff) camera_differentials(S.camera, pf, pl, mk(r.ox, r.oy, r.oz), mk(r.dx, r.dy, r.dz), S.cam_diff_scale, W.cam_diff + 3ll * p);
    *pf_out = pf;
    return r;
}
B2_D void zt_seq_dead(const Wave& W, int p) {
    W.ray[0][2 * p] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
    W.ray[0][2 * p + 1] = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
    W.qpid[0][p] = p;
    W.L[p] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    W.beta[p] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    W.hidx[p] = 0ull;
    W.meta[p] = meta_pack(0, 255, 0);
}
// One step of every tile of the group: the next sample of the current pixel, or start_pixel on the next pixel(s).
__global__ void __launch_bounds__(64) k_zt_seq_next(DeviceScene S, Wave W, ZtSeq Z, int n_slots, int first_step, const int* __restrict__ row_index,
                                                    long long first_local, float2* __restrict__ p_film_out, unsigned long long* __restrict__ totals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots) return;
    const int tile = Z.tile[p], tx = tile % Z.ntx, ty = tile / Z.ntx;
    const int x0 = S.sb[0] + tx * 16, x1 = min(x0 + 16, S.sb[2]), y0 = S.sb[1] + ty * 16, y1 = min(y0 + 16, S.sb[3]);
    const int w = x1 - x0, npix = w * (y1 - y0), spp = Z.spp;
    DPcg32 rng;
    int cur, smp;
    if (first_step) { pcg_set_sequence(rng, (unsigned long long)tile); cur = -1; smp = spp - 1; }  // clone_sampler(tile_idx) -> RNG::new(seed)
    else { rng = Z.rng[p]; cur = Z.cursor[p]; smp = Z.samp[p]; }
    if (cur < npix) {
        if (cur >= 0 && smp + 1 < spp) smp += 1;  // start_next_sample
        else {
            for (++cur; cur < npix; ++cur) {
                zt_start_pixel(rng, Z.dims, spp, true, Z.scr1 + (size_t)p * Z.dims, Z.perm1 + (size_t)p * Z.dims * spp, Z.scr2 + (size_t)p * Z.dims * 2,
                               Z.perm2 + (size_t)p * Z.dims * spp);
                const int px = x0 + cur % w, py = y0 + cur / w;
                if (px >= S.pb[0] && px < S.pb[2] && py >= S.pb[1] && py < S.pb[3]) break;  // sampler_integrator.rs:348: checked after start_pixel
            }
            smp = 0;
        }
    }
    Z.rng[p] = rng; Z.cursor[p] = cur; Z.samp[p] = smp;
    if (cur >= npix) {
        // the tile has no sample left: its slot idles through the wave; take it out of the camera / closest-hit ray totals
        // the wave bookkeeping adds for every slot (k_wave_end)
        zt_seq_dead(W, p); Z.dst[p] = -1;
        atomicAdd(totals, ~0ull); atomicAdd(totals + 1, ~0ull);
        return;
    }
    const int px = x0 + cur % w, py = y0 + cur / w;
    P2 pf;
    zt_seq_emit(S, W, p, px, py, smp, true, &pf);
    const int krow = row_index[py - S.sb[1]];
    int dst = -1;
    if (krow >= 0) {
        dst = (int)((((long long)krow * (S.sb[2] - S.sb[0]) + (px - S.sb[0])) * spp + smp) - first_local);
        p_film_out[dst] = make_float2(pf.x, pf.y);
    }
    Z.dst[p] = dst;
}
// Explicit (x, y, sample) triples (b200pt_li_batch): every entry gets a fresh copy of its tile's stream positioned as
// render_tile would reach the pixel if no earlier path had drawn from it - start_pixel on every earlier pixel of the
// tile, then on the pixel itself (the oracle's sampler_at convention; a render's true stream position also depends on
// the earlier paths, which only the render itself reproduces).
__global__ void __launch_bounds__(64) k_zt_seq_list(DeviceScene S, Wave W, ZtSeq Z, int n, const int* __restrict__ list, float4* __restrict__ rays_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int px = list[3 * p], py = list[3 * p + 1], smp = list[3 * p + 2];
    const int tx = (px - S.sb[0]) / 16, ty = (py - S.sb[1]) / 16;
    const int x0 = S.sb[0] + tx * 16, x1 = min(x0 + 16, S.sb[2]), y0 = S.sb[1] + ty * 16;
    DPcg32 rng;
    pcg_set_sequence(rng, (unsigned long long)(ty * Z.ntx + tx));
    const int w = x1 - x0, own = (py - y0) * w + (px - x0);
    for (int cur = 0; cur <= own; ++cur)
        zt_start_pixel(rng, Z.dims, Z.spp, cur == own, Z.scr1 + (size_t)p * Z.dims, Z.perm1 + (size_t)p * Z.dims * Z.spp, Z.scr2 + (size_t)p * Z.dims * 2,
                       Z.perm2 + (size_t)p * Z.dims * Z.spp);
    Z.rng[p] = rng;
    P2 pf;
    const Ray32 r = zt_seq_emit(S, W, p, px, py, smp, true, &pf);
    if (rays_out) { rays_out[2 * p] = make_float4(r.ox, r.oy, r.oz, r.tmax); rays_out[2 * p + 1] = make_float4(r.dx, r.dy, r.dz, r.time); }
}
// The finished paths' radiances go to their (pixel, sample) places (sanitised like k_store_samples).
__global__ void __launch_bounds__(256) k_zt_seq_store(Wave W, const int* __restrict__ dst, int n, float4* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int d = dst[p];
    if (d < 0) return;
    const float4 l = W.L[p];
    RGB c = rgb(l.x, l.y, l.z);
    const float y = lum_y(c);
    if (isnan(c.r) || isnan(c.g) || isnan(c.b) || y < -1e-5f || isinf(y)) c = rgb1(0.0f);
    out[d] = make_float4(c.r, c.g, c.b, l.w);
}

// ---- K4': resolve pending direct lighting ------------------------------------------------------
__global__ void __launch_bounds__(256) k_resolve(DeviceScene S, Wave W) {
    const int n_pend = W.counters[3];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pend; i += gridDim.x * blockDim.x) {
    const int pid = W.pend_q[i];
    float4 a = W.pend_a[pid], b = W.pend_b[pid], c = W.pend_c[pid];
    int4 d = W.pend_d[pid];
    RGB ld = rgb1(0.0f);
    if (d.y >= 0 && !W.sh_occ[d.y]) ld = ld + rgb(a.x, a.y, a.z);  // common.rs:205-225
    if (d.z >= 0) {  // common.rs:266-296
        const DLight& light = S.lights[d.x];
        float4 mh = W.mis_hit[d.z];
        float4 md = W.mis_ray[2 * d.z + 1];
        V3 wi = mk(md.x, md.y, md.z);
        uint32_t prim = __float_as_uint(mh.y);
        RGB Li = rgb1(0.0f);
        if (prim != 0xffffffffu) {
            V3 p0, p1, p2; int mat, al; uint32_t fl;
            load_prim(S, prim, &p0, &p1, &p2, &mat, &al, &fl);
            if (al == d.x) Li = area_l(light, emitter_hit_normal(S, prim, p0, p1, p2, fl, mh.z, mh.w, W.mis_b2[d.z]), -wi);
        } else if (light.type == LT_INFINITE) {
            Li = infinite_le(light, S.inf_distr[light.inf_slot], wi);
        }
        if (!is_black(Li)) ld = ld + rgb(b.x, b.y, b.z) * Li * rgb1(1.0f) * b.w / c.w;
    }
    RGB add = rgb(c.x, c.y, c.z) * (ld / a.w);  // path.rs:165: beta * (estimate / light_pdf)
    float4 Lw = W.L[pid];
    W.L[pid] = make_float4(Lw.x + add.r, Lw.y + add.g, Lw.z + add.b, Lw.w);
    }
}

// Wave bookkeeping kept on the device: the control blocks are cleared and the first queue size set when a wave starts;
// when it ends the rays it traced are added to the scene's totals (read once, at the end of the render).
__global__ void k_wave_begin(int* ctl, int n_ints, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ints; i += gridDim.x * blockDim.x) ctl[i] = i == 0 ? n : 0;
}
__global__ void k_wave_end(const int* ctl, int n_blocks, int n_camera, unsigned long long* totals) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long closest = 0, shadow = 0, flags = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const int* c = ctl + b * kCtl;
        // the queue this block's counts opened (block 0: the camera rays); the last block's queue is traced by the next
        // segment, as its block 0, or not at all
        if (b + 1 < n_blocks) closest += (unsigned long long)c[0];
        if (b > 0) { closest += (unsigned long long)c[2]; shadow += (unsigned long long)c[1]; }
        if (c[26]) flags |= 1ull;
    }
    if (ctl[(n_blocks - 1) * kCtl] != 0) flags |= 2ull;      // paths still alive after the last iteration that was launched
    totals[0] += (unsigned long long)n_camera; totals[1] += closest; totals[2] += shadow; totals[3] |= flags;
}

// Copies the wave's final radiances into the per-sample store (sanitised, sampler_integrator.rs:374-401).
__global__ void __launch_bounds__(256) k_store_samples(Wave W, int n, float4* __restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 l = W.L[p];
    RGB c = rgb(l.x, l.y, l.z);
    float y = lum_y(c);
    if (isnan(c.r) || isnan(c.g) || isnan(c.b) || y < -1e-5f || isinf(y)) c = rgb1(0.0f);
    out[p] = make_float4(c.r, c.g, c.b, l.w);
}

// ---- K5: film --------------------------------------------------------------------------------
struct DFilm {
    int crop[4];
    float rx, ry, inv_rx, inv_ry;
    float max_lum;
    int sb[4];
    int tile;  // reference tile size (16): a sample only reaches pixels of its own tile's pixel bounds
};
// One thread per film pixel of rows [y_begin, y_end): continues the pixel's running sums with every sample of the
// wave [first, first + n) (pixel-major sample indices of this shard) whose filter window covers the pixel, in
// pixel-major then sample order (film_tile.rs:62-108).  A pixel's samples reach it in ascending sample index whatever
// the wave boundaries are, so the sums - kept in `acc` between waves - do not depend on how the render was cut into
// waves (tests: wave / pass splitting leaves every bit of the image unchanged).
__global__ void __launch_bounds__(128) k_film(DFilm F, const float* __restrict__ table, const float4* __restrict__ sample_L,
                                              const float2* __restrict__ p_film, long long first, int n, int spp, const int* __restrict__ row_index,
                                              int y_begin, int y_end, float4* __restrict__ acc) {
    int w = F.crop[2] - F.crop[0];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * (y_end - y_begin)) return;
    int x = F.crop[0] + i % w, y = y_begin + i / w;
    int sw = F.sb[2] - F.sb[0];
    // candidate source pixels: those whose sample positions can reach (x, y)
    int qx0 = (int)floorf((float)x + 0.5f - F.rx) - 1, qx1 = (int)ceilf((float)x + 0.5f + F.rx) + 1;
    int qy0 = (int)floorf((float)y + 0.5f - F.ry) - 1, qy1 = (int)ceilf((float)y + 0.5f + F.ry) + 1;
    qx0 = max(qx0, F.sb[0]); qx1 = min(qx1, F.sb[2]);
    qy0 = max(qy0, F.sb[1]); qy1 = min(qy1, F.sb[3]);
    const long long o = (long long)(y - F.crop[1]) * w + (x - F.crop[0]);
    RGB sum = rgb1(0.0f);
    float wsum = 0.0f;
    bool loaded = false;
    for (int qy = qy0; qy < qy1; ++qy) {
        const int krow = row_index[qy - F.sb[1]];  // position of sample row qy among this shard's rows, -1 = not ours
        if (krow < 0) continue;
        for (int qx = qx0; qx < qx1; ++qx) {
            // pixel bounds of the reference tile that owns sample pixel (qx, qy) (film/mod.rs:182-198)
            int tx0 = F.sb[0] + ((qx - F.sb[0]) / F.tile) * F.tile, ty0 = F.sb[1] + ((qy - F.sb[1]) / F.tile) * F.tile;
            int tx1 = min(tx0 + F.tile, F.sb[2]), ty1 = min(ty0 + F.tile, F.sb[3]);
            int bx0 = max((int)ceilf((float)tx0 - 0.5f - F.rx), F.crop[0]), by0 = max((int)ceilf((float)ty0 - 0.5f - F.ry), F.crop[1]);
            int bx1 = min((int)floorf((float)tx1 - 0.5f + F.rx) + 1, F.crop[2]), by1 = min((int)floorf((float)ty1 - 0.5f + F.ry) + 1, F.crop[3]);
            if (x < bx0 || x >= bx1 || y < by0 || y >= by1) continue;
            const long long base = ((long long)krow * sw + (qx - F.sb[0])) * spp;
            // the part of this source pixel's samples that lies in the wave
            const int s_lo = (int)max(0ll, first - base), s_hi = (int)min((long long)spp, first + n - base);
            for (int s = s_lo; s < s_hi; ++s) {
                // the 8-byte position first: with the box filter 8 of the 9 candidate pixels fail the window test, and
                // their 16-byte radiance is then never fetched
                const long long k = base + s - first;
                float2 pf = p_film[k];
                float dx = pf.x - 0.5f, dy = pf.y - 0.5f;
                int p0x = (int)ceilf(dx - F.rx), p0y = (int)ceilf(dy - F.ry);
                int p1x = (int)floorf(dx + F.rx) + 1, p1y = (int)floorf(dy + F.ry) + 1;
                if (x < p0x || x >= p1x || y < p0y || y >= p1y) continue;
                float4 l = sample_L[k];
                if (l.w == 0.0f) continue;  // pixel outside the integrator's pixel bounds
                if (!loaded) { const float4 a = acc[o]; sum = rgb(a.x, a.y, a.z); wsum = a.w; loaded = true; }
                RGB c = rgb(l.x, l.y, l.z);
                float ly = lum_y(c);
                if (ly > F.max_lum) c = c * F.max_lum / ly;
                float fx = pabs(((float)x - dx) * F.inv_rx * 16.0f), fy = pabs(((float)y - dy) * F.inv_ry * 16.0f);
                int ix = (int)pmin(floorf(fx), 15.0f), iy = (int)pmin(floorf(fy), 15.0f);
                float fw = table[iy * 16 + ix];
                sum = sum + c * 1.0f * fw;  // contrib_sum += l * sample_weight * filter_weight
                wsum += fw;
            }
        }
    }
    if (loaded) acc[o] = make_float4(sum.r, sum.g, sum.b, wsum);
}

// Film::merge_film_tile: RGB sums -> XYZ (film/mod.rs:243-248), written to the shard's film {X, Y, Z, weight sum}.
__global__ void __launch_bounds__(256) k_film_finish(const float4* __restrict__ acc, long long n_pix, float4* __restrict__ film) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_pix) return;
    const float4 a = acc[o];
    float X = 0.412453f * a.x + 0.357580f * a.y + 0.180423f * a.z;
    float Y = 0.212671f * a.x + 0.715160f * a.y + 0.072169f * a.z;
    float Z = 0.019334f * a.x + 0.119193f * a.y + 0.950227f * a.z;
    film[o] = make_float4(X, Y, Z, a.w);
}

// ------------------------------------------------------------------------------------------------
struct SceneImpl {
    int device = -1;          // CUDA device the scene lives on; every entry point switches to it
    AccelImpl accel;
    Accel2Impl accel2;        // two-level scenes (instancing)
    AlphaImpl alpha;          // alpha-mask textures
    bool instanced = false;
    bool whitted = false;     // a recursive SamplerIntegrator (Whitted / DirectLighting) instead of PathIntegrator
    int tree_mode = 0;        // kTreeWhitted / kTreeDirectAll / kTreeDirectOne
    bool has_null_material = false;  // some primitive has no material: paths pass through it without counting a bounce
    bool has_kd_tex = false;         // some matte / plastic material has a textured Kd: their shade kernels are the KM_TEX instantiations
    bool needs_cam_diff = false;     // ... and one of the textures filters over the camera ray's differentials (closedform checkerboard)
    uint32_t material_classes = 0;   // bit t = some material of type t exists (which shade kernels a bounce launches)
    int n_point_lights = 0;          // delta lights: estimate_direct traces no BSDF-sampled ray for them
    DeviceScene dev;
    b200pt_film film;
    b200pt_sampler sampler;
    std::vector<void*> allocs;
    float* d_filter_table = nullptr;
    Wave wave;
    int wave_cap = 0;
    std::vector<void*> wave_ptrs;
    int* d_ctl = nullptr;                  // (kSegIters + 1) control blocks of the current wave segment
    long long wave_launches = 0;           // kernel launches one run_wave call enqueues (for the launch counter of graph replays)
    unsigned long long* d_totals = nullptr;  // camera / closest-hit / shadow rays of the render, error flags
    int* h_pinned = nullptr;               // pinned host words for the few read-backs that remain
    size_t mem_budget = 0;                 // bytes the wave state may take (0: B200PT_MEM_BUDGET, else a share of the free memory)
    // streams / events that let a bounce's shadow, MIS and next closest-hit traversals overlap (run_wave)
    cudaStream_t shade_main = nullptr;  // the stream of the current launch_shade call (the classes fork from / join to it)
    // CUDA graphs of the bounce loop, one per wave size (run_wave_graphed): the ~105 launches of a wave are replayed by the
    // device without the host in between
    struct WaveGraph { int n = 0; int uses = 0; cudaGraphExec_t exec = nullptr; };
    std::vector<WaveGraph> graphs;
    cudaStream_t graph_stream = nullptr;
    cudaEvent_t ev_graph_in = nullptr, ev_graph_out = nullptr;
    bool graph_failed = false;
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_shade = nullptr, ev_fork = nullptr;
    bool aux_ready = false, overlap = true;
    uint64_t rays[3] = {0, 0, 0};
    uint64_t voxels_built = 0;  // SpatialLightDistribution voxels computed so far
    std::mutex mu;
    int sample_bounds[4];
    // the wave's samples (radiance + film position), the running film sums, and film staging; grown on demand, kept across renders
    float4* d_sample_L = nullptr;
    float2* d_sample_pf = nullptr;
    long long sample_cap = 0;
    float4* d_acc = nullptr;
    size_t acc_cap = 0;
    float4* d_film = nullptr;
    size_t film_cap = 0;
    // (0,2)-sequence tables of the current shard
    uint32_t *d_zt_scr1 = nullptr, *d_zt_scr2 = nullptr;
    uint16_t *d_zt_perm1 = nullptr, *d_zt_perm2 = nullptr, *d_zt_scratch = nullptr;
    long long zt_pix_cap = 0;
    // tile-sequential (0,2) mode ("dimensions" below the path's worst case): one path per tile, see ZtSeq
    bool zt_seq = false;
    ZtSeq ztq{};
    int ztq_cap = 0;
    int spp = 1;                 // samples per pixel actually taken (rounded up to a power of two for the (0,2) sampler)
    int* d_rows = nullptr;       // sample rows owned by the current shard
    int* d_row_index = nullptr;  // sample row -> position in d_rows, -1 = not owned
};

// Bytes of wave state per path (what wave_alloc takes for `cap` paths, plus the wave's sample store).
static size_t wave_bytes_per_path(const SceneImpl* s) {
    const size_t nl = (size_t)std::max(1, s->dev.n_lights);
    const size_t sh_mul = (s->whitted && s->tree_mode != kTreeDirectOne) ? nl : 1;
    const size_t mis_mul = (s->whitted && s->tree_mode != kTreeWhitted) ? sh_mul : 1;
    size_t b = 2 * (32 + 4) + 16 + 4 + 4;               // ray queues + path ids, hit, b2, instance
    b += sh_mul * (32 + 1) + mis_mul * (32 + 16 + 4);   // shadow / MIS queues
    b += 16 + 16 + 8 + 4 + 4 * 16 + 4 + 1 + 4;          // L, beta, sample index, meta, pending records, pending list, key, sorted
    if (s->dev.spatial) b += 4;
    if (s->needs_cam_diff) b += 48;
    if (s->whitted) {
        b += sh_mul * 16 + (size_t)std::max(1, s->dev.max_depth) * (s->needs_cam_diff ? 96 : 48);
        if (s->tree_mode != kTreeWhitted) b += sh_mul * (16 + 8);
    }
    return b + 24;                                       // per-sample radiance + film position
}

// Paths per wave.  Every bounce of a wave costs three traversal launches (each a persistent kernel with its own tail) and
// a few small kernels; measured on C3 (1080p @ 64 spp = 1.3e8 paths): 256 ms per image with 2^22-path waves, 203 ms with
// 2^24, 186 ms with 2^26 (profiles/r1_wave_size.txt).  The wave is therefore as large as the render needs, up to
// kWaveCapMax paths, within the memory budget: b200pt_scene_set_memory_budget / B200PT_MEM_BUDGET (bytes; K / M / G
// suffix), default 60 % of what is free on the device.  Memory no longer grows with resolution x samples per pixel: the
// film keeps running sums between waves (k_film).  B200PT_WAVE_LOG2 overrides the cap (A/B knob).
static const long long kWaveCapMax = 1ll << 26;
static size_t env_bytes(const char* name) {
    const char* e = std::getenv(name);
    if (!e || !*e) return 0;
    char* end = nullptr;
    double v = std::strtod(e, &end);
    if (end && (*end == 'k' || *end == 'K')) v *= 1024.0;
    else if (end && (*end == 'm' || *end == 'M')) v *= 1024.0 * 1024.0;
    else if (end && (*end == 'g' || *end == 'G')) v *= 1024.0 * 1024.0 * 1024.0;
    return v > 0.0 ? (size_t)v : 0;
}
static int wave_cap_for(const SceneImpl* s, long long n_paths) {
    long long cap = kWaveCapMax;
    if (const char* e = std::getenv("B200PT_WAVE_LOG2")) {
        int v = std::atoi(e);
        if (v >= 10 && v <= 27) cap = 1ll << v;
    }
    long long need = 1ll << 14;
    while (need < n_paths && need < cap) need <<= 1;
    cap = std::min(cap, need);
    size_t budget = s->mem_budget ? s->mem_budget : env_bytes("B200PT_MEM_BUDGET");
    if (!budget) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            budget = (size_t)(0.6 * (double)(free_b + (size_t)s->wave_cap * wave_bytes_per_path(s)));  // what this scene's wave already holds is reusable
        else cudaGetLastError();
    }
    if (budget) {
        const size_t per_path = wave_bytes_per_path(s);
        while (cap > 1024 && (size_t)cap * per_path > budget) cap >>= 1;
    }
    // Whitted / DirectLighting "all" test one shadow ray per light and node: keep paths x lights within 2^26 slots
    if (s->whitted && s->tree_mode != kTreeDirectOne) cap = std::min(cap, std::max<long long>((1ll << 26) / std::max(1, s->dev.n_lights), 1024));
    return (int)cap;
}

template <class T> static int dev_upload(SceneImpl* s, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    B2_CUDA(cudaMalloc(&p, bytes));
    s->allocs.push_back(p);
    if (!v.empty()) B2_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return B200PT_OK;
}
template <class T> static int dev_alloc(SceneImpl* s, size_t n, T** out) {
    void* p = nullptr;
    B2_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    s->allocs.push_back(p);
    *out = (T*)p;
    return B200PT_OK;
}

// TrowbridgeReitzDistribution::roughness_to_alpha (trowbridge_reitz.rs:45-53), host f32
static float roughness_to_alpha(float roughness) {
    roughness = roughness > 1e-3f ? roughness : 1e-3f;
    float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
static float clamp0(float v) { return v < 0.0f ? 0.0f : (v > INFINITY ? INFINITY : v); }
static bool black3(const float* c) { return !(c[0] != 0.0f) && !(c[1] != 0.0f) && !(c[2] != 0.0f); }
static float alpha_clamp(float a) { return 0.001f > a ? 0.001f : a; }  // TrowbridgeReitzDistribution::new: max(0.001, alpha)

// materials/src/{matte,plastic,glass,metal}.rs compute_scattering_functions with constant textures,
// evaluated once per material instead of once per intersection.
static DMaterial make_material(const b200pt_material& m, bool allow_multiple_lobes) {
    DMaterial d;
    std::memset(&d, 0, sizeof(d));
    d.type = m.type;
    auto lobe = [&](int kind, uint32_t type) -> DBxDF& {
        DBxDF& x = d.bx[d.n_bxdf++];
        x.kind = kind; x.type = type;
        x.fr_eta_i = x.fr_eta_t = 1.0f; x.ax = x.ay = 1.0f; x.eta_a = x.eta_b = 1.0f;
        return x;
    };
    switch (m.type) {
        case B200PT_MAT_MATTE: {
            float r[3] = {clamp0(m.kd[0]), clamp0(m.kd[1]), clamp0(m.kd[2])};
            float sig = m.sigma < 0.0f ? 0.0f : (m.sigma > 90.0f ? 90.0f : m.sigma);
            if (!black3(r)) {
                if (sig == 0.0f) { DBxDF& x = lobe(BX_LAMBERT, BSDF_REFLECTION | BSDF_DIFFUSE); std::memcpy(x.r, r, 12); }
                else {
                    DBxDF& x = lobe(BX_OREN_NAYAR, BSDF_REFLECTION | BSDF_DIFFUSE);
                    std::memcpy(x.r, r, 12);
                    float sg = sig * (3.14159265358979323846f / 180.0f), s2 = sg * sg;  // oren_nayar.rs:20-31
                    x.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
                    x.on_b = 0.45f * s2 / (s2 + 0.09f);
                }
            }
            break;
        }
        case B200PT_MAT_PLASTIC: {
            float kd[3] = {clamp0(m.kd[0]), clamp0(m.kd[1]), clamp0(m.kd[2])}, ks[3] = {clamp0(m.ks[0]), clamp0(m.ks[1]), clamp0(m.ks[2])};
            if (!black3(kd)) { DBxDF& x = lobe(BX_LAMBERT, BSDF_REFLECTION | BSDF_DIFFUSE); std::memcpy(x.r, kd, 12); }
            if (!black3(ks)) {
                DBxDF& x = lobe(BX_MF_REFL, BSDF_REFLECTION | BSDF_GLOSSY);
                std::memcpy(x.r, ks, 12);
                x.conductor = 0; x.fr_eta_i = 1.5f; x.fr_eta_t = 1.0f;
                float rough = m.urough;
                if (m.remap_roughness) rough = roughness_to_alpha(rough);
                x.ax = x.ay = alpha_clamp(rough);
            }
            break;
        }
        case B200PT_MAT_GLASS: {
            float eta = m.eta[0], ur = m.urough, vr = m.vrough;
            float r[3] = {clamp0(m.ks[0]), clamp0(m.ks[1]), clamp0(m.ks[2])}, t[3] = {clamp0(m.kt[0]), clamp0(m.kt[1]), clamp0(m.kt[2])};
            if (!(black3(r) && black3(t))) {
                bool is_spec = ur == 0.0f && vr == 0.0f;
                if (is_spec && !allow_multiple_lobes) {  // whitted.rs:76 passes false: two delta lobes (glass.rs:112-120)
                    if (!black3(r)) {
                        DBxDF& x = lobe(BX_SPEC_REFL, BSDF_REFLECTION | BSDF_SPECULAR);
                        std::memcpy(x.r, r, 12);
                        x.conductor = 0; x.fr_eta_i = 1.0f; x.fr_eta_t = eta;
                    }
                    if (!black3(t)) {
                        DBxDF& x = lobe(BX_SPEC_TRANS, BSDF_TRANSMISSION | BSDF_SPECULAR);
                        std::memcpy(x.t, t, 12);
                        x.eta_a = 1.0f; x.eta_b = eta;
                    }
                } else if (is_spec) {  // allow_multiple_lobes is true on the path integrator (path.rs:145)
                    DBxDF& x = lobe(BX_FRESNEL_SPECULAR, BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR);
                    std::memcpy(x.r, r, 12); std::memcpy(x.t, t, 12);
                    x.eta_a = 1.0f; x.eta_b = eta;
                } else {
                    if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
                    if (!black3(r)) {
                        DBxDF& x = lobe(BX_MF_REFL, BSDF_REFLECTION | BSDF_GLOSSY);
                        std::memcpy(x.r, r, 12);
                        x.conductor = 0; x.fr_eta_i = 1.0f; x.fr_eta_t = eta;
                        x.ax = alpha_clamp(ur); x.ay = alpha_clamp(vr);
                    }
                    if (!black3(t)) {
                        DBxDF& x = lobe(BX_MF_TRANS, BSDF_TRANSMISSION | BSDF_GLOSSY);
                        std::memcpy(x.t, t, 12);
                        x.eta_a = 1.0f; x.eta_b = eta;
                        x.ax = alpha_clamp(ur); x.ay = alpha_clamp(vr);
                    }
                }
            }
            break;
        }
        case B200PT_MAT_MIRROR: {  // mirror.rs:45-56: SpecularReflection(Kr, FresnelNoOp)
            float r[3] = {clamp0(m.ks[0]), clamp0(m.ks[1]), clamp0(m.ks[2])};
            if (!black3(r)) {
                DBxDF& x = lobe(BX_SPEC_REFL, BSDF_REFLECTION | BSDF_SPECULAR);
                std::memcpy(x.r, r, 12);
                x.conductor = 2;
            }
            break;
        }
        case B200PT_MAT_METAL: {
            float ur = m.urough, vr = m.vrough;
            if (m.remap_roughness) { ur = roughness_to_alpha(ur); vr = roughness_to_alpha(vr); }
            DBxDF& x = lobe(BX_MF_REFL, BSDF_REFLECTION | BSDF_GLOSSY);
            x.r[0] = x.r[1] = x.r[2] = 1.0f;
            x.conductor = 1;
            std::memcpy(x.c_eta_t, m.eta, 12); std::memcpy(x.c_k, m.k, 12);
            x.ax = alpha_clamp(ur); x.ay = alpha_clamp(vr);
            break;
        }
    }
    return d;
}

// Host f32 helpers with the reference's operation order (host code is built with -ffp-contract=off).
static RGB h_inf_lookup(const float* L, float sx, float sy) {
    float s = sx * 1.0f - 0.5f, t = sy * 1.0f - 0.5f;
    float s0 = std::floor(s), t0 = std::floor(t);
    float ds = s - s0, dt = t - t0;
    RGB l = rgb(L[0], L[1], L[2]);
    return l * (1.0f - ds) * (1.0f - dt) + l * (1.0f - ds) * dt + l * ds * (1.0f - dt) + l * ds * dt;
}
struct HostDistr1D {  // core/src/sampling/distribution_1d.rs:22-48
    std::vector<float> func, cdf;
    float func_int = 0.0f;
    void init(const std::vector<float>& f) {
        func = f;
        size_t n = f.size();
        cdf.assign(n + 1, 0.0f);
        for (size_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (float)n;
        func_int = cdf[n];
        if (func_int == 0.0f) { for (size_t i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n; }
        else { for (size_t i = 1; i < n + 1; ++i) cdf[i] /= func_int; }
    }
};

static int wave_alloc(SceneImpl* s, int cap) {
    Wave& W = s->wave;
    if (!s->d_totals) {
        B2_CUDA(cudaMalloc(&s->d_totals, 4 * sizeof(unsigned long long)));
        B2_CUDA(cudaMallocHost(&s->h_pinned, 64 * sizeof(int)));
    }
    if (s->wave_cap == cap) return B200PT_OK;  // shrinks too: a smaller budget must be honoured
    for (auto& g : s->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);  // captured launches hold the old buffers' addresses
    s->graphs.clear();
    for (void* p : s->wave_ptrs) cudaFree(p);  // grow: a later render needs a larger wave
    s->wave_ptrs.clear();
    s->wave_cap = 0;
    const size_t n_before = s->allocs.size();
    struct MoveOut {  // the wave's buffers are owned by wave_ptrs, not by the scene-lifetime list
        SceneImpl* s; size_t n0;
        ~MoveOut() { while (s->allocs.size() > n0) { s->wave_ptrs.push_back(s->allocs.back()); s->allocs.pop_back(); } }
    } move_out{s, n_before};
    for (int k = 0; k < 2; ++k) {
        int rc = dev_alloc(s, (size_t)cap * 2, &W.ray[k]); if (rc) return rc;
        rc = dev_alloc(s, (size_t)cap, &W.qpid[k]); if (rc) return rc;
    }
    int rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.hit))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.hit_b2))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.hit_inst))) return rc;
    const size_t sh_cap = (s->whitted && s->tree_mode != kTreeDirectOne) ? (size_t)cap * (size_t)std::max(1, s->dev.n_lights) : (size_t)cap;
    const size_t mis_cap = (s->whitted && s->tree_mode != kTreeWhitted) ? sh_cap : (size_t)cap;
    if ((rc = dev_alloc(s, sh_cap * 2, &W.sh_ray))) return rc;
    if ((rc = dev_alloc(s, sh_cap, &W.sh_occ))) return rc;
    W.wstack = nullptr; W.sh_c = nullptr; W.dp_b = nullptr; W.dp_c = nullptr;
    if (s->whitted) {
        if ((rc = dev_alloc(s, sh_cap, &W.sh_c))) return rc;
        W.wstack_n = s->needs_cam_diff ? 6 : 3;  // with ray differentials a parked transmission child keeps its own (3 float4 more)
        if ((rc = dev_alloc(s, (size_t)cap * (size_t)std::max(1, s->dev.max_depth) * (size_t)W.wstack_n, &W.wstack))) return rc;
        if (s->tree_mode != kTreeWhitted) {
            if ((rc = dev_alloc(s, sh_cap, &W.dp_b))) return rc;
            if ((rc = dev_alloc(s, sh_cap, &W.dp_c))) return rc;
        }
    }
    if ((rc = dev_alloc(s, mis_cap * 2, &W.mis_ray))) return rc;
    if ((rc = dev_alloc(s, mis_cap, &W.mis_hit))) return rc;
    if ((rc = dev_alloc(s, mis_cap, &W.mis_b2))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.L))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.beta))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.hidx))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.meta))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.pend_a))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.pend_b))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.pend_c))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.pend_d))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.pend_q))) return rc;
    if ((rc = dev_alloc(s, (size_t)(kSegIters + 1) * kCtl, &s->d_ctl))) return rc;
    W.counters = s->d_ctl;
    if ((rc = dev_alloc(s, (size_t)cap, &W.key))) return rc;
    if ((rc = dev_alloc(s, (size_t)cap, &W.sorted))) return rc;
    W.deferred = nullptr;
    if (s->dev.spatial && (rc = dev_alloc(s, (size_t)cap, &W.deferred))) return rc;
    W.cam_diff = nullptr;
    if (s->needs_cam_diff && (rc = dev_alloc(s, (size_t)cap * 3, &W.cam_diff))) return rc;
    s->wave_cap = cap;
    return B200PT_OK;
}

// The shade stage of one bounce: one launch per material class present in the scene over its range of the sorted queue
// (bin 0 = escaped rays), each compiled with only the lobes that class can produce.  Register budgets (CTAs per SM) per
// class from -Xptxas -v / ncu: profiles/r2_shade_split.txt.
static const uint32_t kKmMatte = (1u << BX_LAMBERT) | (1u << BX_OREN_NAYAR);
static const uint32_t kKmPlastic = (1u << BX_LAMBERT) | (1u << BX_MF_REFL) | KM_DIEL;
static const uint32_t kKmGlass = (1u << BX_FRESNEL_SPECULAR) | (1u << BX_MF_REFL) | (1u << BX_MF_TRANS) | KM_DIEL;
static const uint32_t kKmMetal = (1u << BX_MF_REFL) | KM_COND;
static const uint32_t kKmMirror = (1u << BX_SPEC_REFL);
static int shade_grid(const SceneImpl* s, int n_upper, int per_sm) {
    const DevCtx* c = dev_ctx(s->device);
    return std::max(1, std::min((n_upper + 127) / 128, (c ? c->sm_count : 148) * per_sm));
}
template <uint32_t KM>
static void launch_shade_class(SceneImpl* s, const Wave& W, int cur, int n_upper, int bin, int blocks, cudaStream_t st) {
    if (st != s->shade_main) cudaStreamWaitEvent(st, s->ev_fork, 0);  // a class on its own stream starts after the sort
    if constexpr ((KM & KM_TEX) != 0 || KM == kKmMirror) k_shade<KM, kShadeHit, 4><<<shade_grid(s, n_upper, 16), 128, 0, st>>>(s->dev, W, cur, bin, bin + 1);  // textured Kd, mirror: one register budget
    else switch (blocks) {
        case 3: k_shade<KM, kShadeHit, 3><<<shade_grid(s, n_upper, 12), 128, 0, st>>>(s->dev, W, cur, bin, bin + 1); break;
        case 4: k_shade<KM, kShadeHit, 4><<<shade_grid(s, n_upper, 16), 128, 0, st>>>(s->dev, W, cur, bin, bin + 1); break;
        case 6: k_shade<KM, kShadeHit, 6><<<shade_grid(s, n_upper, 24), 128, 0, st>>>(s->dev, W, cur, bin, bin + 1); break;
        default: k_shade<KM, kShadeHit, 5><<<shade_grid(s, n_upper, 20), 128, 0, st>>>(s->dev, W, cur, bin, bin + 1); break;
    }
    if (st != s->shade_main) {  // join: the main stream continues after this class
        for (int k = 0; k < 3; ++k)
            if (st == s->aux[k]) { cudaEventRecord(s->ev_aux[k], st); cudaStreamWaitEvent(s->shade_main, s->ev_aux[k], 0); }
    }
}
static void launch_shade(SceneImpl* s, const Wave& W, int cur, int n_upper, cudaStream_t st) {
    static const int mode = [] { const char* e = std::getenv("B200PT_SHADE_SPLIT"); return e ? std::atoi(e) : 1; }();
    if (mode == 0) {  // A/B: one generic kernel over the whole sorted queue
        k_shade<KM_ALL, kShadeHit, 5><<<shade_grid(s, n_upper, 20), 128, 0, st>>>(s->dev, W, cur, 0, kBins);
        g_launches.fetch_add(1);
        return;
    }
    // CTAs per SM per class: 4 = 118-128 registers, no spills (5 spills 120-212 B, 6 spills 230-400 B; C3: 4444 180 ms, 4555 190 ms,
    // 5555 191 ms, 3333 193 ms).  A/B knob: B200PT_SHADE_BLOCKS = four digits, matte plastic glass metal
    static const int blocks[4] = {
        [] { const char* e = std::getenv("B200PT_SHADE_BLOCKS"); return e && std::strlen(e) == 4 ? e[0] - '0' : 4; }(),
        [] { const char* e = std::getenv("B200PT_SHADE_BLOCKS"); return e && std::strlen(e) == 4 ? e[1] - '0' : 4; }(),
        [] { const char* e = std::getenv("B200PT_SHADE_BLOCKS"); return e && std::strlen(e) == 4 ? e[2] - '0' : 4; }(),
        [] { const char* e = std::getenv("B200PT_SHADE_BLOCKS"); return e && std::strlen(e) == 4 ? e[3] - '0' : 4; }()};
    const int n_classes = __builtin_popcount(s->material_classes & 0xfu);
    int launches = 0;
    // Default (mode 1): one launch per class, each class on its own stream.  With a full queue every launch fills the GPU
    // and they run one after the other as before; with a nearly empty one (late bounces, small renders, one GPU's share
    // of a multi-GPU render) the ~50 us dependent chains of the classes overlap instead of adding up
    // (profiles/r2_launches_floor_c3_tiny.csv: 183 us of a 256 us bounce floor were four class launches in a row).
    // Putting the classes into ONE kernel (mode 3, k_shade_classes) overlaps them too but mixes four code paths on every
    // SM: C3 174 -> 216 ms per image (instruction cache).  Mode 2: all classes on one stream (the first form of the split).
    s->shade_main = st;
    cudaStream_t cs[4] = {st, st, st, st};
    if (mode == 1 && s->overlap && n_classes > 1) {
        cudaEventRecord(s->ev_fork, st);
        int k = 0;
        for (int c = 0; c < 4; ++c)
            if (s->material_classes & (1u << c)) { cs[c] = k == 0 ? st : s->aux[k - 1]; ++k; }
    }
    k_shade<KM_ALL, kShadeMiss, 8><<<shade_grid(s, n_upper, 16), 128, 0, st>>>(s->dev, W, cur, 0, 1);
    ++launches;
    if (mode == 3 && n_classes > 1 && !s->has_kd_tex && !(s->material_classes & (1u << B200PT_MAT_MIRROR))) {
        k_shade_classes<4><<<std::max(1, std::min((n_upper + 127) / 128 + 4, (dev_ctx(s->device) ? dev_ctx(s->device)->sm_count : 148) * 16)), 128, 0, st>>>(s->dev, W, cur);
        ++launches;
    } else {
        // the main stream's class first: the joins of the other streams (enqueued by launch_shade_class) must come after it
        for (int pass = 0; pass < 2; ++pass) {
            const bool main_pass = pass == 0;
            if ((s->material_classes & (1u << B200PT_MAT_MATTE)) && (cs[0] == st) == main_pass) {
                if (s->has_kd_tex) launch_shade_class<kKmMatte | KM_TEX>(s, W, cur, n_upper, 1, blocks[0], cs[0]);
                else launch_shade_class<kKmMatte>(s, W, cur, n_upper, 1, blocks[0], cs[0]);
                ++launches;
            }
            if ((s->material_classes & (1u << B200PT_MAT_PLASTIC)) && (cs[1] == st) == main_pass) {
                if (s->has_kd_tex) launch_shade_class<kKmPlastic | KM_TEX>(s, W, cur, n_upper, 2, blocks[1], cs[1]);
                else launch_shade_class<kKmPlastic>(s, W, cur, n_upper, 2, blocks[1], cs[1]);
                ++launches;
            }
            if ((s->material_classes & (1u << B200PT_MAT_GLASS)) && (cs[2] == st) == main_pass) { launch_shade_class<kKmGlass>(s, W, cur, n_upper, 3, blocks[2], cs[2]); ++launches; }
            if ((s->material_classes & (1u << B200PT_MAT_METAL)) && (cs[3] == st) == main_pass) { launch_shade_class<kKmMetal>(s, W, cur, n_upper, 4, blocks[3], cs[3]); ++launches; }
            if ((s->material_classes & (1u << B200PT_MAT_MIRROR)) && main_pass) { launch_shade_class<kKmMirror>(s, W, cur, n_upper, 5, 4, st); ++launches; }  // a small kernel, on the main stream
        }
    }
    if (s->has_null_material) { k_shade<KM_ALL, kShadeNull, 8><<<shade_grid(s, n_upper, 16), 128, 0, st>>>(s->dev, W, cur, kBinNull, kBinNull + 1); ++launches; }
    g_launches.fetch_add(launches);
}

// Runs the bounce loop for the n paths currently initialised in the wave (queue 0).
// Nothing in the loop waits for the device: every queue size a bounce produces stays in the wave's control blocks and the
// next kernels read it from there (grids are sized for the wave and stride; persistent traversal kernels take the count
// pointer), so the host only enqueues.  The three traversals a bounce produces - shadow rays, MIS rays, and the next
// bounce's closest-hit rays - are independent of one another and go to three streams so that the drain of one persistent
// kernel overlaps the others; resolve, which needs the first two, and the next shade, which needs all three, follow on the
// main stream behind events.  B200PT_OVERLAP=0 puts everything back on one stream (A/B).
static int aux_setup(SceneImpl* s) {
    if (s->aux_ready) return B200PT_OK;
    for (int i = 0; i < 3; ++i) {
        B2_CUDA(cudaStreamCreateWithFlags(&s->aux[i], cudaStreamNonBlocking));
        B2_CUDA(cudaEventCreateWithFlags(&s->ev_aux[i], cudaEventDisableTiming));
    }
    B2_CUDA(cudaEventCreateWithFlags(&s->ev_shade, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    const char* e = std::getenv("B200PT_OVERLAP");
    s->overlap = !(e && e[0] == '0');
    s->aux_ready = true;
    return B200PT_OK;
}

static int run_wave(SceneImpl* s, int n, cudaStream_t st) {
    int rc = aux_setup(s);
    if (rc) return rc;
    if (n <= 0) return B200PT_OK;
    const long long launches_before = g_launches.load();
    struct CountLaunches { SceneImpl* s; long long before; ~CountLaunches() { s->wave_launches = g_launches.load() - before; } } count_launches{s, launches_before};
    const Wave& W0 = s->wave;
    int* const ctl = s->d_ctl;
    const DevCtx* dc = dev_ctx(s->device);
    const int sms = dc ? dc->sm_count : 148;
    const int g256 = std::max(1, std::min((n + 255) / 256, sms * 16));
    auto work_ctr = [&](int block, int which) { return reinterpret_cast<unsigned long long*>(ctl + block * kCtl + 32) + which; };
    auto closest = [&](int q, int block) {  // the queue whose size is control block `block`[0]
        TraceLaunch tl; tl.n_dev = ctl + block * kCtl; tl.work_ctr = work_ctr(block, 0);
        return s->instanced ? launch_intersect2(s->accel2.dev, W0.ray[q], n, W0.hit, st, W0.hit_b2, W0.hit_inst, 0, &tl)
                            : launch_intersect(s->dev.accel, W0.ray[q], n, W0.hit, st, 0, W0.hit_b2, &tl);
    };
    // iteration k shades the vertices with `bounces` == k; paths end at max_depth, except through null materials
    int iters_left = s->has_null_material ? 0x7fffffff : s->dev.max_depth + 1;
    int cur = 0, n_seg = n;
    bool first_seg = true;
    while (n_seg > 0 && iters_left > 0) {
        const int seg_iters = std::min(kSegIters, iters_left);
        k_wave_begin<<<1, 256, 0, st>>>(ctl, (kSegIters + 1) * kCtl, n_seg);
        g_launches.fetch_add(1);
        if ((rc = closest(cur, 0))) return rc;
        for (int it = 0; it < seg_iters; ++it) {
            Wave W = W0;
            W.counters = ctl + (it + 1) * kCtl;
            const int* n_act = ctl + it * kCtl;
            k_bin_count<<<g256, 256, 0, st>>>(s->dev, W, n_act, 0);
            k_bin_scatter<<<g256, 256, 0, st>>>(W, n_act, 0);
            g_launches.fetch_add(2);
            launch_shade(s, W, cur, n, st);
            if (s->dev.spatial) {
                // SpatialLightDistribution: fill the voxels this bounce touched for the first time, then shade the parked slots
                k_voxel_contrib<<<sms * 8, 128, 0, st>>>(s->dev, W.counters + 24);
                k_voxel_finish<<<sms, 128, 0, st>>>(s->dev, W.counters + 24);
                k_shade<KM_ALL, kShadeHit, 5><<<shade_grid(s, n, 20), 128, 0, st>>>(s->dev, W, cur, -1, -1);
                g_launches.fetch_add(3);
            }
            cudaStream_t s_sh = s->overlap ? s->aux[0] : st, s_mis = s->overlap ? s->aux[1] : st;
            if (s->overlap) {
                B2_CUDA(cudaEventRecord(s->ev_shade, st));
                B2_CUDA(cudaStreamWaitEvent(s_sh, s->ev_shade, 0));
                B2_CUDA(cudaStreamWaitEvent(s_mis, s->ev_shade, 0));
            }
            {
                TraceLaunch tl; tl.n_dev = W.counters + 1; tl.work_ctr = work_ctr(it + 1, 1);
                rc = s->instanced ? launch_occluded2(s->accel2.dev, W.sh_ray, n, W.sh_occ, s_sh, 0, &tl) : launch_occluded(s->dev.accel, W.sh_ray, n, W.sh_occ, s_sh, 0, &tl);
                if (rc) return rc;
                if (s->overlap) B2_CUDA(cudaEventRecord(s->ev_aux[0], s_sh));
            }
            if (s->dev.n_lights > s->n_point_lights) {  // estimate_direct samples the BSDF only for lights that can be hit
                TraceLaunch tl; tl.n_dev = W.counters + 2; tl.work_ctr = work_ctr(it + 1, 2);
                rc = s->instanced ? launch_intersect2(s->accel2.dev, W.mis_ray, n, W.mis_hit, s_mis, W.mis_b2, nullptr, 0, &tl)
                                  : launch_intersect(s->dev.accel, W.mis_ray, n, W.mis_hit, s_mis, 0, W.mis_b2, &tl);
                if (rc) return rc;
            }
            if (s->overlap) B2_CUDA(cudaEventRecord(s->ev_aux[1], s_mis));
            if (it + 1 < seg_iters && (rc = closest(cur ^ 1, it + 1))) return rc;  // the next bounce's rays, concurrently with the two above
            if (s->overlap) {
                B2_CUDA(cudaStreamWaitEvent(st, s->ev_aux[0], 0));
                B2_CUDA(cudaStreamWaitEvent(st, s->ev_aux[1], 0));
            }
            k_resolve<<<g256, 256, 0, st>>>(s->dev, W);
            g_launches.fetch_add(1);
            cur ^= 1;
        }
        k_wave_end<<<1, 32, 0, st>>>(ctl, seg_iters + 1, first_seg ? n : 0, s->d_totals);
        g_launches.fetch_add(1);
        first_seg = false;
        iters_left -= seg_iters;
        if (iters_left <= 0) break;
        // deeper than one segment of control blocks (maxdepth > 15, or null materials): the only read-back of the loop
        B2_CUDA(cudaMemcpyAsync(s->h_pinned, ctl + seg_iters * kCtl, sizeof(int), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        n_seg = s->h_pinned[0];
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wavefront kernels");
    return B200PT_OK;
}

// The bounce loop of a wave replayed from a CUDA graph.  run_wave only enqueues (no read-back when maxdepth + 1 fits one
// control segment), so its launches - sort, shade classes on their streams, three traversals, resolve, per bounce - can be
// captured once per wave size and replayed: the device then runs a wave's ~105 launches without waiting for the host,
// which matters on a busy host (eight ranks and their clock samplers on 16 threads: single renders showed 37-56 ms
// spikes around a 24.7 ms median, profiles/r2_shard_and_floor.txt).  The first wave of a given size runs eagerly (first-use
// setup of streams, counters and occupancy queries must not happen inside a capture), the second is captured, later ones
// are replayed.  Any capture error falls back to the eager loop for the rest of the scene's life.  B200PT_GRAPH=0 disables it.
static int run_wave_graphed(SceneImpl* s, int n, cudaStream_t st) {
    static const bool enabled = [] { const char* e = std::getenv("B200PT_GRAPH"); return !(e && e[0] == '0'); }();
    const bool eligible = enabled && !s->graph_failed && !s->whitted && !s->zt_seq && !s->has_null_material && s->dev.max_depth + 1 <= kSegIters &&
                          s->dev.sampler_type != B200PT_SAMPLER_ZEROTWO && n > 0;
    if (!eligible) return run_wave(s, n, st);
    SceneImpl::WaveGraph* g = nullptr;
    for (auto& e : s->graphs) if (e.n == n) g = &e;
    if (!g) { s->graphs.push_back({}); g = &s->graphs.back(); g->n = n; }
    if (g->uses++ == 0) {
        // first wave of this size: eager, then the same launches are captured for the next one (nothing runs during capture)
        int rc = run_wave(s, n, st);
        if (rc) return rc;
        if (!s->graph_stream) {
            if (cudaStreamCreateWithFlags(&s->graph_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&s->ev_graph_in, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&s->ev_graph_out, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); s->graph_failed = true; return B200PT_OK; }
        }
        // capture on an own stream (the caller's may be the legacy default stream, which cannot be captured)
        cudaGraph_t graph = nullptr;
        const long long launches_before = g_launches.load();
        cudaError_t e = cudaStreamBeginCapture(s->graph_stream, cudaStreamCaptureModeThreadLocal);
        int rc2 = B200PT_OK;
        if (e == cudaSuccess) {
            rc2 = run_wave(s, n, s->graph_stream);
            e = cudaStreamEndCapture(s->graph_stream, &graph);
        }
        g_launches.fetch_sub(g_launches.load() - launches_before > 0 ? s->wave_launches : 0);  // nothing was launched during the capture
        if (e == cudaSuccess && rc2 == B200PT_OK && graph) e = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess || rc2 != B200PT_OK || !g->exec) {
            cudaGetLastError();
            g->exec = nullptr;
            s->graph_failed = true;
        }
        return B200PT_OK;
    }
    if (!g->exec) return run_wave(s, n, st);
    B2_CUDA(cudaEventRecord(s->ev_graph_in, st));
    B2_CUDA(cudaStreamWaitEvent(s->graph_stream, s->ev_graph_in, 0));
    B2_CUDA(cudaGraphLaunch(g->exec, s->graph_stream));
    B2_CUDA(cudaEventRecord(s->ev_graph_out, s->graph_stream));
    B2_CUDA(cudaStreamWaitEvent(st, s->ev_graph_out, 0));
    // launches of the replayed wave (same count as the eager loop enqueues)
    g_launches.fetch_add(s->wave_launches);
    return B200PT_OK;
}

// WhittedIntegrator: one tree node per path and iteration until every path has walked its whole tree.
static int run_wave_whitted(SceneImpl* s, int n, cudaStream_t st) {
    Wave& W = s->wave;
    int rc = aux_setup(s);
    if (rc) return rc;
    int cur = 0, n_active = n;
    s->rays[0] += (uint64_t)n;
    const int nl = s->dev.n_lights;
    const long long max_iter = 1ll << std::min(std::max(s->dev.max_depth, 1), 24);  // a binary tree of depth max_depth
    auto closest = [&](int q, int count) {
        s->rays[1] += (uint64_t)count;
        return s->instanced ? launch_intersect2(s->accel2.dev, W.ray[q], count, W.hit, st, W.hit_b2, W.hit_inst)
                            : launch_intersect(s->dev.accel, W.ray[q], count, W.hit, st, 0, W.hit_b2);
    };
    if (n_active > 0 && (rc = closest(cur, n_active))) return rc;
    for (long long iter = 0; n_active > 0 && iter < max_iter; ++iter) {
        B2_CUDA(cudaMemsetAsync(W.counters, 0, kCtl * sizeof(int), st));
        k_bin_count<<<(n_active + 255) / 256, 256, 0, st>>>(s->dev, W, nullptr, n_active);
        k_bin_scatter<<<(n_active + 255) / 256, 256, 0, st>>>(W, nullptr, n_active);
        const int gs = (n_active + 127) / 128;
        if (s->tree_mode == kTreeWhitted) k_shade_tree<kTreeWhitted><<<gs, 128, 0, st>>>(s->dev, W, cur, n_active);
        else if (s->tree_mode == kTreeDirectAll) k_shade_tree<kTreeDirectAll><<<gs, 128, 0, st>>>(s->dev, W, cur, n_active);
        else k_shade_tree<kTreeDirectOne><<<gs, 128, 0, st>>>(s->dev, W, cur, n_active);
        g_launches.fetch_add(3);
        int cnt[7];
        B2_CUDA(cudaMemcpyAsync(cnt, W.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (cnt[4]) {
            b200pt_set_error("whitted / directlighting: a camera sample needs more sampler dimensions (lights x tree nodes) than the sampler has (halton 1000, sobol 1024); the reference asserts here (samplers/src/halton.rs:106-110)");
            return B200PT_ERR_UNSUPPORTED;
        }
        // shadow rays, MIS rays and the next node's closest-hit rays are independent: three streams (see run_wave)
        cudaStream_t s_sh = s->overlap ? s->aux[0] : st, s_mis = s->overlap ? s->aux[1] : st;
        const bool with_mis = s->tree_mode != kTreeWhitted;
        if (cnt[3] > 0) {
            const int64_t n_sh = (int64_t)cnt[3] * (s->tree_mode == kTreeDirectOne ? 1 : nl);
            rc = s->instanced ? launch_occluded2(s->accel2.dev, W.sh_ray, n_sh, W.sh_occ, s_sh) : launch_occluded(s->dev.accel, W.sh_ray, n_sh, W.sh_occ, s_sh, 0);
            if (rc) return rc;
            s->rays[2] += (uint64_t)cnt[5];
            if (s->overlap) B2_CUDA(cudaEventRecord(s->ev_aux[0], s_sh));
            if (with_mis) {  // the BSDF-sampled MIS rays of estimate_direct
                rc = s->instanced ? launch_intersect2(s->accel2.dev, W.mis_ray, n_sh, W.mis_hit, s_mis, W.mis_b2, nullptr)
                                  : launch_intersect(s->dev.accel, W.mis_ray, n_sh, W.mis_hit, s_mis, 0, W.mis_b2);
                if (rc) return rc;
                s->rays[1] += (uint64_t)cnt[6];
                if (s->overlap) B2_CUDA(cudaEventRecord(s->ev_aux[1], s_mis));
            }
        }
        const bool more = cnt[0] > 0 && iter + 1 < max_iter;
        if (more && (rc = closest(cur ^ 1, cnt[0]))) return rc;
        if (cnt[3] > 0) {
            if (s->overlap) {
                B2_CUDA(cudaStreamWaitEvent(st, s->ev_aux[0], 0));
                if (with_mis) B2_CUDA(cudaStreamWaitEvent(st, s->ev_aux[1], 0));
            }
            const int gr = (cnt[3] + 255) / 256;
            if (s->tree_mode == kTreeWhitted) k_resolve_tree<kTreeWhitted><<<gr, 256, 0, st>>>(s->dev, W, cnt[3]);
            else if (s->tree_mode == kTreeDirectAll) k_resolve_tree<kTreeDirectAll><<<gr, 256, 0, st>>>(s->dev, W, cnt[3]);
            else k_resolve_tree<kTreeDirectOne><<<gr, 256, 0, st>>>(s->dev, W, cnt[3]);
            g_launches.fetch_add(1);
        }
        cur ^= 1;
        n_active = cnt[0];
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wavefront kernels (whitted)");
    return B200PT_OK;
}

}  // namespace b2

struct b200pt_scene {
    b2::SceneImpl impl;
};

using namespace b2;

extern "C" {

int b200pt_scene_create(const b200pt_scene_desc* d, b200pt_scene** out) {
    if (!out) { b200pt_set_error("b200pt_scene_create: out is null"); return B200PT_ERR_INVALID; }
    *out = nullptr;
    int rc = require_device();
    if (rc) return rc;
    if (!d || d->n_prims < 0 || d->n_nodes < 0 || (d->n_prims > 0 && (!d->nodes || !d->ordered_prims || !d->tri_verts || !d->prim_material)) ||
        (d->n_materials > 0 && !d->materials) || (d->n_lights > 0 && !d->lights)) {
        b200pt_set_error("b200pt_scene_create: invalid scene description");
        return B200PT_ERR_INVALID;
    }
    if (d->sampler.type != B200PT_SAMPLER_HALTON && d->sampler.type != B200PT_SAMPLER_ZEROTWO && d->sampler.type != B200PT_SAMPLER_SOBOL) {
        b200pt_set_error("b200pt_scene_create: unknown sampler type (halton, 02sequence and sobol are on this path)");
        return B200PT_ERR_UNSUPPORTED;
    }
    if (d->sampler.type == B200PT_SAMPLER_SOBOL && !d->sobol_matrices_32) {
        b200pt_set_error("b200pt_scene_create: the sobol sampler needs scene_desc.sobol_matrices_32 (SOBOL_MATRICES_32, 1024 x 52 u32)");
        return B200PT_ERR_INVALID;
    }
    if (d->integrator.type != B200PT_INTEGRATOR_PATH && d->integrator.type != B200PT_INTEGRATOR_WHITTED && d->integrator.type != B200PT_INTEGRATOR_DIRECT) {
        b200pt_set_error("b200pt_scene_create: unknown integrator type (path, whitted and directlighting are on this path)");
        return B200PT_ERR_UNSUPPORTED;
    }
    if (d->integrator.type == B200PT_INTEGRATOR_DIRECT && d->integrator.direct_strategy != B200PT_DIRECT_ALL && d->integrator.direct_strategy != B200PT_DIRECT_ONE) {
        b200pt_set_error("b200pt_scene_create: unknown directlighting strategy");
        return B200PT_ERR_INVALID;
    }
    if (d->integrator.type != B200PT_INTEGRATOR_PATH) {
        // The number of get_2d() calls of one camera sample depends on the tree it spawns; the (0,2) sampler would fall
        // back to the tile RNG (see below).  Halton is a pure function of (pixel, sample, dimension).
        if (d->sampler.type == B200PT_SAMPLER_ZEROTWO) { b200pt_set_error("b200pt_scene_create: the whitted / directlighting integrators need the halton or sobol sampler on this path"); return B200PT_ERR_UNSUPPORTED; }
        if (d->integrator.max_depth < 0 || d->integrator.max_depth > 24) { b200pt_set_error("b200pt_scene_create: whitted / directlighting maxdepth must be in [0, 24]"); return B200PT_ERR_UNSUPPORTED; }
        if ((long long)d->n_lights * 1024 > (1ll << 24)) { b200pt_set_error("b200pt_scene_create: whitted: more than 16384 lights"); return B200PT_ERR_UNSUPPORTED; }
    }
    if (d->sampler.type == B200PT_SAMPLER_ZEROTWO) {
        // Past its pre-generated slots the reference's PixelSampler draws from the TILE's RNG inside li(), which makes the
        // stream position of every later pixel depend on earlier path lengths (SURVEY §7 "ZeroTwo sequencing").  When
        // "dimensions" covers the path's worst case li() never touches the RNG and all samples run in parallel; otherwise
        // (the reference's default is 4) the render takes the tile-sequential mode (ZtSeq), one path per tile at a time.
        if (d->sampler.dimensions < 0 || d->sampler.dimensions > 255) {
            b200pt_set_error("b200pt_scene_create: 02sequence \"dimensions\" must be in [0, 255]");
            return B200PT_ERR_UNSUPPORTED;
        }
        if (d->sampler.spp > 32768) { b200pt_set_error("b200pt_scene_create: 02sequence pixelsamples > 32768"); return B200PT_ERR_UNSUPPORTED; }
    }
    if (d->integrator.type == B200PT_INTEGRATOR_PATH) {
        // Sampler tables bound the path length: a vertex draws up to 8 dimensions after the 5 of the camera sample; the
        // reference asserts past its tables (samplers/src/halton.rs:106-110).  The bounce counter is 8 bits (255 = unused pixel).
        const int max_dims = d->sampler.type == B200PT_SAMPLER_SOBOL ? 1024 : 1000;
        if (d->integrator.max_depth < 0 || d->integrator.max_depth > 254 ||
            (d->sampler.type != B200PT_SAMPLER_ZEROTWO && 5 + 8 * (long long)d->integrator.max_depth > max_dims)) {
            b200pt_set_error("b200pt_scene_create: path maxdepth beyond what the sampler's tables cover (halton: 124, sobol: 127)");
            return B200PT_ERR_UNSUPPORTED;
        }
    }
    bool has_null = false;
    for (int64_t i = 0; i < d->n_prims; ++i) {
        // material -1: Material "" / "none" (api/src/lib.rs make_material -> None): the path integrator passes through it
        // (path.rs:146-150); an emissive primitive still emits
        if (d->prim_material[i] < -1 || d->prim_material[i] >= d->n_materials) { b200pt_set_error("b200pt_scene_create: primitive material index out of range"); return B200PT_ERR_INVALID; }
        if (d->prim_material[i] < 0) has_null = true;
        if (d->prim_light && (d->prim_light[i] < -1 || d->prim_light[i] >= d->n_lights)) { b200pt_set_error("b200pt_scene_create: primitive light index out of range"); return B200PT_ERR_INVALID; }
    }
    if (has_null && d->integrator.type != B200PT_INTEGRATOR_PATH) {
        b200pt_set_error("b200pt_scene_create: primitives without a material are supported by the path integrator only");
        return B200PT_ERR_UNSUPPORTED;
    }
    b200pt_scene* sc = new b200pt_scene();
    SceneImpl* s = &sc->impl;
    s->device = current_device();
    s->has_null_material = has_null;
    for (int i = 0; i < d->n_materials; ++i)
        if (d->materials[i].type >= 0 && d->materials[i].type < 5) s->material_classes |= 1u << d->materials[i].type;
        else { b200pt_set_error("b200pt_scene_create: unknown material type"); delete sc; return B200PT_ERR_INVALID; }
    for (int i = 0; i < d->n_lights; ++i) if (d->lights[i].type == B200PT_LIGHT_POINT || d->lights[i].type == B200PT_LIGHT_DISTANT || d->lights[i].type == B200PT_LIGHT_SPOT || d->lights[i].type == B200PT_LIGHT_GONIOMETRIC || d->lights[i].type == B200PT_LIGHT_PROJECTION) ++s->n_point_lights;  // delta lights
    auto fail = [&](int code) { b200pt_scene_destroy(sc); return code; };
    DeviceScene& D = s->dev;
    std::memset(&D, 0, sizeof(D));
    if (d->n_objects > 0) {
        if (!d->objects || (d->n_instances > 0 && !d->instances) || d->n_top_tris < 0 || d->n_top_tris > d->n_prims) {
            b200pt_set_error("b200pt_scene_create: invalid instancing description");
            return fail(B200PT_ERR_INVALID);
        }
        for (int64_t i = d->n_top_tris; i < d->n_prims; ++i)
            if (d->prim_light && d->prim_light[i] >= 0) { b200pt_set_error("b200pt_scene_create: area lights inside object instances are not supported (as in pbrt)"); return fail(B200PT_ERR_UNSUPPORTED); }
        rc = accel2_build_device(d, &s->accel2);
        if (rc) return fail(rc);
        s->instanced = true;
        D.accel = s->accel2.dev.top;
        D.instances = s->accel2.dev.instances;
    } else {
        rc = accel_build_device(d->nodes, d->n_nodes, d->ordered_prims, d->tri_verts, d->prim_flags, d->n_prims, &s->accel, d->tri_uvs);
        if (rc) return fail(rc);
        D.accel = s->accel.dev;
    }
    // alpha-mask textures (triangle.rs:587-607, 840-899): evaluated inside the traversal kernels' accept path
    rc = alpha_build_device(d->float_textures, d->n_float_textures, d->prim_alpha_tex, d->tri_uvs, d->prim_flags, d->n_prims, d->noise_perm, &s->alpha);
    if (rc) return fail(rc);
    D.accel.alpha = s->alpha.dev;
    if (s->instanced) s->accel2.dev.top.alpha = s->alpha.dev; else s->accel.dev.alpha = s->alpha.dev;
    s->film = d->film;
    s->sampler = d->sampler;
    s->spp = d->sampler.spp;
    if (d->sampler.type == B200PT_SAMPLER_ZEROTWO || d->sampler.type == B200PT_SAMPLER_SOBOL) { int p2 = 1; while (p2 < s->spp) p2 <<= 1; s->spp = p2; }  // zero_two_sequence.rs:23-32, sobol.rs:28-37
    s->zt_seq = d->sampler.type == B200PT_SAMPLER_ZEROTWO && (d->sampler.dimensions < 1 + 2 * d->integrator.max_depth || d->sampler.dimensions < 2 + 3 * d->integrator.max_depth);
    D.sampler_type = d->sampler.type;

    // primitives in original order with their material / light / flags
    std::vector<float4> pv((size_t)d->n_prims * 3);
    parallel_for(d->n_prims, [&](int64_t i_begin, int64_t i_end) {
    for (int64_t i = i_begin; i < i_end; ++i) {
        const float* v = d->tri_verts + 9 * i;
        int32_t mat = d->prim_material[i], lt = d->prim_light ? d->prim_light[i] : -1;
        uint32_t fl = d->prim_flags ? d->prim_flags[i] : 0u;
        float fm, fl2, ff;
        std::memcpy(&fm, &mat, 4); std::memcpy(&fl2, &lt, 4); std::memcpy(&ff, &fl, 4);
        pv[3 * i] = make_float4(v[0], v[1], v[2], fm);
        pv[3 * i + 1] = make_float4(v[3], v[4], v[5], fl2);
        pv[3 * i + 2] = make_float4(v[6], v[7], v[8], ff);
    }
    });
    if ((rc = dev_upload(s, pv, &D.prim_verts))) return fail(rc);
    // optional vertex attributes (triangle.rs:384-394, 631-721)
    bool any_uv = false, any_n = false, any_s = false;
    for (int64_t i = 0; i < d->n_prims && d->prim_flags; ++i) {
        const uint32_t fl = d->prim_flags[i];
        any_uv |= (fl & B200PT_PRIM_HAS_UV) != 0; any_n |= (fl & B200PT_PRIM_HAS_NORMALS) != 0; any_s |= (fl & B200PT_PRIM_HAS_TANGENTS) != 0;
    }
    if ((any_uv && !d->tri_uvs) || (any_n && !d->tri_normals) || (any_s && !d->tri_tangents)) {
        b200pt_set_error("b200pt_scene_create: a primitive flag announces uvs / normals / tangents but the array is NULL");
        return fail(B200PT_ERR_INVALID);
    }
    if (any_uv) {
        std::vector<float4> duv((size_t)d->n_prims);
        for (int64_t i = 0; i < d->n_prims; ++i) duv[(size_t)i] = record_duv((d->prim_flags[i] & B200PT_PRIM_HAS_UV) ? d->tri_uvs + 6 * i : nullptr);
        if ((rc = dev_upload(s, duv, &D.prim_duv))) return fail(rc);
    }
    if (any_n) {
        std::vector<float> vn(d->tri_normals, d->tri_normals + 9 * d->n_prims);
        if ((rc = dev_upload(s, vn, &D.prim_n))) return fail(rc);
    }
    if (any_s) {
        std::vector<float> vs(d->tri_tangents, d->tri_tangents + 9 * d->n_prims);
        if ((rc = dev_upload(s, vs, &D.prim_s))) return fail(rc);
    }

    // Textured "Kd" (matte / plastic): the material's lobes are built with a placeholder Kd so that the diffuse lobe exists
    // as lobe 0; the shade kernels put the texture's value in per intersection (wavefront.cuh: textured_material).
    std::vector<int> kd_tex((size_t)std::max(1, d->n_materials), -1);
    bool closedform_tex = false;
    if (d->material_kd_tex && d->spectrum_textures && d->n_spectrum_textures > 0) {
        for (int i = 0; i < d->n_materials; ++i) {
            const b200pt_material& m = d->materials[i];
            const int t = d->material_kd_tex[i];
            if (t < 0 || (m.type != B200PT_MAT_MATTE && m.type != B200PT_MAT_PLASTIC)) continue;
            if (t >= d->n_spectrum_textures) { b200pt_set_error("b200pt_scene_create: material_kd_tex index out of range"); return fail(B200PT_ERR_INVALID); }
            const b200pt_spectrum_texture& T = d->spectrum_textures[t];
            if (T.type != B200PT_STEX_CONSTANT && T.type != B200PT_STEX_CHECKERBOARD) { b200pt_set_error("b200pt_scene_create: unknown spectrum texture type (constant, checkerboard)"); return fail(B200PT_ERR_UNSUPPORTED); }
            kd_tex[(size_t)i] = t;
            s->has_kd_tex = true;
            if (T.type == B200PT_STEX_CHECKERBOARD && T.aa_closedform) closedform_tex = true;
        }
    }
    std::vector<DMaterial> mats;
    for (int i = 0; i < d->n_materials; ++i) {
        b200pt_material m = d->materials[i];
        if (kd_tex[(size_t)i] >= 0) m.kd[0] = m.kd[1] = m.kd[2] = 1.0f;
        mats.push_back(make_material(m, d->integrator.type == B200PT_INTEGRATOR_PATH));
    }
    if ((rc = dev_upload(s, mats, &D.materials))) return fail(rc);
    if (s->has_kd_tex) {
        std::vector<DSpecTex> st((size_t)d->n_spectrum_textures);
        for (int k = 0; k < d->n_spectrum_textures; ++k) {
            const b200pt_spectrum_texture& T = d->spectrum_textures[k];
            DSpecTex& o = st[(size_t)k];
            o.type = T.type; o.su = T.su; o.sv = T.sv; o.du = T.du; o.dv = T.dv; o.closedform = T.aa_closedform ? 1 : 0;
            std::memcpy(o.tex1, T.tex1, 12); std::memcpy(o.tex2, T.tex2, 12);
        }
        std::vector<float> uv6((size_t)d->n_prims * 6);
        for (int64_t i = 0; i < d->n_prims; ++i) {
            const bool has = d->tri_uvs && d->prim_flags && (d->prim_flags[i] & B200PT_PRIM_HAS_UV);
            static const float dflt[6] = {0.0f, 0.0f, 1.0f, 0.0f, 1.0f, 1.0f};  // triangle.rs:384-394
            std::memcpy(&uv6[(size_t)i * 6], has ? d->tri_uvs + 6 * i : dflt, 24);
        }
        if ((rc = dev_upload(s, kd_tex, &D.mat_kd_tex))) return fail(rc);
        if ((rc = dev_upload(s, st, &D.spec_tex))) return fail(rc);
        if ((rc = dev_upload(s, uv6, &D.prim_uv6))) return fail(rc);
        s->needs_cam_diff = closedform_tex;
        D.cam_diff_scale = 1.0f / std::sqrt((float)s->spp);
    }

    // Scene::new (core/src/scene.rs:50-77): world bound, infinite lights, Light::preprocess
    V3 wc = mk(0, 0, 0);
    float radius = 0.0f;
    if (d->n_nodes > 0) {  // Bounds3::bounding_sphere, bounds3.rs:196-208
        const float* b = d->nodes[0].bounds;
        V3 lo = mk(b[0], b[1], b[2]), hi = mk(b[3], b[4], b[5]);
        wc = (1.0f - 0.5f) * lo + 0.5f * hi;
        bool inside = (wc.x >= lo.x && wc.x <= hi.x) && (wc.y >= lo.y && wc.y <= hi.y) && (wc.z >= lo.z && wc.z <= hi.z);
        radius = inside ? std::sqrt(length_squared(wc - hi)) : 0.0f;
    }
    D.world_radius = radius;
    std::vector<DLight> lights((size_t)d->n_lights);
    std::vector<int> inf_ids;
    std::vector<DInfDistr> inf_distr;
    std::vector<float> power_y((size_t)d->n_lights);
    for (int i = 0; i < d->n_lights; ++i) {
        const b200pt_light& l = d->lights[i];
        DLight& o = lights[(size_t)i];
        std::memset(&o, 0, sizeof(o));
        o.type = l.type; o.prim = l.prim; o.two_sided = l.two_sided; o.inf_slot = -1;
        std::memcpy(o.pos, l.pos, 12); std::memcpy(o.L, l.L, 12);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) { o.l2w[3 * r + c] = l.light_to_world[4 * r + c]; o.w2l[3 * r + c] = l.world_to_light[4 * r + c]; }
        RGB Lr = rgb(l.L[0], l.L[1], l.L[2]);
        RGB power = rgb1(0.0f);
        if (l.type == B200PT_LIGHT_POINT) power = kFourPi * Lr;  // point.rs:96-98
        else if (l.type == B200PT_LIGHT_DISTANT) power = Lr * kPi * radius * radius;  // distant.rs:92-95
        else if (l.type == B200PT_LIGHT_SPOT) {  // spot.rs:109-111
            power = Lr * (kPi * 2.0f) * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width));
            o.area = l.cos_total_width; o.cos_falloff_start = l.cos_falloff_start;
        }
        else if (l.type == B200PT_LIGHT_AREA) {
            if (l.prim < 0 || l.prim >= d->n_prims) { b200pt_set_error("b200pt_scene_create: area light primitive out of range"); return fail(B200PT_ERR_INVALID); }
            const float* v = d->tri_verts + 9 * (size_t)l.prim;
            V3 p0 = mk(v[0], v[1], v[2]), p1 = mk(v[3], v[4], v[5]), p2 = mk(v[6], v[7], v[8]);
            o.area = 0.5f * std::sqrt(length_squared(cross(p1 - p0, p2 - p0)));  // Triangle::area, triangle.rs:906-911
            float sgn = l.two_sided ? 2.0f : 1.0f;
            power = sgn * Lr * o.area * kPi;  // diffuse.rs:131-134
        } else if (l.type == B200PT_LIGHT_INFINITE) {
            o.inf_slot = (int)inf_distr.size();
            inf_ids.push_back(i);
            // InfiniteAreaLight::new (infinite.rs:61-92): MIPMap, importance image, Distribution2D
            b2host::EnvMapTables em;
            b2host::build_envmap(l.map_rgb, l.map_width, l.map_height, l.L, &em);
            DInfDistr dd;
            std::memset(&dd, 0, sizeof(dd));
            dd.width = em.width; dd.height = em.height; dd.nu = em.nu; dd.nv = em.nv; dd.mfunc_int = em.marg_int;
            const float* tex = nullptr;
            if ((rc = dev_upload(s, em.texels, &tex))) return fail(rc);
            dd.texels = (const float4*)tex;
            if ((rc = dev_upload(s, em.cond_func, &dd.func))) return fail(rc);
            if ((rc = dev_upload(s, em.cond_cdf, &dd.cdf))) return fail(rc);
            if ((rc = dev_upload(s, em.cond_int, &dd.func_int))) return fail(rc);
            if ((rc = dev_upload(s, em.marg_func, &dd.mfunc))) return fail(rc);
            if ((rc = dev_upload(s, em.marg_cdf, &dd.mcdf))) return fail(rc);
            inf_distr.push_back(dd);
            RGB spec = rgb(em.power_lookup[0], em.power_lookup[1], em.power_lookup[2]);  // infinite.rs:177-186
            power = kPi * radius * radius * spec;
        } else if (l.type == B200PT_LIGHT_GONIOMETRIC || l.type == B200PT_LIGHT_PROJECTION) {
            // GonioPhotometricLight::new (goniometric.rs:59-100) / ProjectionLight::new (projection.rs:50-110): MIPMap over the image;
            // power = 4 pi I lookup_triangle((.5, .5), .5) (goniometric.rs:140-150) or lookup * I * 2 pi (1 - cos_total_width) (projection.rs:173-183)
            o.inf_slot = (int)inf_distr.size();
            b2host::EnvMapTables em;
            const float one[3] = {1.0f, 1.0f, 1.0f};
            if (!l.map_rgb || l.map_width <= 0 || l.map_height <= 0) { em.texels = {1.0f, 1.0f, 1.0f, 0.0f}; em.power_lookup[0] = em.power_lookup[1] = em.power_lookup[2] = 1.0f; }  // no map: scale = 1
            else b2host::build_envmap(l.map_rgb, l.map_width, l.map_height, one, &em, false);
            DInfDistr dd;
            std::memset(&dd, 0, sizeof(dd));
            dd.width = em.width; dd.height = em.height;
            const float* tex = nullptr;
            if ((rc = dev_upload(s, em.texels, &tex))) return fail(rc);
            dd.texels = (const float4*)tex;
            inf_distr.push_back(dd);
            const RGB spec = rgb(em.power_lookup[0], em.power_lookup[1], em.power_lookup[2]);
            if (l.type == B200PT_LIGHT_GONIOMETRIC) power = kFourPi * Lr * spec;
            else {
                power = spec * Lr * (kPi * 2.0f) * (1.0f - l.cos_total_width);
                const bool has_map = l.map_rgb && l.map_width > 0 && l.map_height > 0;
                const float aspect = has_map ? (float)l.map_width / (float)l.map_height : 1.0f;
                o.l2w[0] = 1.0f / std::tan(l.fov * (3.14159265358979323846f / 180.0f) / 2.0f);  // Transform::perspective, transform.rs:244
                if (aspect > 1.0f) { o.l2w[1] = -aspect; o.l2w[2] = -1.0f; o.l2w[3] = aspect; o.l2w[4] = 1.0f; }
                else { o.l2w[1] = -1.0f; o.l2w[2] = -1.0f / aspect; o.l2w[3] = 1.0f; o.l2w[4] = 1.0f / aspect; }
            }
        } else { b200pt_set_error("b200pt_scene_create: unknown light type"); return fail(B200PT_ERR_INVALID); }
        power_y[(size_t)i] = lum_y(power);
    }
    if ((rc = dev_upload(s, lights, &D.lights))) return fail(rc);
    if ((rc = dev_upload(s, inf_ids, &D.infinite_lights))) return fail(rc);
    if ((rc = dev_upload(s, inf_distr, &D.inf_distr))) return fail(rc);
    D.n_lights = d->n_lights;
    D.n_infinite = (int)inf_ids.size();
    // create_light_sample_distribution (light_distrib/mod.rs:59-70)
    int strat = d->n_lights == 1 ? B200PT_LIGHTS_UNIFORM : d->integrator.light_strategy;
    if (strat != B200PT_LIGHTS_UNIFORM && strat != B200PT_LIGHTS_POWER && strat != B200PT_LIGHTS_SPATIAL) {
        b200pt_set_error("b200pt_scene_create: unknown light sample strategy");
        return fail(B200PT_ERR_INVALID);
    }
    D.spatial = 0;
    if (strat == B200PT_LIGHTS_SPATIAL && d->integrator.type == B200PT_INTEGRATOR_PATH && d->n_nodes > 0) {
        // SpatialLightDistribution::new(scene, 64), spatial.rs:57-88
        const float* b = d->nodes[0].bounds;
        std::memcpy(D.wb, b, 24);
        float diag[3] = {b[3] - b[0], b[4] - b[1], b[5] - b[2]};
        int me = (diag[0] > diag[1] && diag[0] > diag[2]) ? 0 : (diag[1] > diag[2] ? 1 : 2);
        long long n_vox = 1;
        for (int i = 0; i < 3; ++i) {
            float r = std::round(diag[i] / diag[me] * 64.0f);
            long long v = (!(r == r) || r <= 0.0f) ? 0 : (long long)r;
            D.n_voxels[i] = (int)std::max<long long>(1, v);
            n_vox *= D.n_voxels[i];
        }
        // Rows (func[n_lights], cdf[n_lights + 1], func_int) are handed out on first touch from a pool: only the voxels
        // paths actually reach - a shell around the surfaces - get one, as in the reference's lazily filled hash table.
        const long long row = 2ll * d->n_lights + 2;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
        size_t pool_bytes = env_bytes("B200PT_SPATIAL_BUDGET");
        if (!pool_bytes) pool_bytes = std::min<size_t>((size_t)8 << 30, free_b / 4);
        long long pool_rows = std::min<long long>(n_vox, std::max<long long>(64, (long long)(pool_bytes / (size_t)(row * 4))));
        if ((rc = dev_alloc(s, (size_t)(pool_rows * row), &D.vox_table))) return fail(rc);
        if ((rc = dev_alloc(s, (size_t)n_vox, &D.vox_state))) return fail(rc);
        if ((rc = dev_alloc(s, (size_t)n_vox, &D.vox_row))) return fail(rc);
        if ((rc = dev_alloc(s, (size_t)n_vox, &D.vox_work))) return fail(rc);
        if ((rc = dev_alloc(s, (size_t)1, &D.vox_pool_next))) return fail(rc);
        B2_CUDA(cudaMemset(D.vox_state, 0, (size_t)n_vox * sizeof(int)));
        B2_CUDA(cudaMemset(D.vox_pool_next, 0, sizeof(int)));
        D.vox_pool_cap = (int)std::min<long long>(pool_rows, 0x7fffffff);
        D.spatial = 1;
    }
    HostDistr1D ld;
    std::vector<float> lf;
    for (int i = 0; i < d->n_lights; ++i) lf.push_back(strat == B200PT_LIGHTS_UNIFORM ? 1.0f : power_y[(size_t)i]);
    ld.init(lf);
    if ((rc = dev_upload(s, ld.func, &D.light_func))) return fail(rc);
    if ((rc = dev_upload(s, ld.cdf, &D.light_cdf))) return fail(rc);
    D.light_func_int = ld.func_int;

    // Film::get_sample_bounds (film/mod.rs:150-159)
    const b200pt_film& f = d->film;
    s->sample_bounds[0] = (int)std::floor((float)f.crop[0] + 0.5f - f.filter_radius[0]);
    s->sample_bounds[1] = (int)std::floor((float)f.crop[1] + 0.5f - f.filter_radius[1]);
    s->sample_bounds[2] = (int)std::ceil((float)f.crop[2] - 0.5f + f.filter_radius[0]);
    s->sample_bounds[3] = (int)std::ceil((float)f.crop[3] - 0.5f + f.filter_radius[1]);
    std::memcpy(D.sb, s->sample_bounds, 16);
    std::memcpy(D.pb, d->integrator.pixel_bounds, 16);
    D.max_depth = d->integrator.max_depth;
    s->whitted = d->integrator.type != B200PT_INTEGRATOR_PATH;
    s->tree_mode = d->integrator.type == B200PT_INTEGRATOR_WHITTED ? kTreeWhitted : (d->integrator.direct_strategy == B200PT_DIRECT_ONE ? kTreeDirectOne : kTreeDirectAll);
    D.rr_threshold = d->integrator.rr_threshold;

    // HaltonSampler::new over the sample bounds (samplers/src/halton.rs:61-100, 262-275)
    const b2host::HaltonTables& ht = b2host::halton_tables();
    b2host::HaltonParams hp = b2host::halton_params(D.sb[2] - D.sb[0], D.sb[3] - D.sb[1]);
    if ((rc = dev_upload(s, ht.perms, &D.halton.perms))) return fail(rc);
    if ((rc = dev_upload(s, ht.primes, &D.halton.primes))) return fail(rc);
    if ((rc = dev_upload(s, ht.prime_sums, &D.halton.prime_sums))) return fail(rc);
    if ((rc = dev_upload(s, ht.div_m, &D.halton.div_m))) return fail(rc);
    if ((rc = dev_upload(s, ht.div_sh, &D.halton.div_sh))) return fail(rc);
    for (int i = 0; i < 2; ++i) { D.halton.base_scale[i] = hp.base_scale[i]; D.halton.base_exp[i] = hp.base_exp[i]; D.halton.mult_inv[i] = hp.mult_inv[i]; }
    D.halton.stride = hp.stride;
    D.halton.sample_at_center = d->sampler.sample_at_center;
    if (d->sampler.type == B200PT_SAMPLER_SOBOL) {  // SobolSampler::new, sobol.rs:25-50
        const int ext = std::max(s->sample_bounds[2] - s->sample_bounds[0], s->sample_bounds[3] - s->sample_bounds[1]);
        uint32_t res = 1;
        while (res < (uint32_t)std::max(ext, 1)) res <<= 1;
        int m = 0;
        while ((1u << m) < res) ++m;
        std::vector<unsigned long long> vdc(52), inv(52);
        static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
        if (!b2host::sobol_interval_tables(d->sobol_matrices_32, m, (uint64_t*)vdc.data(), (uint64_t*)inv.data())) {
            b200pt_set_error("b200pt_scene_create: sobol_matrices_32 does not hold the Sobol' generator matrices (dimensions 0 / 1 are not a (0,2)-sequence)");
            return fail(B200PT_ERR_INVALID);
        }
        std::vector<uint32_t> m32(d->sobol_matrices_32, d->sobol_matrices_32 + 1024 * 52);
        if ((rc = dev_upload(s, m32, &D.sobol.m32))) return fail(rc);
        if ((rc = dev_upload(s, vdc, &D.sobol.vdc))) return fail(rc);
        if ((rc = dev_upload(s, inv, &D.sobol.vdc_inv))) return fail(rc);
        D.sobol.log2_res = m; D.sobol.res = (int)res;
        D.sobol.sb_min[0] = s->sample_bounds[0]; D.sobol.sb_min[1] = s->sample_bounds[1];
    }

    std::memcpy(D.camera.r2c, d->camera.raster_to_camera, 64);
    std::memcpy(D.camera.c2w, d->camera.camera_to_world, 64);
    D.camera.lens_radius = d->camera.lens_radius; D.camera.focal_distance = d->camera.focal_distance;
    D.camera.shutter_open = d->camera.shutter_open; D.camera.shutter_close = d->camera.shutter_close;
    if (d->camera.type != B200PT_CAMERA_PERSPECTIVE && d->camera.type != B200PT_CAMERA_ORTHOGRAPHIC && d->camera.type != B200PT_CAMERA_ENVIRONMENT) {
        b200pt_set_error("b200pt_scene_create: unknown camera type (perspective, orthographic, environment)");
        return fail(B200PT_ERR_UNSUPPORTED);
    }
    D.camera.type = d->camera.type; D.camera.xres = d->film.xres; D.camera.yres = d->film.yres;

    std::vector<float> tab(f.filter_table, f.filter_table + 256);
    const float* dt = nullptr;
    if ((rc = dev_upload(s, tab, &dt))) return fail(rc);
    s->d_filter_table = (float*)dt;
    *out = sc;
    return B200PT_OK;
}

void b200pt_scene_destroy(b200pt_scene* sc) {
    if (!sc) return;
    if (sc->impl.device >= 0) cudaSetDevice(sc->impl.device);
    for (void* p : sc->impl.allocs) cudaFree(p);
    for (void* p : sc->impl.wave_ptrs) cudaFree(p);
    for (int i = 0; i < 3; ++i) { if (sc->impl.aux[i]) cudaStreamDestroy(sc->impl.aux[i]); if (sc->impl.ev_aux[i]) cudaEventDestroy(sc->impl.ev_aux[i]); }
    if (sc->impl.ev_fork) cudaEventDestroy(sc->impl.ev_fork);
    for (auto& g : sc->impl.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (sc->impl.graph_stream) cudaStreamDestroy(sc->impl.graph_stream);
    if (sc->impl.ev_graph_in) cudaEventDestroy(sc->impl.ev_graph_in);
    if (sc->impl.ev_graph_out) cudaEventDestroy(sc->impl.ev_graph_out);
    if (sc->impl.d_sample_L) cudaFree(sc->impl.d_sample_L);
    if (sc->impl.d_sample_pf) cudaFree(sc->impl.d_sample_pf);
    if (sc->impl.d_film) cudaFree(sc->impl.d_film);
    if (sc->impl.d_acc) cudaFree(sc->impl.d_acc);
    if (sc->impl.d_totals) cudaFree(sc->impl.d_totals);
    if (sc->impl.h_pinned) cudaFreeHost(sc->impl.h_pinned);
    if (sc->impl.ev_shade) cudaEventDestroy(sc->impl.ev_shade);
    for (void* p : {(void*)sc->impl.ztq.rng, (void*)sc->impl.ztq.tile, (void*)sc->impl.ztq.cursor, (void*)sc->impl.ztq.samp, (void*)sc->impl.ztq.dst, (void*)sc->impl.ztq.scr1,
                    (void*)sc->impl.ztq.scr2, (void*)sc->impl.ztq.perm1, (void*)sc->impl.ztq.perm2})
        if (p) cudaFree(p);
    for (void* p : {(void*)sc->impl.d_zt_scr1, (void*)sc->impl.d_zt_scr2, (void*)sc->impl.d_zt_perm1, (void*)sc->impl.d_zt_perm2, (void*)sc->impl.d_zt_scratch})
        if (p) cudaFree(p);
    if (sc->impl.d_rows) cudaFree(sc->impl.d_rows);
    if (sc->impl.d_row_index) cudaFree(sc->impl.d_row_index);
    alpha_free_device(&sc->impl.alpha);
    accel_free_device(&sc->impl.accel);
    accel2_free_device(&sc->impl.accel2);
    delete sc;
}

int b200pt_scene_ray_counts(const b200pt_scene* s, uint64_t counts[3]) {
    if (!s || !counts) { b200pt_set_error("b200pt_scene_ray_counts: null argument"); return B200PT_ERR_INVALID; }
    counts[0] = s->impl.rays[0]; counts[1] = s->impl.rays[1]; counts[2] = s->impl.rays[2];
    return B200PT_OK;
}

// Builds the (0,2)-sequence tables for the rows currently described by s->d_row_index (no-op for Halton).
static int zerotwo_prepare(SceneImpl* s, long long n_pix, cudaStream_t st) {
    if (s->dev.sampler_type != B200PT_SAMPLER_ZEROTWO || s->zt_seq) return B200PT_OK;
    const int* sb = s->sample_bounds;
    const int spp = s->spp, dims = s->sampler.dimensions;
    const int n1 = 1 + 2 * s->dev.max_depth, n2 = 2 + 3 * s->dev.max_depth;
    const int ntx = (sb[2] - sb[0] + 15) / 16, nty = (sb[3] - sb[1] + 15) / 16;
    if (n_pix > s->zt_pix_cap) {
        for (void** p : {(void**)&s->d_zt_scr1, (void**)&s->d_zt_scr2, (void**)&s->d_zt_perm1, (void**)&s->d_zt_perm2}) { if (*p) cudaFree(*p); *p = nullptr; }
        s->zt_pix_cap = 0;
        B2_CUDA(cudaMalloc(&s->d_zt_scr1, (size_t)n_pix * n1 * sizeof(uint32_t)));
        B2_CUDA(cudaMalloc(&s->d_zt_scr2, (size_t)n_pix * n2 * 2 * sizeof(uint32_t)));
        B2_CUDA(cudaMalloc(&s->d_zt_perm1, (size_t)n_pix * n1 * spp * sizeof(uint16_t)));
        B2_CUDA(cudaMalloc(&s->d_zt_perm2, (size_t)n_pix * n2 * spp * sizeof(uint16_t)));
        s->zt_pix_cap = n_pix;
    }
    if (!s->d_zt_scratch) B2_CUDA(cudaMalloc(&s->d_zt_scratch, (size_t)ntx * nty * spp * sizeof(uint16_t)));
    k_zerotwo_tiles<<<(ntx * nty + 63) / 64, 64, 0, st>>>(sb[0], sb[1], sb[2], sb[3], ntx, nty, dims, n1, n2, spp, s->d_row_index, s->d_zt_scr1,
                                                          s->d_zt_perm1, s->d_zt_scr2, s->d_zt_perm2, s->d_zt_scratch);
    g_launches.fetch_add(1);
    s->dev.zt.scr1 = s->d_zt_scr1; s->dev.zt.perm1 = s->d_zt_perm1; s->dev.zt.scr2 = s->d_zt_scr2; s->dev.zt.perm2 = s->d_zt_perm2;
    s->dev.zt.n1 = n1; s->dev.zt.n2 = n2; s->dev.zt.spp = spp;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_zerotwo_tiles");
    return B200PT_OK;
}

// Per-slot state of the tile-sequential (0,2) mode for `slots` tiles (or list entries) in flight.
static int ztq_alloc(SceneImpl* s, int slots) {
    ZtSeq& Z = s->ztq;
    const int dims = std::max(1, s->sampler.dimensions), spp = s->spp;
    if (slots > s->ztq_cap) {
        for (void** p : {(void**)&Z.rng, (void**)&Z.tile, (void**)&Z.cursor, (void**)&Z.samp, (void**)&Z.dst, (void**)&Z.scr1, (void**)&Z.scr2, (void**)&Z.perm1, (void**)&Z.perm2}) {
            if (*p) cudaFree(*p);
            *p = nullptr;
        }
        s->ztq_cap = 0;
        B2_CUDA(cudaMalloc(&Z.rng, (size_t)slots * sizeof(DPcg32)));
        B2_CUDA(cudaMalloc(&Z.tile, (size_t)slots * sizeof(int)));
        B2_CUDA(cudaMalloc(&Z.cursor, (size_t)slots * sizeof(int)));
        B2_CUDA(cudaMalloc(&Z.samp, (size_t)slots * sizeof(int)));
        B2_CUDA(cudaMalloc(&Z.dst, (size_t)slots * sizeof(int)));
        B2_CUDA(cudaMalloc(&Z.scr1, (size_t)slots * dims * sizeof(uint32_t)));
        B2_CUDA(cudaMalloc(&Z.scr2, (size_t)slots * dims * 2 * sizeof(uint32_t)));
        B2_CUDA(cudaMalloc(&Z.perm1, (size_t)slots * dims * spp * sizeof(uint16_t)));
        B2_CUDA(cudaMalloc(&Z.perm2, (size_t)slots * dims * spp * sizeof(uint16_t)));
        s->ztq_cap = slots;
    }
    const int* sb = s->sample_bounds;
    Z.dims = s->sampler.dimensions; Z.spp = spp; Z.ntx = (sb[2] - sb[0] + 15) / 16;
    DZeroTwo& z = s->dev.zt;
    z.scr1 = Z.scr1; z.perm1 = Z.perm1; z.scr2 = Z.scr2; z.perm2 = Z.perm2;
    z.n1 = z.n2 = Z.dims; z.spp = spp; z.rng = Z.rng;
    return B200PT_OK;
}

// The wave loop of the tile-sequential (0,2) mode: the shard's tile rows are taken in groups whose samples fit the sample
// store; a group's tiles each carry one path per step, 256 x spp steps (fewer for clipped tiles), then k_film continues
// the film sums with the group's samples exactly as after a wave of the parallel modes.
static int render_zt_seq(SceneImpl* s, const std::vector<int>& srows, cudaStream_t st, const std::function<void(long long, long long, int, int)>& film_pass) {
    const int* sb = s->sample_bounds;
    const int sw = sb[2] - sb[0], spp = s->spp, ntx = (sw + 15) / 16;
    struct TRow { int ty; size_t k0, k1; };
    std::vector<TRow> trows;
    for (size_t k = 0; k < srows.size();) {
        const int ty = (srows[k] - sb[1]) / 16;
        size_t k1 = k;
        while (k1 < srows.size() && (srows[k1] - sb[1]) / 16 == ty) ++k1;
        trows.push_back({ty, k, k1});
        k = k1;
    }
    int rc = B200PT_OK;
    std::vector<int> tiles;
    for (size_t i = 0; i < trows.size() && !rc;) {
        size_t j = i;
        long long ns = 0;
        int max_h = 0;
        while (j < trows.size()) {
            const long long add = (long long)(trows[j].k1 - trows[j].k0) * sw * spp;
            if (j > i && (ns + add > s->wave_cap || (long long)(j - i + 1) * ntx > s->wave_cap)) break;
            ns += add;
            max_h = std::max(max_h, std::min(16, sb[3] - (sb[1] + trows[j].ty * 16)));
            ++j;
        }
        if (ns > s->wave_cap || ntx > s->wave_cap) {
            b200pt_set_error("render: the memory budget does not hold the samples of one row of 16x16 tiles (02sequence with \"dimensions\" below the path's worst case renders tile-sequentially)");
            return B200PT_ERR_OOM;
        }
        tiles.clear();
        for (size_t t = i; t < j; ++t)
            for (int tx = 0; tx < ntx; ++tx) tiles.push_back(trows[t].ty * ntx + tx);
        const int T = (int)tiles.size();
        if ((rc = ztq_alloc(s, T))) return rc;
        B2_CUDA(cudaMemcpyAsync(s->ztq.tile, tiles.data(), (size_t)T * sizeof(int), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemsetAsync(s->d_sample_L, 0, (size_t)ns * sizeof(float4), st));   // weight 0: pixels outside the integrator's pixel bounds
        B2_CUDA(cudaMemsetAsync(s->d_sample_pf, 0, (size_t)ns * sizeof(float2), st));
        B2_CUDA(cudaStreamSynchronize(st));  // `tiles` is reused by the next group
        const long long first_local = (long long)trows[i].k0 * sw * spp;
        const long long steps = (long long)std::min(16, sw) * max_h * spp;
        for (long long step = 0; step < steps && !rc; ++step) {
            k_zt_seq_next<<<(T + 63) / 64, 64, 0, st>>>(s->dev, s->wave, s->ztq, T, step == 0 ? 1 : 0, s->d_row_index, first_local, s->d_sample_pf, s->d_totals);
            g_launches.fetch_add(1);
            rc = run_wave(s, T, st);
            if (rc) break;
            k_zt_seq_store<<<(T + 255) / 256, 256, 0, st>>>(s->wave, s->ztq.dst, T, s->d_sample_L);
            g_launches.fetch_add(1);
        }
        if (!rc) film_pass(first_local, ns, srows[trows[i].k0], srows[trows[j - 1].k1 - 1]);
        i = j;
    }
    return rc;
}

// Reads the render's ray totals and error flags back (the one read-back of a render besides the film).
static int finish_totals(SceneImpl* s, cudaStream_t st) {
    unsigned long long t[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(s->h_pinned, s->d_totals, sizeof(t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "render");
    std::memcpy(t, s->h_pinned, sizeof(t));
    s->rays[0] += t[0]; s->rays[1] += t[1]; s->rays[2] += t[2];
    if (s->dev.spatial) {
        int built = 0;
        B2_CUDA(cudaMemcpy(&built, s->dev.vox_pool_next, sizeof(int), cudaMemcpyDeviceToHost));
        s->voxels_built = (uint64_t)std::min(built, s->dev.vox_pool_cap);
    }
    if (t[3] & 1ull) {
        b200pt_set_error("render: the spatial light distribution touched more voxels than its row pool holds (raise B200PT_SPATIAL_BUDGET, or use lightsamplestrategy power / uniform)");
        return B200PT_ERR_OOM;
    }
    return B200PT_OK;
}

// Renders the sample rows listed in `srows` (ascending, inside the sample bounds) into a zero-initialised film of
// the full cropped window.  Shards with disjoint row sets sum to the whole image (each sample is taken once).
// The shard's samples (pixel-major) are cut into waves that fit the memory budget; after each wave k_film continues the
// running sums of the film pixels the wave's samples reach, so memory does not grow with resolution x samples per pixel.
static int render_rows_impl(SceneImpl* s, const std::vector<int>& srows, void* d_film_xyzw, cudaStream_t st, bool raw = false) {
    int rc = B200PT_OK;
    const b200pt_film& f = s->film;
    const int cw = f.crop[2] - f.crop[0], ch = f.crop[3] - f.crop[1];
    const long long n_film = (long long)cw * ch;
    s->rays[0] = s->rays[1] = s->rays[2] = 0;
    if (srows.empty() || n_film <= 0) {
        if (n_film > 0) B2_CUDA(cudaMemsetAsync(d_film_xyzw, 0, (size_t)n_film * sizeof(float4), st));
        B2_CUDA(cudaStreamSynchronize(st));
        return B200PT_OK;
    }
    const int* sb = s->sample_bounds;
    const int sw = sb[2] - sb[0], sh = sb[3] - sb[1], spp = s->spp;
    const long long n_samples = (long long)srows.size() * sw * spp;
    if ((rc = wave_alloc(s, wave_cap_for(s, n_samples)))) return rc;
    if (s->wave_cap > s->sample_cap) {
        if (s->d_sample_L) cudaFree(s->d_sample_L);
        if (s->d_sample_pf) cudaFree(s->d_sample_pf);
        s->d_sample_L = nullptr; s->d_sample_pf = nullptr; s->sample_cap = 0;
        B2_CUDA(cudaMalloc(&s->d_sample_L, (size_t)s->wave_cap * sizeof(float4)));
        B2_CUDA(cudaMalloc(&s->d_sample_pf, (size_t)s->wave_cap * sizeof(float2)));
        s->sample_cap = s->wave_cap;
    }
    if ((size_t)n_film * sizeof(float4) > s->acc_cap) {
        if (s->d_acc) cudaFree(s->d_acc);
        s->d_acc = nullptr; s->acc_cap = 0;
        B2_CUDA(cudaMalloc(&s->d_acc, (size_t)n_film * sizeof(float4)));
        s->acc_cap = (size_t)n_film * sizeof(float4);
    }
    if (!s->d_rows) {
        B2_CUDA(cudaMalloc(&s->d_rows, (size_t)sh * sizeof(int)));
        B2_CUDA(cudaMalloc(&s->d_row_index, (size_t)sh * sizeof(int)));
    }
    std::vector<int> row_index((size_t)sh, -1);
    for (size_t k = 0; k < srows.size(); ++k) row_index[(size_t)(srows[k] - sb[1])] = (int)k;
    B2_CUDA(cudaMemcpyAsync(s->d_rows, srows.data(), srows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(s->d_row_index, row_index.data(), (size_t)sh * sizeof(int), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemsetAsync(s->d_acc, 0, (size_t)n_film * sizeof(float4), st));
    B2_CUDA(cudaMemsetAsync(s->d_totals, 0, 4 * sizeof(unsigned long long), st));
    B2_CUDA(cudaStreamSynchronize(st));  // the host vectors above go out of scope
    if ((rc = zerotwo_prepare(s, (long long)srows.size() * sw, st))) return rc;
    DFilm F;
    std::memcpy(F.crop, f.crop, 16);
    F.rx = f.filter_radius[0]; F.ry = f.filter_radius[1];
    F.inv_rx = 1.0f / F.rx; F.inv_ry = 1.0f / F.ry;
    F.max_lum = f.max_sample_luminance;
    std::memcpy(F.sb, sb, 16);
    F.tile = 16;
    const int reach = (int)std::ceil(F.ry) + 2;  // film rows a sample row can touch, generously
    float4* d_L = s->d_sample_L;
    float2* d_pf = s->d_sample_pf;
    if (s->zt_seq) {
        rc = render_zt_seq(s, srows, st, [&](long long first, long long n, int y_lo, int y_hi) {
            const int y0 = std::max(f.crop[1], y_lo - reach), y1 = std::min(f.crop[3], y_hi + reach + 1);
            if (y1 <= y0) return;
            const long long npix = (long long)cw * (y1 - y0);
            k_film<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(F, s->d_filter_table, d_L, d_pf, first, (int)n, spp, s->d_row_index, y0, y1, s->d_acc);
            g_launches.fetch_add(1);
        });
    }
    for (long long first = 0; first < n_samples && !rc && !s->zt_seq; first += s->wave_cap) {
        int n = (int)std::min<long long>(s->wave_cap, n_samples - first);
        k_raygen<<<(n + 255) / 256, 256, 0, st>>>(s->dev, s->wave, first, n, spp, s->d_rows, nullptr, d_pf, nullptr);
        g_launches.fetch_add(1);
        rc = s->whitted ? run_wave_whitted(s, n, st) : run_wave_graphed(s, n, st);
        if (rc) break;
        k_store_samples<<<(n + 255) / 256, 256, 0, st>>>(s->wave, n, d_L);
        // every film pixel near the wave's sample rows continues its sums (also pixels of a neighbouring shard when the
        // filter is wider than a pixel; the shards' films are summed afterwards)
        const int y_lo = srows[(size_t)(first / ((long long)sw * spp))], y_hi = srows[(size_t)((first + n - 1) / ((long long)sw * spp))];
        const int y0 = std::max(f.crop[1], y_lo - reach), y1 = std::min(f.crop[3], y_hi + reach + 1);
        if (y1 > y0) {
            const long long npix = (long long)cw * (y1 - y0);
            k_film<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(F, s->d_filter_table, d_L, d_pf, first, n, spp, s->d_row_index, y0, y1, s->d_acc);
        }
        g_launches.fetch_add(2);
    }
    if (!rc) {
        // raw: the running sums themselves {sum of filter-weighted RGB, sum of filter weights} (shard films are combined in
        // that space, b200pt_film_finish_device converts afterwards); otherwise Film::merge_film_tile's RGB -> XYZ
        if (raw) B2_CUDA(cudaMemcpyAsync(d_film_xyzw, s->d_acc, (size_t)n_film * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        else { k_film_finish<<<(unsigned)((n_film + 255) / 256), 256, 0, st>>>(s->d_acc, n_film, (float4*)d_film_xyzw); g_launches.fetch_add(1); }
        rc = finish_totals(s, st);
    }
    return rc;
}

// sample rows of the pixel rows [r0, r1) of the cropped window; the rows of the sample bounds above / below the
// window (filters wider than a pixel) go with the first / last pixel row
static void append_rows(const SceneImpl* s, int r0, int r1, std::vector<int>* out) {
    const b200pt_film& f = s->film;
    const int ch = f.crop[3] - f.crop[1];
    const int* sb = s->sample_bounds;
    int a = r0 == 0 ? sb[1] : f.crop[1] + r0, b = r1 == ch ? sb[3] : f.crop[1] + r1;
    for (int y = a; y < b; ++y) out->push_back(y);
}

int b200pt_render_rows_device(b200pt_scene* sc, int32_t row_begin, int32_t row_end, void* d_film_xyzw, void* stream) {
    if (!sc || !d_film_xyzw) { b200pt_set_error("b200pt_render_rows_device: null argument"); return B200PT_ERR_INVALID; }
    SceneImpl* s = &sc->impl;
    std::lock_guard<std::mutex> g(s->mu);  // render serialises per scene
    int rc = use_device(s->device);
    if (rc) return rc;
    const int ch = s->film.crop[3] - s->film.crop[1];
    if (row_begin < 0 || row_end > ch || row_begin > row_end) { b200pt_set_error("b200pt_render_rows_device: row range outside the cropped window"); return B200PT_ERR_INVALID; }
    std::vector<int> rows;
    if (row_begin < row_end) append_rows(s, row_begin, row_end, &rows);
    return render_rows_impl(s, rows, d_film_xyzw, (cudaStream_t)stream);
}

int32_t b200pt_band_owner(int32_t band, int32_t n_shards) {
    if (n_shards < 1 || band < 0) return 0;
    static const bool round_robin = [] { const char* e = std::getenv("B200PT_BAND_ORDER"); return e && std::strcmp(e, "roundrobin") == 0; }();  // A/B
    if (round_robin) return band % n_shards;
    const int32_t pass = band / n_shards, k = band % n_shards;
    return (pass & 1) ? n_shards - 1 - k : k;
}
static int render_shard_impl(b200pt_scene* sc, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_xyzw, void* stream, bool raw);
int b200pt_render_shard_device(b200pt_scene* sc, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_xyzw, void* stream) {
    return render_shard_impl(sc, shard, n_shards, band_rows, d_film_xyzw, stream, false);
}
int b200pt_render_shard_device_raw(b200pt_scene* sc, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_rgbw, void* stream) {
    return render_shard_impl(sc, shard, n_shards, band_rows, d_film_rgbw, stream, true);
}
int b200pt_film_finish_device(const void* d_film_rgbw, int64_t n_pix, void* d_film_xyzw, void* stream) {
    if (!d_film_rgbw || !d_film_xyzw || n_pix < 0) { b200pt_set_error("b200pt_film_finish_device: invalid argument"); return B200PT_ERR_INVALID; }
    if (n_pix == 0) return B200PT_OK;
    k_film_finish<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_film_rgbw, n_pix, (float4*)d_film_xyzw);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? B200PT_OK : cuda_fail(e, "k_film_finish");
}
static int render_shard_impl(b200pt_scene* sc, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_xyzw, void* stream, bool raw) {
    if (!sc || !d_film_xyzw || n_shards < 1 || shard < 0 || shard >= n_shards || band_rows < 1) { b200pt_set_error("b200pt_render_shard_device: invalid argument"); return B200PT_ERR_INVALID; }
    SceneImpl* s = &sc->impl;
    std::lock_guard<std::mutex> g(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    if (s->zt_seq && n_shards > 1 && (band_rows % 16 != 0 || (s->film.crop[1] - s->sample_bounds[1]) % 16 != 0)) {
        // a tile is one sequential stream: a shard that owns part of a tile's rows has to render the whole tile anyway
        b200pt_set_error("b200pt_render_shard_device: the tile-sequential (0,2) mode shards by whole 16-row tiles: band_rows must be a multiple of 16 (and the crop window must start on a tile row)");
        return B200PT_ERR_INVALID;
    }
    const int ch = s->film.crop[3] - s->film.crop[1];
    std::vector<int> rows;
    for (int r0 = 0, band = 0; r0 < ch; r0 += band_rows, ++band)
        if (b200pt_band_owner(band, n_shards) == shard) append_rows(s, r0, std::min(ch, r0 + band_rows), &rows);
    return render_rows_impl(s, rows, d_film_xyzw, (cudaStream_t)stream, raw);
}

int b200pt_render_rows(b200pt_scene* sc, int32_t row_begin, int32_t row_end, float* film_xyzw) {
    if (!sc || !film_xyzw) { b200pt_set_error("b200pt_render_rows: null argument"); return B200PT_ERR_INVALID; }
    int rc = use_device(sc->impl.device);
    if (rc) return rc;
    const b200pt_film& f = sc->impl.film;
    size_t bytes = (size_t)(f.crop[2] - f.crop[0]) * (f.crop[3] - f.crop[1]) * sizeof(float4);
    SceneImpl* s = &sc->impl;
    if (bytes > s->film_cap) {
        if (s->d_film) cudaFree(s->d_film);
        s->d_film = nullptr; s->film_cap = 0;
        B2_CUDA(cudaMalloc(&s->d_film, std::max<size_t>(bytes, 16)));
        s->film_cap = bytes;
    }
    rc = b200pt_render_rows_device(sc, row_begin, row_end, s->d_film, nullptr);
    if (!rc) {
        cudaError_t e = cudaMemcpy(film_xyzw, s->d_film, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "film download");
    }
    return rc;
}

// Film::write_image / get_pixel_rgb (film/mod.rs:356-417), no splats on this path.
int b200pt_film_resolve(const b200pt_film* f, const float* film_xyzw, float* rgb_out) {
    if (!f || !film_xyzw || !rgb_out) { b200pt_set_error("b200pt_film_resolve: null argument"); return B200PT_ERR_INVALID; }
    size_t n = (size_t)(f->crop[2] - f->crop[0]) * (f->crop[3] - f->crop[1]);
    for (size_t i = 0; i < n; ++i) {
        const float* p = film_xyzw + 4 * i;
        float c[3];
        c[0] = 3.240479f * p[0] - 1.537150f * p[1] - 0.498535f * p[2];
        c[1] = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
        c[2] = 0.055648f * p[0] - 0.204043f * p[1] + 1.057311f * p[2];
        for (int k = 0; k < 3; ++k) {
            float v = c[k];
            if (p[3] != 0.0f) { float inv = 1.0f / p[3]; v = (v * inv) > 0.0f ? (v * inv) : 0.0f; }
            v *= f->scale;
            rgb_out[3 * i + k] = v;
        }
    }
    return B200PT_OK;
}

int b200pt_li_batch(b200pt_scene* sc, const int32_t* pixel_sample, int64_t n, float* li_out, b200pt_ray* rays_out) {
    if (!sc || n < 0 || (n > 0 && (!pixel_sample || !li_out))) { b200pt_set_error("b200pt_li_batch: invalid argument"); return B200PT_ERR_INVALID; }
    SceneImpl* s = &sc->impl;
    std::lock_guard<std::mutex> g(s->mu);
    int rc = use_device(s->device);
    if (rc) return rc;
    if ((rc = wave_alloc(s, wave_cap_for(s, n)))) return rc;
    B2_CUDA(cudaMemset(s->d_totals, 0, 4 * sizeof(unsigned long long)));
    if (s->dev.sampler_type == B200PT_SAMPLER_ZEROTWO) {  // explicit lists may name any pixel: own every sample row
        const int* sb = s->sample_bounds;
        const int sw = sb[2] - sb[0], sh = sb[3] - sb[1];
        for (int64_t i = 0; i < n; ++i) {
            const int32_t* e = pixel_sample + 3 * i;
            if (e[0] < sb[0] || e[0] >= sb[2] || e[1] < sb[1] || e[1] >= sb[3] || e[2] < 0 || e[2] >= s->spp) {
                b200pt_set_error("b200pt_li_batch: (pixel, sample) outside the sample bounds / sample count");
                return B200PT_ERR_INVALID;
            }
        }
        if (!s->d_rows) {
            B2_CUDA(cudaMalloc(&s->d_rows, (size_t)sh * sizeof(int)));
            B2_CUDA(cudaMalloc(&s->d_row_index, (size_t)sh * sizeof(int)));
        }
        std::vector<int> ident((size_t)sh);
        for (int k = 0; k < sh; ++k) ident[(size_t)k] = k;
        B2_CUDA(cudaMemcpy(s->d_row_index, ident.data(), (size_t)sh * sizeof(int), cudaMemcpyHostToDevice));
        if ((rc = zerotwo_prepare(s, (long long)sh * sw, 0))) return rc;
    }
    int* d_list = nullptr;
    float4 *d_L = nullptr, *d_rays = nullptr;
    const int cap = (int)std::min<int64_t>(s->wave_cap, std::max<int64_t>(n, 1));
    if (s->zt_seq && (rc = ztq_alloc(s, cap))) return rc;
    B2_CUDA(cudaMalloc(&d_list, (size_t)cap * 3 * sizeof(int)));
    cudaMalloc(&d_L, (size_t)cap * sizeof(float4));
    cudaMalloc(&d_rays, (size_t)cap * 2 * sizeof(float4));
    std::vector<float4> hL;
    for (int64_t first = 0; first < n && !rc; first += cap) {
        int m = (int)std::min<int64_t>(cap, n - first);
        cudaMemcpy(d_list, pixel_sample + 3 * first, (size_t)m * 3 * sizeof(int), cudaMemcpyHostToDevice);
        if (s->zt_seq) k_zt_seq_list<<<(m + 63) / 64, 64>>>(s->dev, s->wave, s->ztq, m, d_list, d_rays);
        else k_raygen<<<(m + 255) / 256, 256>>>(s->dev, s->wave, 0, m, s->spp, nullptr, d_list, nullptr, d_rays);
        g_launches.fetch_add(1);
        rc = s->whitted ? run_wave_whitted(s, m, 0) : run_wave(s, m, 0);
        if (rc) break;
        k_store_samples<<<(m + 255) / 256, 256>>>(s->wave, m, d_L);
        g_launches.fetch_add(1);
        hL.resize((size_t)m);
        cudaError_t e = cudaMemcpy(hL.data(), d_L, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = cuda_fail(e, "li download"); break; }
        for (int i = 0; i < m; ++i) { li_out[3 * (first + i)] = hL[i].x; li_out[3 * (first + i) + 1] = hL[i].y; li_out[3 * (first + i) + 2] = hL[i].z; }
        if (rays_out) cudaMemcpy(rays_out + first, d_rays, (size_t)m * sizeof(b200pt_ray), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_list); cudaFree(d_L); cudaFree(d_rays);
    if (!rc) rc = finish_totals(s, 0);
    return rc;
}

int b200pt_scene_set_memory_budget(b200pt_scene* sc, uint64_t bytes) {
    if (!sc) { b200pt_set_error("b200pt_scene_set_memory_budget: null scene"); return B200PT_ERR_INVALID; }
    std::lock_guard<std::mutex> g(sc->impl.mu);
    sc->impl.mem_budget = (size_t)bytes;
    return B200PT_OK;
}

}  // extern "C"
