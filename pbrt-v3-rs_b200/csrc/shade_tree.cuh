// Recursive SamplerIntegrators (Whitted, DirectLighting) as wavefront stages.  Internal header of b200pt_scene.cu.
#pragma once
#include "wavefront.cuh"

namespace b2 {

// ---- K4w: WhittedIntegrator::li (integrators/src/whitted.rs:60-126) and DirectLightingIntegrator::li
// (integrators/src/direct_lighting.rs:82-146) as a wavefront stage ---------------------------------------------------
// Both recurse: reflect subtree, then transmit subtree, drawing sampler dimensions in that depth-first order.
// Here every path walks its own tree depth-first, one node per wave iteration: the specular-transmission child of a
// node is computed at the node (delta lobes ignore the sample value) and parked on the path's stack while the
// reflection subtree runs; popping it consumes the two dimensions specular_transmit's get_2d() would have drawn at that
// point, so every later light sample sees the reference's dimension.  A node's own radiance l = Le + direct light is
// formed in the reference's order by k_resolve_tree and enters the pixel as L += beta * l with beta the product of
// f * |wi . ns| / pdf down the tree (the reference multiplies on the way back up: same value up to f32 rounding,
// identical for depth-0 nodes).
//
// Direct light per node, kMode:
//   kTreeWhitted    every light once: f * Li * |wi . ns| / pdf if unoccluded (whitted.rs:89-112)
//   kTreeDirectAll  uniform_sample_all_lights (integrator/common.rs:25-87).  The tile samplers are made by
//                   clone_sampler(), which drops the sample arrays the integrator requested in preprocess()
//                   (halton.rs:176-182), so get_2d_array() is always empty and every light takes the single-sample
//                   branch: u_light = get_2d(), u_scattering = get_2d(), one estimate_direct with MIS
//   kTreeDirectOne  uniform_sample_one_light with no distribution: light = min(u * n, n - 1), estimate / (1 / n)
// Pending record k owns the slots [k * stride, (k + 1) * stride) of the shadow and MIS queues (stride = number of
// lights, or 1 for kTreeDirectOne); unused slots carry t_max = -1 rays.
enum { kTreeWhitted = 0, kTreeDirectAll = 1, kTreeDirectOne = 2 };

B2_D void wstack_store(const Wave& W, int max_depth, int pid, int sp, V3 o, V3 d, float time, RGB beta, int depth, bool valid) {
    float4* e = W.wstack + ((long long)pid * max_depth + sp) * W.wstack_n;
    e[0] = make_float4(o.x, o.y, o.z, time);
    e[1] = make_float4(d.x, d.y, d.z, __int_as_float(valid ? depth : -1));
    e[2] = make_float4(beta.r, beta.g, beta.b, 0.0f);
}
B2_D void tree_slot_clear(const Wave& W, long long sl, bool with_mis) {
    W.sh_c[sl] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    store_ray(W.sh_ray, (int)sl, mk(0, 0, 0), mk(0, 0, 1), -1.0f, 0.0f);
    if (with_mis) store_ray(W.mis_ray, (int)sl, mk(0, 0, 0), mk(0, 0, 1), -1.0f, 0.0f);
}

template <int kMode>
__global__ void __launch_bounds__(128) k_shade_tree(DeviceScene S, Wave W, int cur, int n_active) {
    int i_sorted = blockIdx.x * blockDim.x + threadIdx.x;
    if (i_sorted >= n_active) return;
    const int slot = W.sorted[i_sorted];
    const int pid = W.qpid[cur][slot];
    const float4 r1 = W.ray[cur][2 * slot + 1];
    const float4 hit = W.hit[slot];
    const float hb2 = W.hit_b2[slot];
    const V3 ray_d = mk(r1.x, r1.y, r1.z);
    const float time = r1.w;
    int meta = W.meta[pid];
    int dim = meta & 0xffff, depth = (meta >> 16) & 0xff, sp = (meta >> 24) & 0xff;
    if (depth == 255) return;  // pixel outside the integrator's pixel bounds: no sample is taken
    const float4 Lw = W.L[pid], bw = W.beta[pid];
    RGB L = rgb(Lw.x, Lw.y, Lw.z);
    const RGB beta = rgb(bw.x, bw.y, bw.z);
    const unsigned long long hidx = W.hidx[pid];
    const uint32_t prim = __float_as_uint(hit.y);
    const bool found = prim != 0xffffffffu;
    const int stride = kMode == kTreeDirectOne ? 1 : S.n_lights;

    bool have_next = false;
    V3 next_o = mk(0, 0, 0), next_d = next_o;
    RGB next_beta = beta;
    int next_depth = 0;

    if (!found) {  // whitted.rs:117-121, direct_lighting.rs:138-143
        RGB l = rgb1(0.0f);
        for (int i = 0; i < S.n_infinite; ++i) {
            const DLight& il = S.lights[S.infinite_lights[i]];
            l = l + infinite_le(il, S.inf_distr[il.inf_slot], ray_d);
        }
        L = L + beta * l;
    } else {
        const int dims_here = kMode == kTreeWhitted ? 2 * S.n_lights : (kMode == kTreeDirectAll ? 4 * S.n_lights : 5);
        if (dim + dims_here + 4 > (S.sampler_type == B200PT_SAMPLER_SOBOL ? 1024 : 1000)) {  // HaltonSampler asserts at 1000 dimensions (halton.rs:106-110), Sobol at 1024
            atomicExch(&W.counters[4], 1);
            return;
        }
        HitCtx hc;
        surface_at(S, W, slot, prim, hit, hb2, ray_d, true, &hc);
        const SurfHit& sh = hc.sh;
        const V3 wo = hc.wo;
        BSDF bsdf;
        bsdf.ns = sh.ns; bsdf.ng = sh.n;
        bsdf.ss = normalize(sh.dpdu);
        bsdf.ts = cross(bsdf.ns, bsdf.ss);
        bsdf.m = S.materials + hc.mat;  // built with allow_multiple_lobes = false (whitted.rs:76, direct_lighting.rs:91)
        // Ray differentials (scenes with a closed-form checkerboard): every ray of the tree carries them - the camera ray's, then
        // the ones specular_reflect / specular_transmit derive for their children - in W.cam_diff[pid]; si.der of this hit:
        const bool have_diff = W.cam_diff != nullptr;
        HitDerivs HD;
        HD.dudx = HD.dvdx = HD.dudy = HD.dvdy = 0.0f;
        HD.dpdx = HD.dpdy = mk(0.0f, 0.0f, 0.0f);
        V3 dndu = mk(0.0f, 0.0f, 0.0f), dndv = dndu;
        float tex_u = 0.0f, tex_v = 0.0f;
        float4 parent_diff[3];
        if (have_diff) {
            for (int k = 0; k < 3; ++k) parent_diff[k] = W.cam_diff[3ll * pid + k];
            tree_hit_derivs(S, W, slot, prim, hit, hb2, sh, parent_diff, &HD, &tex_u, &tex_v, &dndu, &dndv);
        }
        DMaterial tex_mat;
        if (S.mat_kd_tex) {  // textured Kd (matte.rs:63, plastic.rs:81)
            const int tex = S.mat_kd_tex[hc.mat];
            if (tex >= 0) {
                if (have_diff) {
                    RGB kd = spectrum_texture_eval(S.spec_tex[tex], tex_u, tex_v, HD.dudx, HD.dvdx, HD.dudy, HD.dvdy);
                    kd = rgb(clamp0inf(kd.r), clamp0inf(kd.g), clamp0inf(kd.b));
                    tex_mat = S.materials[hc.mat];
                    if (is_black(kd)) { tex_mat.n_bxdf -= 1; tex_mat.bx[0] = tex_mat.bx[1]; }
                    else { tex_mat.bx[0].r[0] = kd.r; tex_mat.bx[0].r[1] = kd.g; tex_mat.bx[0].r[2] = kd.b; }
                } else textured_material(S, W, slot, prim, hit, hb2, sh, tex, nullptr, S.materials[hc.mat], &tex_mat);
                bsdf.m = &tex_mat;
            }
        }
        RGB le = rgb1(0.0f);
        if (hc.alight >= 0) le = area_l(S.lights[hc.alight], sh.n, wo);  // isect.le(&wo)
        int rec = -1, n_real_sh = 0, n_real_mis = 0;
        const int n_est = kMode == kTreeDirectOne ? (S.n_lights > 0 ? 1 : 0) : S.n_lights;
        for (int e = 0; e < n_est; ++e) {
            int li = e;
            if (kMode == kTreeDirectOne) {  // common.rs:111-116
                float u = smp_1d(S, hidx, dim);
                float fn = (float)S.n_lights;
                li = (int)pmin(u * fn, fn - 1.0f);
            }
            const DLight& light = S.lights[li];
            bool want_sh = false, want_mis = false;
            RGB c = rgb1(0.0f);
            DirectEst de;
            if (kMode == kTreeWhitted) {
                P2 u = smp_2d(S, hidx, dim);
                const LightSample ls = sample_light(S, light, sh, u);
                de.shadow = false; de.mis = false;
                if (ls.valid && !is_black(ls.Li) && ls.pdf != 0.0f) {
                    RGB f = bsdf_f(bsdf, wo, ls.wi, BSDF_ALL);
                    if (!is_black(f)) {
                        want_sh = true;
                        c = f * ls.Li * abs_dot(ls.wi, sh.ns) / ls.pdf;
                        de.sh_o = offset_ray_origin(sh.p, sh.p_error, sh.n, ls.p1 - sh.p);  // Hit::spawn_ray_to_hit
                        V3 target = offset_ray_origin(ls.p1, ls.p1_err, ls.p1_n, de.sh_o - ls.p1);
                        de.sh_d = target - de.sh_o;
                    }
                }
            } else {
                P2 u_light = smp_2d(S, hidx, dim);
                P2 u_scatter = smp_2d(S, hidx, dim);
                de = estimate_direct_rays(S, light, sh, wo, bsdf, u_light, u_scatter);
                want_sh = de.shadow; want_mis = de.mis;
                c = de.ld_light;
            }
            const bool want = want_sh || want_mis;
            n_real_sh += want_sh ? 1 : 0;
            n_real_mis += want_mis ? 1 : 0;
            if (want && rec < 0) {
                rec = atomicAdd(&W.counters[3], 1);
                for (int j = 0; j < e; ++j) tree_slot_clear(W, (long long)rec * stride + j, kMode != kTreeWhitted);
            }
            if (rec >= 0) {
                const long long sl = (long long)rec * stride + e;
                if (!want) tree_slot_clear(W, sl, kMode != kTreeWhitted);
                else {
                    if (want_sh) store_ray(W.sh_ray, (int)sl, de.sh_o, de.sh_d, 1.0f - kShadowEps, time);
                    else store_ray(W.sh_ray, (int)sl, mk(0, 0, 0), mk(0, 0, 1), -1.0f, 0.0f);
                    W.sh_c[sl] = make_float4(c.r, c.g, c.b, __int_as_float((want_sh ? 1 : 0) | (want_mis ? 2 : 0)));
                    if (kMode != kTreeWhitted) {
                        if (want_mis) store_ray(W.mis_ray, (int)sl, de.mis_o, de.mis_d, __int_as_float(0x7f800000), time);
                        else store_ray(W.mis_ray, (int)sl, mk(0, 0, 0), mk(0, 0, 1), -1.0f, 0.0f);
                        W.dp_b[sl] = make_float4(de.mis_f.r, de.mis_f.g, de.mis_f.b, de.mis_w);
                        W.dp_c[sl] = make_float2(de.mis_pdf, __int_as_float(li));
                    }
                }
            }
        }
        if (n_real_sh) atomicAdd(&W.counters[5], n_real_sh);  // rays the reference traces too (the placeholder slots are not counted)
        if (n_real_mis) atomicAdd(&W.counters[6], n_real_mis);
        if (rec >= 0) {
            W.pend_q[rec] = pid;
            W.pend_a[pid] = make_float4(le.r, le.g, le.b, 0.0f);
            W.pend_c[pid] = make_float4(beta.r, beta.g, beta.b, 0.0f);
        } else {
            L = L + beta * le;
        }
        // specular_reflect / specular_transmit (sampler_integrator.rs:79-238)
        if (depth + 1 < S.max_depth) {
            const V3 wo_l = bsdf_to_local(bsdf, wo);
            bool r_ok = false, t_ok = false;
            V3 r_wi = mk(0, 0, 0), t_wi = r_wi;
            RGB r_beta = beta, t_beta = beta;
            for (int k = 0; k < bsdf.m->n_bxdf; ++k) {
                const DBxDF& b = bsdf.m->bx[k];
                if (wo_l.z == 0.0f) break;  // BSDF::sample_f, bsdf.rs:227-229
                if (b.kind == BX_SPEC_REFL && !r_ok) {
                    BxDFSample bs = spec_refl_sample_f(b, wo_l);
                    V3 wi = bsdf_to_world(bsdf, bs.wi);
                    if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(wi, sh.ns) != 0.0f) { r_ok = true; r_wi = wi; r_beta = beta * (bs.f * abs_dot(wi, sh.ns) / bs.pdf); }
                } else if (b.kind == BX_SPEC_TRANS && !t_ok) {
                    BxDFSample bs = spec_trans_sample_f(b, wo_l);
                    V3 wi = bsdf_to_world(bsdf, bs.wi);
                    if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(wi, sh.ns) != 0.0f) { t_ok = true; t_wi = wi; t_beta = beta * (bs.f * abs_dot(wi, sh.ns) / bs.pdf); }
                }
            }
            dim += 2;  // specular_reflect: sampler.get_2d()
            if (r_ok) {
                have_next = true;
                next_o = offset_ray_origin(sh.p, sh.p_error, sh.n, r_wi);  // Hit::spawn_ray
                next_d = r_wi; next_beta = r_beta; next_depth = depth + 1;
                // specular_transmit runs after the whole reflection subtree: park it (its get_2d is charged at the pop)
                V3 to = t_ok ? offset_ray_origin(sh.p, sh.p_error, sh.n, t_wi) : mk(0, 0, 0);
                wstack_store(W, S.max_depth, pid, sp, to, t_wi, time, t_beta, depth + 1, t_ok);
                if (have_diff) {
                    if (t_ok) specular_child_differentials(parent_diff, HD, sh.p, wo, sh.ns, dndu, dndv, t_wi, true, 1.0f, W.wstack + ((long long)pid * S.max_depth + sp) * W.wstack_n + 3);
                    specular_child_differentials(parent_diff, HD, sh.p, wo, sh.ns, dndu, dndv, r_wi, false, 1.0f, W.cam_diff + 3ll * pid);
                }
                sp += 1;
            } else {
                dim += 2;  // specular_transmit: sampler.get_2d()
                if (t_ok) {
                    have_next = true;
                    next_o = offset_ray_origin(sh.p, sh.p_error, sh.n, t_wi);
                    next_d = t_wi; next_beta = t_beta; next_depth = depth + 1;
                    if (have_diff) specular_child_differentials(parent_diff, HD, sh.p, wo, sh.ns, dndu, dndv, t_wi, true, 1.0f, W.cam_diff + 3ll * pid);
                }
            }
        }
    }
    float next_time = time;
    while (!have_next && sp > 0) {  // return to the innermost node that still owes its specular_transmit
        sp -= 1;
        const float4* e = W.wstack + ((long long)pid * S.max_depth + sp) * W.wstack_n;
        const float4 e0 = e[0], e1 = e[1], e2 = e[2];
        dim += 2;
        const int d = __float_as_int(e1.w);
        if (d >= 0) {
            have_next = true;
            if (W.cam_diff) for (int k = 0; k < 3; ++k) W.cam_diff[3ll * pid + k] = e[3 + k];
            next_o = mk(e0.x, e0.y, e0.z); next_d = mk(e1.x, e1.y, e1.z); next_time = e0.w;
            next_beta = rgb(e2.x, e2.y, e2.z); next_depth = d;
        }
    }
    W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
    if (!have_next) return;
    W.beta[pid] = make_float4(next_beta.r, next_beta.g, next_beta.b, bw.w);
    W.meta[pid] = meta_pack(dim, next_depth, sp);
    int ns = atomicAdd(&W.counters[0], 1);
    store_ray(W.ray[cur ^ 1], ns, next_o, next_d, __int_as_float(0x7f800000), next_time);
    W.qpid[cur ^ 1][ns] = pid;
}

// l = Le + the node's direct light in the reference's order; L += beta * l
template <int kMode>
__global__ void __launch_bounds__(256) k_resolve_tree(DeviceScene S, Wave W, int n_pend) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pend) return;
    const int pid = W.pend_q[i];
    const float4 a = W.pend_a[pid], b = W.pend_c[pid];
    RGB l = rgb(a.x, a.y, a.z);
    const int stride = kMode == kTreeDirectOne ? 1 : S.n_lights;
    RGB direct = rgb1(0.0f);  // uniform_sample_all_lights accumulates its own sum (common.rs:33-86)
    for (int e = 0; e < stride; ++e) {
        const long long sl = (long long)i * stride + e;
        const float4 c = W.sh_c[sl];
        const int flags = __float_as_int(c.w);
        if (kMode == kTreeWhitted) {
            if ((flags & 1) && !W.sh_occ[sl]) l = l + rgb(c.x, c.y, c.z);  // whitted.rs:104-109
            continue;
        }
        if (flags == 0) { if (kMode == kTreeDirectAll) direct = direct + rgb1(0.0f); continue; }
        RGB ld = rgb1(0.0f);
        if ((flags & 1) && !W.sh_occ[sl]) ld = ld + rgb(c.x, c.y, c.z);  // common.rs:205-225
        if (flags & 2) {                                                  // common.rs:266-296
            const float4 mf = W.dp_b[sl];
            const float2 mp = W.dp_c[sl];
            const int li = __float_as_int(mp.y);
            const DLight& light = S.lights[li];
            const float4 mh = W.mis_hit[sl];
            const float4 md = W.mis_ray[2 * sl + 1];
            const V3 wi = mk(md.x, md.y, md.z);
            const uint32_t prim = __float_as_uint(mh.y);
            RGB Li = rgb1(0.0f);
            if (prim != 0xffffffffu) {
                V3 p0, p1, p2; int mat, al; uint32_t fl;
                load_prim(S, prim, &p0, &p1, &p2, &mat, &al, &fl);
                if (al == li) Li = area_l(light, emitter_hit_normal(S, prim, p0, p1, p2, fl, mh.z, mh.w, W.mis_b2[sl]), -wi);
            } else if (light.type == LT_INFINITE) {
                Li = infinite_le(light, S.inf_distr[light.inf_slot], wi);
            }
            if (!is_black(Li)) ld = ld + rgb(mf.x, mf.y, mf.z) * Li * rgb1(1.0f) * mf.w / mp.x;
        }
        if (kMode == kTreeDirectAll) direct = direct + ld;
        else direct = ld / (1.0f / (float)S.n_lights);  // common.rs:115, 133: estimate / light_pdf
    }
    if (kMode != kTreeWhitted) l = l + direct;
    float4 Lw = W.L[pid];
    RGB L = rgb(Lw.x, Lw.y, Lw.z) + rgb(b.x, b.y, b.z) * l;
    W.L[pid] = make_float4(L.r, L.g, L.b, Lw.w);
}

}  // namespace b2
