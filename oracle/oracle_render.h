// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY: pinned by execution for Whitted-class renders, see oracle_math.h.
//
// PathIntegrator::li, direct lighting, BSDFs, materials, lights, camera, film
// and the tile-parallel render loop, restated from the reference files cited at
// each function.  Scene interchange structs come from include/b200pt.h (the
// C ABI both the oracle and the CUDA path consume).
#pragma once
#include <mutex>
#include <unordered_map>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/b200pt.h"
#include "oracle_bvh.h"
#include "oracle_envmap.h"
#include "oracle_math.h"
#include "oracle_rng.h"
#include "oracle_sobol.h"

namespace orc {

// ---------------------------------------------------------------------------
// core/src/sampling/common.rs
inline P2 concentric_sample_disk(P2 u) {  // :138-155
    Float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return P2(0.0f, 0.0f);
    Float r, theta;
    if (pabs(ox) > pabs(oy)) { r = ox; theta = kPiOver4 * (oy / ox); }
    else { r = oy; theta = kPiOver2 - kPiOver4 * (ox / oy); }
    return P2(r * std::cos(theta), r * std::sin(theta));
}
inline V3 cosine_sample_hemisphere(P2 u) {  // :207-211
    P2 d = concentric_sample_disk(u);
    Float z = std::sqrt(pmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return V3(d.x, d.y, z);
}
inline P2 uniform_sample_triangle(P2 u) {  // :198-201
    Float su0 = std::sqrt(u.x);
    return P2(1.0f - su0, u.y * su0);
}
inline Float power_heuristic(int nf, Float f_pdf, int ng, Float g_pdf) {  // :239-243
    Float f = (Float)nf * f_pdf, g = (Float)ng * g_pdf;
    return (f * f) / (f * f + g * g);
}

// core/src/pbrt/common.rs:251-276
template <class Pred> inline size_t find_interval(size_t size, Pred pred) {
    size_t first = 0, len = size;
    while (len > 0) {
        size_t half = len >> 1, middle = first + half;
        if (pred(middle)) { first = middle + 1; len -= half + 1; }
        else len = half;
    }
    if (first == 0) return 0;
    return pclamp<size_t>(first - 1, 0, size - 2);
}

// core/src/sampling/distribution_1d.rs
struct Distribution1D {
    std::vector<Float> func, cdf;
    Float func_int = 0;
    Distribution1D() {}
    explicit Distribution1D(const std::vector<Float>& f) : func(f) {  // :22-48
        size_t n = f.size();
        cdf.resize(n + 1);
        cdf[0] = 0.0f;
        for (size_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (Float)n;
        func_int = cdf[n];
        if (func_int == 0.0f) { for (size_t i = 1; i < n + 1; ++i) cdf[i] = (Float)i / (Float)n; }
        else { for (size_t i = 1; i < n + 1; ++i) cdf[i] /= func_int; }
    }
    size_t count() const { return func.size(); }
    Float sample_continuous(Float u, Float* pdf, size_t* off) const {  // :56-76
        size_t offset = find_interval(cdf.size(), [&](size_t i) { return cdf[i] <= u; });
        Float du = u - cdf[offset];
        if (cdf[offset + 1] - cdf[offset] > 0.0f) du /= cdf[offset + 1] - cdf[offset];
        *pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
        if (off) *off = offset;
        return ((Float)offset + du) / (Float)count();
    }
    size_t sample_discrete(Float u, Float* pdf) const {  // :81-94
        size_t offset = find_interval(cdf.size(), [&](size_t i) { return cdf[i] <= u; });
        *pdf = func_int > 0.0f ? func[offset] / (func_int * (Float)count()) : 0.0f;
        return offset;
    }
};
// core/src/sampling/distribution_2d.rs
struct Distribution2D {
    std::vector<Distribution1D> cond;
    Distribution1D marginal;
    void init(const std::vector<std::vector<Float>>& f) {
        cond.clear();
        std::vector<Float> mf;
        for (auto& row : f) { cond.emplace_back(row); mf.push_back(cond.back().func_int); }
        marginal = Distribution1D(mf);
    }
    P2 sample_continuous(P2 u, Float* pdf) const {
        Float pdf1, pdf0; size_t v;
        Float d1 = marginal.sample_continuous(u.y, &pdf1, &v);
        Float d0 = cond[v].sample_continuous(u.x, &pdf0, nullptr);
        *pdf = pdf0 * pdf1;
        return P2(d0, d1);
    }
    static size_t to_index(Float v, size_t n) {  // `as usize` saturating, then clamp
        size_t i = (!(v == v) || v <= 0.0f) ? 0 : (size_t)v;
        return pclamp<size_t>(i, 0, n - 1);
    }
    Float pdf(P2 p) const {
        size_t iu = to_index(p.x * (Float)cond[0].count(), cond[0].count());
        size_t iv = to_index(p.y * (Float)marginal.count(), marginal.count());
        return cond[iv].func[iu] / marginal.func_int;
    }
};

// ---------------------------------------------------------------------------
// core/src/reflection/common.rs — shading-frame trigonometry.
inline Float cos_theta(V3 w) { return w.z; }
inline Float cos2_theta(V3 w) { return w.z * w.z; }
inline Float abs_cos_theta(V3 w) { return pabs(w.z); }
inline Float sin2_theta(V3 w) { return pmax(0.0f, 1.0f - cos2_theta(w)); }
inline Float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline Float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline Float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : pclamp(w.x / s, -1.0f, 1.0f); }
inline Float sin_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 0.0f : pclamp(w.y / s, -1.0f, 1.0f); }
inline Float cos2_phi(V3 w) { Float c = cos_phi(w); return c * c; }
inline Float sin2_phi(V3 w) { Float c = sin_phi(w); return c * c; }
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
// common.rs:136-152
inline bool refract(V3 wi, V3 n, Float eta, V3* wt) {
    Float cos_i = dot(n, wi);
    Float sin2_i = pmax(0.0f, 1.0f - cos_i * cos_i);
    Float sin2_t = eta * eta * sin2_i;
    if (sin2_t >= 1.0f) return false;
    Float cos_t = std::sqrt(1.0f - sin2_t);
    *wt = eta * -wi + (eta * cos_i - cos_t) * n;
    return true;
}
// common.rs:155-158:  -wo + 2.0 * wo.dot(n) * n
inline V3 reflect(V3 wo, V3 n) { return -wo + (2.0f * dot(wo, n)) * n; }

// core/src/reflection/fresnel.rs:152-185
inline Float fr_dielectric(Float cos_i, Float eta_i, Float eta_t) {
    cos_i = pclamp(cos_i, -1.0f, 1.0f);
    bool entering = cos_i > 0.0f;
    if (!entering) { Float t = eta_i; eta_i = eta_t; eta_t = t; cos_i = pabs(cos_i); }
    Float sin_i = std::sqrt(std::fmax(0.0f, 1.0f - cos_i * cos_i));
    Float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1.0f) return 1.0f;
    Float cos_t = std::sqrt(std::fmax(0.0f, 1.0f - sin_t * sin_t));
    Float r_parl = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    Float r_perp = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
// fresnel.rs:187-210 — QUIRK kept: sin^2(theta) = 1 - cos(theta) (not 1 - cos^2).
inline RGB fr_conductor(Float cos_i, RGB eta_i, RGB eta_t, RGB k) {
    cos_i = pclamp(cos_i, -1.0f, 1.0f);
    RGB eta = eta_t / eta_i;
    RGB eta_k = k / eta_i;
    Float cos2 = cos_i * cos_i;
    Float sin2 = 1.0f - cos_i;
    RGB eta2 = eta * eta;
    RGB etak2 = eta_k * eta_k;
    RGB t0 = eta2 - etak2 - RGB(sin2);
    RGB a2pb2 = rgb_sqrt(t0 * t0 + 4.0f * eta2 * etak2);
    RGB t1 = a2pb2 + RGB(cos2);
    RGB a = rgb_sqrt(0.5f * (a2pb2 + t0));
    RGB t2 = 2.0f * cos_i * a;
    RGB rs = (t1 - t2) / (t1 + t2);
    RGB t3 = cos2 * a2pb2 + RGB(sin2 * sin2);
    RGB t4 = t2 * sin2;
    RGB rp = rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (rp + rs);
}

// core/src/microfacet/trowbridge_reitz.rs
inline Float tr_roughness_to_alpha(Float roughness) {  // :45-53
    roughness = pmax(roughness, 1e-3f);
    Float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
struct TRDist {
    Float ax, ay;  // alpha_x, alpha_y (already max(0.001, .)), sample_visible_area = true on this path
    Float d(V3 wh) const {  // :64-78
        Float t2 = tan2_theta(wh);
        if (std::isinf(t2)) return 0.0f;
        Float cos4 = cos2_theta(wh) * cos2_theta(wh);
        Float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * t2;
        return 1.0f / (kPi * ax * ay * cos4 * (1.0f + e) * (1.0f + e));
    }
    Float lambda(V3 w) const {  // :82-96
        Float att = pabs(tan_theta(w));
        if (std::isinf(att)) return 0.0f;
        Float alpha = std::sqrt(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
        Float a2t2 = (alpha * att) * (alpha * att);
        return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
    }
    // core/src/microfacet/mod.rs:55-90
    Float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }
    Float g(V3 wo, V3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    Float pdf(V3 wo, V3 wh) const { return d(wh) * g1(wo) * abs_dot(wo, wh) / abs_cos_theta(wo); }
    // trowbridge_reitz.rs:144-200
    static void sample11(Float cos_t, Float u1, Float u2, Float* sx, Float* sy) {
        if (cos_t > 0.9999f) {
            Float r = std::sqrt(u1 / (1.0f - u1));
            Float phi = kTwoPi * u2;
            *sx = r * std::cos(phi);
            *sy = r * std::sin(phi);
            return;
        }
        Float sin_t = std::sqrt(pmax(0.0f, 1.0f - cos_t * cos_t));
        Float tan_t = sin_t / cos_t;
        Float a = 1.0f / tan_t;
        Float g1 = 2.0f / (1.0f + std::sqrt(1.0f + 1.0f / (a * a)));
        a = 2.0f * u1 / g1 - 1.0f;
        Float tmp = 1.0f / (a * a - 1.0f);
        if (tmp > 1e10f) tmp = 1e10f;
        Float b = tan_t;
        Float dd = std::sqrt(pmax(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
        Float sx1 = b * tmp - dd, sx2 = b * tmp + dd;
        *sx = (a < 0.0f || sx2 > 1.0f / tan_t) ? sx1 : sx2;
        Float s;
        if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
        else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
        Float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) /
                  (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
        *sy = s * z * std::sqrt(1.0f + *sx * *sx);
    }
    // :202-220
    static V3 sample(V3 wi, Float ax, Float ay, Float u1, Float u2) {
        V3 ws = normalize(V3(ax * wi.x, ay * wi.y, wi.z));
        Float sx, sy;
        sample11(cos_theta(ws), u1, u2, &sx, &sy);
        Float tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
        sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
        sx = tmp;
        sx *= ax;
        sy *= ay;
        return normalize(V3(-sx, -sy, 1.0f));
    }
    V3 sample_wh(V3 wo, P2 u) const {  // :100-141, sample_visible_area branch
        bool flip = wo.z < 0.0f;
        V3 wh = sample(flip ? -wo : wo, ax, ay, u.x, u.y);
        return flip ? -wh : wh;
    }
};

// core/src/reflection/bsdf.rs:10-20
enum : uint8_t { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };

enum BxDFKind { BX_LAMBERT, BX_OREN_NAYAR, BX_MF_REFL, BX_MF_TRANS, BX_FRESNEL_SPECULAR, BX_SPEC_REFL, BX_SPEC_TRANS };
enum FresnelKind { FR_DIELECTRIC, FR_CONDUCTOR, FR_NOOP };

struct BxDFSample {
    RGB f;
    Float pdf = 0;
    V3 wi;
    uint8_t type = 0;
};

struct BxDF {
    BxDFKind kind;
    uint8_t type;
    RGB r, t;
    Float on_a = 0, on_b = 0;        // Oren-Nayar A, B
    FresnelKind fr = FR_DIELECTRIC;  // microfacet reflection
    Float fr_eta_i = 1, fr_eta_t = 1;
    RGB c_eta_i, c_eta_t, c_k;
    TRDist dist{1, 1};
    Float eta_a = 1, eta_b = 1;  // transmission / fresnel specular

    RGB fresnel(Float cos_i) const {  // fresnel.rs:13-21, 60-62, 95-98
        if (fr == FR_DIELECTRIC) return RGB(fr_dielectric(cos_i, fr_eta_i, fr_eta_t));
        if (fr == FR_NOOP) return RGB(1.0f);  // FresnelNoOp, fresnel.rs:124-138
        return fr_conductor(pabs(cos_i), c_eta_i, c_eta_t, c_k);
    }

    RGB f(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERT: return r * kInvPi;  // lambertian_reflection.rs:38
            case BX_OREN_NAYAR: {               // oren_nayar.rs:36-57
                Float sin_i = sin_theta(wi), sin_o = sin_theta(wo);
                Float max_cos = 0.0f;
                if (sin_i > 1e-4f && sin_o > 1e-4f) {
                    Float sp_i = sin_phi(wi), cp_i = cos_phi(wi), sp_o = sin_phi(wo), cp_o = cos_phi(wo);
                    Float d_cos = cp_i * cp_o + sp_i * sp_o;
                    max_cos = pmax(0.0f, d_cos);
                }
                Float aco = abs_cos_theta(wo), aci = abs_cos_theta(wi);
                Float sin_alpha, tan_beta;
                if (aci > aco) { sin_alpha = sin_o; tan_beta = sin_i / aci; }
                else { sin_alpha = sin_i; tan_beta = sin_o / aco; }
                return r * kInvPi * (on_a + on_b * max_cos * sin_alpha * tan_beta);
            }
            case BX_MF_REFL: {  // microfacet_reflection.rs:48-66
                Float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
                V3 wh = wi + wo;
                if ((cos_i == 0.0f || cos_o == 0.0f) || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f)) return RGB();
                wh = normalize(wh);
                RGB F = fresnel(dot(wi, face_forward(wh, V3(0.0f, 0.0f, 1.0f))));
                return r * dist.d(wh) * dist.g(wo, wi) * F / (4.0f * cos_i * cos_o);
            }
            case BX_MF_TRANS: {  // microfacet_transmission.rs:70-123
                if (same_hemisphere(wo, wi)) return RGB();
                Float cos_o = cos_theta(wo), cos_i = cos_theta(wi);
                if (cos_i == 0.0f || cos_o == 0.0f) return RGB();
                Float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize(wo + wi * eta);
                if (wh.z < 0.0f) wh = -wh;
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return RGB();
                RGB F = RGB(fr_dielectric(dot(wo, wh), eta_a, eta_b));
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float factor = 1.0f / eta;  // TransportMode::Radiance
                return (RGB(1.0f) - F) * t *
                       pabs(dist.d(wh) * dist.g(wo, wi) * eta * eta * abs_dot(wi, wh) * abs_dot(wo, wh) * factor * factor /
                            (cos_i * cos_o * sqrt_denom * sqrt_denom));
            }
            case BX_FRESNEL_SPECULAR: return RGB();  // fresnel_specular.rs:63-66
            case BX_SPEC_REFL: return RGB();         // specular_reflection.rs:40-43
            case BX_SPEC_TRANS: return RGB();        // specular_transmission.rs:55-58
        }
        return RGB();
    }

    Float pdf(V3 wo, V3 wi) const {
        switch (kind) {
            case BX_LAMBERT:
            case BX_OREN_NAYAR:  // reflection/mod.rs:160-167 default arm
                return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0.0f;
            case BX_MF_REFL: {  // microfacet_reflection.rs:96-103
                if (!same_hemisphere(wo, wi)) return 0.0f;
                V3 wh = normalize(wo + wi);
                return dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
            }
            case BX_MF_TRANS: {  // microfacet_transmission.rs:151-172
                if (same_hemisphere(wo, wi)) return 0.0f;
                Float eta = cos_theta(wo) > 0.0f ? eta_b / eta_a : eta_a / eta_b;
                V3 wh = normalize(wo + wi * eta);
                if (dot(wo, wh) * dot(wi, wh) > 0.0f) return 0.0f;
                Float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
                Float dwh_dwi = pabs((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
                return dist.pdf(wo, wh) * dwh_dwi;
            }
            case BX_FRESNEL_SPECULAR: return 0.0f;
            case BX_SPEC_REFL: return 0.0f;   // specular_reflection.rs:53-55
            case BX_SPEC_TRANS: return 0.0f;  // specular_transmission.rs:83-85
        }
        return 0.0f;
    }

    BxDFSample sample_f(V3 wo, P2 u) const {
        BxDFSample s;
        s.type = type;
        switch (kind) {
            case BX_LAMBERT:
            case BX_OREN_NAYAR: {  // reflection/mod.rs:132-141 default arm
                V3 wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0f) wi.z *= -1.0f;
                s.pdf = pdf(wo, wi);
                s.f = f(wo, wi);
                s.wi = wi;
                return s;
            }
            case BX_MF_REFL: {  // microfacet_reflection.rs:68-94
                if (wo.z == 0.0f) return s;
                V3 wh = dist.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0f) return s;
                V3 wi = reflect(wo, wh);
                if (!same_hemisphere(wo, wi)) { s.wi = wi; return s; }
                s.pdf = dist.pdf(wo, wh) / (4.0f * dot(wo, wh));
                s.f = f(wo, wi);
                s.wi = wi;
                return s;
            }
            case BX_MF_TRANS: {  // microfacet_transmission.rs:125-149
                if (wo.z == 0.0f) return s;
                V3 wh = dist.sample_wh(wo, u);
                if (dot(wo, wh) < 0.0f) return s;
                Float eta = cos_theta(wo) > 0.0f ? eta_a / eta_b : eta_b / eta_a;
                V3 wi;
                if (!refract(wo, wh, eta, &wi)) return s;
                s.pdf = pdf(wo, wi);
                s.f = f(wo, wi);
                s.wi = wi;
                return s;
            }
            case BX_SPEC_REFL: {  // specular_reflection.rs:45-51 (fresnel = FresnelDielectric(fr_eta_i, fr_eta_t))
                V3 wi(-wo.x, -wo.y, wo.z);
                s.pdf = 1.0f;
                s.f = fresnel(cos_theta(wi)) * r / abs_cos_theta(wi);
                s.wi = wi;
                return s;
            }
            case BX_SPEC_TRANS: {  // specular_transmission.rs:60-81 (fresnel = FresnelDielectric(eta_a, eta_b), Radiance mode)
                bool entering = cos_theta(wo) > 0.0f;
                Float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
                V3 wi;
                if (!refract(wo, face_forward(V3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, &wi)) return s;
                s.pdf = 1.0f;
                RGB ft = t * (RGB(1.0f) - RGB(fr_dielectric(cos_theta(wi), eta_a, eta_b)));
                ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));
                s.f = ft / abs_cos_theta(wi);
                s.wi = wi;
                return s;
            }
            case BX_FRESNEL_SPECULAR: {  // fresnel_specular.rs:68-103
                Float F = fr_dielectric(cos_theta(wo), eta_a, eta_b);
                if (u.x < F) {
                    V3 wi(-wo.x, -wo.y, wo.z);
                    s.type = BSDF_SPECULAR | BSDF_REFLECTION;
                    s.pdf = F;
                    s.f = F * r / abs_cos_theta(wi);
                    s.wi = wi;
                    return s;
                }
                bool entering = cos_theta(wo) > 0.0f;
                Float eta_i = entering ? eta_a : eta_b, eta_t = entering ? eta_b : eta_a;
                s.type = BSDF_SPECULAR | BSDF_TRANSMISSION;
                V3 wi;
                if (!refract(wo, face_forward(V3(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, &wi)) return s;
                RGB ft = t * (1.0f - F);
                ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));  // TransportMode::Radiance
                s.pdf = 1.0f - F;
                s.f = ft / abs_cos_theta(wi);
                s.wi = wi;
                return s;
            }
        }
        return s;
    }
    bool matches(uint8_t flags) const { return (type & flags) == type; }  // reflection/mod.rs:82-85
};

// core/src/reflection/bsdf.rs
struct BSDF {
    V3 ns, ng, ss, ts;
    Float eta = 1.0f;
    BxDF bx[2];
    int n = 0;
    void add(const BxDF& b) { bx[n++] = b; }
    int num_components(uint8_t flags) const { int c = 0; for (int i = 0; i < n; ++i) if (bx[i].matches(flags)) ++c; return c; }
    V3 world_to_local(V3 v) const { return V3(dot(v, ss), dot(v, ts), dot(v, ns)); }
    V3 local_to_world(V3 v) const {
        return V3(ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z);
    }
    RGB f(V3 wo_w, V3 wi_w, uint8_t flags) const {  // :166-192
        V3 wi = world_to_local(wi_w), wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return RGB();
        bool refl = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
        RGB f;
        for (int i = 0; i < n; ++i)
            if (bx[i].matches(flags) && ((refl && (bx[i].type & BSDF_REFLECTION)) || (!refl && (bx[i].type & BSDF_TRANSMISSION))))
                f += bx[i].f(wo, wi);
        return f;
    }
    Float pdf(V3 wo_w, V3 wi_w, uint8_t flags) const {  // :331-356
        if (n == 0) return 0.0f;
        V3 wo = world_to_local(wo_w), wi = world_to_local(wi_w);
        if (wo.z == 0.0f) return 0.0f;
        int m = 0;
        Float p = 0.0f;
        for (int i = 0; i < n; ++i)
            if (bx[i].matches(flags)) { ++m; p += bx[i].pdf(wo, wi); }
        return m > 0 ? p / (Float)m : 0.0f;
    }
    BxDFSample sample_f(V3 wo_w, P2 u, uint8_t flags) const {  // :194-292
        BxDFSample none;
        int m = num_components(flags);
        if (m == 0) return none;
        int comp = (int)pmin<size_t>((size_t)std::floor(u.x * (Float)m), (size_t)(m - 1));
        int count = comp, idx = -1;
        for (int i = 0; i < n; ++i)
            if (bx[i].matches(flags)) { if (count == 0) { idx = i; break; } --count; }
        P2 ur(pmin(u.x * (Float)m - (Float)comp, kOneMinusEpsilon), u.y);
        V3 wo = world_to_local(wo_w);
        if (wo.z == 0.0f) return none;
        BxDFSample s = bx[idx].sample_f(wo, ur);
        if (s.pdf == 0.0f) return none;
        V3 wi_w = local_to_world(s.wi);
        if (!(s.type & BSDF_SPECULAR) && m > 1)
            for (int i = 0; i < n; ++i)
                if (i != idx && bx[i].matches(flags)) s.pdf += bx[i].pdf(wo, s.wi);
        if (m > 1) s.pdf /= (Float)m;
        if (!(s.type & BSDF_SPECULAR)) {
            bool refl = dot(wi_w, ng) * dot(wo_w, ng) > 0.0f;
            s.f = RGB();
            for (int i = 0; i < n; ++i)
                if (bx[i].matches(flags) && ((refl && (bx[i].type & BSDF_REFLECTION)) || (!refl && (bx[i].type & BSDF_TRANSMISSION))))
                    s.f += bx[i].f(wo, s.wi);
        }
        s.wi = wi_w;
        return s;
    }
};

// core/src/interaction: the parts of Hit / SurfaceInteraction this path reads.
struct SurfHit {
    V3 p, p_error, wo, n;  // Hit
    V3 shading_n, dpdu;    // Shading (== geometric values: no vertex normals/tangents, no bump)
    P2 uv;                 // SurfaceInteraction::uv
    V3 dpdu_g, dpdv_g;     // SurfaceInteraction::der.{dpdu, dpdv}: the geometric partials (texture filtering only)
    V3 dndu, dndv;         // Shading::{dndu, dndv} (differentials of specular children only)
    uint32_t prim = 0xffffffffu;
    Float time = 0;
};
// interaction/mod.rs:189-223
inline Ray spawn_ray(const SurfHit& h, V3 d) { return Ray(offset_ray_origin(h.p, h.p_error, h.n, d), d, kInfinity, h.time); }
inline Ray spawn_ray_to(V3 p0, V3 e0, V3 n0, V3 p1, V3 e1, V3 n1, Float time) {
    V3 origin = offset_ray_origin(p0, e0, n0, p1 - p0);
    V3 target = offset_ray_origin(p1, e1, n1, origin - p1);
    V3 d = target - origin;
    return Ray(origin, d, 1.0f - kShadowEpsilon, time);
}

struct RenderScene {
    Accel accel;
    TopLevel top;          // used instead of `accel` for traversal when has_instances
    bool has_instances = false;
    std::vector<int32_t> prim_material, prim_light;
    std::vector<b200pt_material> materials;
    std::vector<SpectrumTexture> spectrum_textures;  // textured "Kd" (matte / plastic)
    std::vector<int32_t> material_kd_tex;            // per material: index into spectrum_textures or -1 (empty = none)
    std::vector<b200pt_light> lights;
    std::vector<int> infinite_lights;
    b200pt_camera camera;
    b200pt_film film;
    b200pt_sampler sampler;
    b200pt_integrator integ;
    M4 raster_to_camera, camera_to_world;
    Bounds3 world_bound;
    V3 world_center;
    Float world_radius = 1.0f;
    Distribution1D light_distr;
    // SpatialLightDistribution (light_distrib/spatial.rs), deterministic: a voxel's distribution is a pure function of
    // the voxel, so it is computed on first use under a lock and every lookup gets it (the reference's lock-free table
    // hands out None -> uniform sampling to threads that look a voxel up while another thread is still computing it).
    bool spatial = false;
    int n_voxels[3] = {1, 1, 1};
    std::mutex spatial_mu;
    std::unordered_map<uint64_t, Distribution1D*> spatial_cache;
    ~RenderScene() { for (auto& kv : spatial_cache) delete kv.second; }
    // per infinite light (indexed by light id): 2x2 distribution of the constant map
    std::vector<Distribution2D> inf_distr;
    std::vector<MipMap> inf_map;           // per light (built for infinite lights only)
    std::vector<Float> light_area;  // area lights: Triangle::area
    int sample_bounds[4];           // Film::get_sample_bounds
    std::atomic<uint64_t> n_camera{0}, n_closest{0}, n_shadow{0};
};

inline M4 m4_from(const float* a) { M4 m; std::memcpy(m.m, a, 64); return m; }

inline RGB light_L(const b200pt_light& l) { return RGB(l.L[0], l.L[1], l.L[2]); }

// InfiniteAreaLight::le, infinite.rs:188-199 (l_map.lookup_triangle(st, 0.0): level-0 bilinear lookup, wrap = repeat;
// without a "mapname" the map is the 1x1 image [L], infinite.rs:66-79)
inline RGB infinite_le(const RenderScene& sc, int li, const Ray& ray) {
    const b200pt_light& l = sc.lights[(size_t)li];
    V3 w = normalize(xf_vector(m4_from(l.world_to_light), ray.d));
    P2 st(spherical_phi(w) * kInvTwoPi, spherical_theta(w) * kInvPi);
    return sc.inf_map[(size_t)li].lookup_triangle(st, 0.0f);
}

// Triangle::area, triangle.rs:906-911
inline Float triangle_area(V3 p0, V3 p1, V3 p2) { return 0.5f * length(cross(p1 - p0, p2 - p0)); }

// Light::power (point.rs:96, diffuse.rs:131-134, infinite.rs:177-186)
inline RGB light_power(const RenderScene& sc, int li) {
    const b200pt_light& l = sc.lights[li];
    if (l.type == B200PT_LIGHT_POINT) return kFourPi * light_L(l);
    if (l.type == B200PT_LIGHT_DISTANT) return light_L(l) * kPi * sc.world_radius * sc.world_radius;  // distant.rs:92-95
    if (l.type == B200PT_LIGHT_PROJECTION) {  // projection.rs:173-183
        RGB spec(1.0f);
        if (!sc.inf_map[(size_t)li].pyramid.empty()) spec = sc.inf_map[(size_t)li].lookup_triangle(P2(0.5f, 0.5f), 0.5f);
        return spec * light_L(l) * kTwoPi * (1.0f - l.cos_total_width);
    }
    if (l.type == B200PT_LIGHT_GONIOMETRIC) {  // goniometric.rs:140-150
        RGB spec(1.0f);
        if (!sc.inf_map[(size_t)li].pyramid.empty()) spec = sc.inf_map[(size_t)li].lookup_triangle(P2(0.5f, 0.5f), 0.5f);
        return (4.0f * kPi) * light_L(l) * spec;
    }
    if (l.type == B200PT_LIGHT_SPOT) return light_L(l) * kTwoPi * (1.0f - 0.5f * (l.cos_falloff_start + l.cos_total_width));  // spot.rs:109-111
    if (l.type == B200PT_LIGHT_AREA) {
        Float s = l.two_sided ? 2.0f : 1.0f;
        return s * light_L(l) * sc.light_area[li] * kPi;
    }
    RGB spec = sc.inf_map[(size_t)li].lookup_triangle(P2(0.5f, 0.5f), 0.5f);  // infinite.rs:177-186
    return kPi * sc.world_radius * sc.world_radius * spec;
}

inline FloatTexture float_texture_from(const b200pt_float_texture& t) {
    FloatTexture f;
    f.type = t.type; f.su = t.su; f.sv = t.sv; f.du = t.du; f.dv = t.dv; f.a = t.value[0]; f.b = t.value[1];
    f.wrap = t.wrap; f.width = t.width; f.height = t.height;
    if (t.type == B200PT_TEX_IMAGEMAP && t.texels) f.texels.assign(t.texels, t.texels + (size_t)t.width * t.height);
    return f;
}

inline RenderScene* scene_create(const b200pt_scene_desc* d) {
    RenderScene* s = new RenderScene();
    s->accel.nodes.resize((size_t)d->n_nodes);
    std::memcpy(s->accel.nodes.data(), d->nodes, (size_t)d->n_nodes * sizeof(LinearBVHNode));
    const int64_t n_top = d->n_objects > 0 ? d->n_top_tris + d->n_instances : d->n_prims;  // primitives of the scene aggregate
    s->accel.ordered.assign(d->ordered_prims, d->ordered_prims + n_top);
    s->accel.verts.assign(d->tri_verts, d->tri_verts + 9 * d->n_prims);
    if (d->prim_flags) s->accel.flags.assign(d->prim_flags, d->prim_flags + d->n_prims);
    if (d->tri_uvs) s->accel.uvs.assign(d->tri_uvs, d->tri_uvs + 6 * d->n_prims);
    if (d->tri_normals) s->accel.normals.assign(d->tri_normals, d->tri_normals + 9 * d->n_prims);
    if (d->tri_tangents) s->accel.tangents.assign(d->tri_tangents, d->tri_tangents + 9 * d->n_prims);
    if (d->float_textures && d->n_float_textures > 0 && d->prim_alpha_tex) {  // alpha masks
        s->accel.alpha_tex.assign(d->prim_alpha_tex, d->prim_alpha_tex + 2 * d->n_prims);
        for (int k = 0; k < d->n_float_textures; ++k) s->accel.textures.push_back(float_texture_from(d->float_textures[k]));
        if (d->noise_perm) s->accel.noise_perm.assign(d->noise_perm, d->noise_perm + 256);
    }
    if (d->prim_material) s->prim_material.assign(d->prim_material, d->prim_material + d->n_prims);
    if (d->prim_light) s->prim_light.assign(d->prim_light, d->prim_light + d->n_prims);
    if (d->n_objects > 0) {
        s->has_instances = true;
        s->top.top = s->accel;  // top-level nodes / ordered + all vertices and flags (global ids)
        s->top.n_top_tris = d->n_top_tris;
        for (int o = 0; o < d->n_objects; ++o) {
            const b200pt_object& ob = d->objects[o];
            Accel a;
            a.nodes.resize((size_t)ob.n_nodes);
            std::memcpy(a.nodes.data(), ob.nodes, (size_t)ob.n_nodes * sizeof(LinearBVHNode));
            a.ordered.assign(ob.ordered_prims, ob.ordered_prims + ob.n_prims);
            a.verts.assign(d->tri_verts + 9 * ob.first_prim, d->tri_verts + 9 * (ob.first_prim + ob.n_prims));
            if (d->prim_flags) a.flags.assign(d->prim_flags + ob.first_prim, d->prim_flags + ob.first_prim + ob.n_prims);
            if (d->tri_uvs) a.uvs.assign(d->tri_uvs + 6 * ob.first_prim, d->tri_uvs + 6 * (ob.first_prim + ob.n_prims));
            if (!s->accel.alpha_tex.empty()) {
                a.alpha_tex.assign(d->prim_alpha_tex + 2 * ob.first_prim, d->prim_alpha_tex + 2 * (ob.first_prim + ob.n_prims));
                a.textures = s->accel.textures;
                a.noise_perm = s->accel.noise_perm;
            }
            s->top.objects.push_back(a);
            s->top.object_first_prim.push_back(ob.first_prim);
        }
        for (int i = 0; i < d->n_instances; ++i) {
            Instance I;
            I.object = d->instances[i].object;
            I.i2w = m4_from(d->instances[i].instance_to_world);
            I.w2i = m4_from(d->instances[i].world_to_instance);
            s->top.instances.push_back(I);
        }
    }
    s->materials.assign(d->materials, d->materials + d->n_materials);
    if (d->material_kd_tex && d->spectrum_textures && d->n_spectrum_textures > 0) {
        s->material_kd_tex.assign(d->material_kd_tex, d->material_kd_tex + d->n_materials);
        for (int k = 0; k < d->n_spectrum_textures; ++k) {
            const b200pt_spectrum_texture& t = d->spectrum_textures[k];
            SpectrumTexture T;
            T.type = t.type; T.su = t.su; T.sv = t.sv; T.du = t.du; T.dv = t.dv;
            for (int c = 0; c < 3; ++c) { T.tex1[c] = t.tex1[c]; T.tex2[c] = t.tex2[c]; }
            T.closedform = t.aa_closedform != 0;
            s->spectrum_textures.push_back(T);
        }
    }
    s->lights.assign(d->lights, d->lights + d->n_lights);
    s->camera = d->camera; s->film = d->film; s->sampler = d->sampler; s->integ = d->integrator;
    s->raster_to_camera = m4_from(d->camera.raster_to_camera);
    s->camera_to_world = m4_from(d->camera.camera_to_world);
    // Scene::new (scene.rs:50-77): world bound = root node bounds; lights preprocess.
    if (!s->accel.nodes.empty()) {
        const Float* b = s->accel.nodes[0].bounds;
        s->world_bound = Bounds3(V3(b[0], b[1], b[2]), V3(b[3], b[4], b[5]));
    }
    bounding_sphere(s->world_bound, &s->world_center, &s->world_radius);  // infinite.rs:113-117
    s->light_area.assign(s->lights.size(), 0.0f);
    s->inf_distr.resize(s->lights.size());
    s->inf_map.resize(s->lights.size());
    for (size_t i = 0; i < s->lights.size(); ++i) {
        const b200pt_light& l = s->lights[i];
        if (l.type == B200PT_LIGHT_AREA)
            s->light_area[i] = triangle_area(s->accel.vert(l.prim, 0), s->accel.vert(l.prim, 1), s->accel.vert(l.prim, 2));
        if ((l.type == B200PT_LIGHT_GONIOMETRIC || l.type == B200PT_LIGHT_PROJECTION) && l.map_rgb && l.map_width > 0 && l.map_height > 0) {  // GonioPhotometricLight::new, goniometric.rs:71-90; ProjectionLight::new, projection.rs:62-75
            std::vector<RGB> texels((size_t)l.map_width * l.map_height);
            for (size_t k = 0; k < texels.size(); ++k) texels[k] = RGB(l.map_rgb[3 * k], l.map_rgb[3 * k + 1], l.map_rgb[3 * k + 2]);
            s->inf_map[i].build(l.map_width, l.map_height, texels);
        }
        if (l.type == B200PT_LIGHT_INFINITE) {
            s->infinite_lights.push_back((int)i);
            // InfiniteAreaLight::new, infinite.rs:61-92: texels = image * L (or the 1x1 image [L]), MIPMap over them
            std::vector<RGB> texels;
            int mw = 1, mh = 1;
            if (l.map_rgb && l.map_width > 0 && l.map_height > 0) {
                mw = l.map_width; mh = l.map_height;
                texels.resize((size_t)mw * mh);
                for (size_t k = 0; k < texels.size(); ++k) texels[k] = RGB(l.map_rgb[3 * k], l.map_rgb[3 * k + 1], l.map_rgb[3 * k + 2]) * light_L(l);
            } else texels.push_back(light_L(l));
            s->inf_map[i].build(mw, mh, texels);
            // compute_scalar_image, infinite.rs:326-369: (2w x 2h) image of y * sin(theta)
            const int width = 2 * s->inf_map[i].width(), height = 2 * s->inf_map[i].height();
            const Float fwidth = 0.5f / (Float)(width < height ? width : height);
            std::vector<std::vector<Float>> img(height);
            for (int v = 0; v < height; ++v) {
                Float vp = ((Float)v + 0.5f) / (Float)height;
                Float sin_t = std::sin(kPi * ((Float)v + 0.5f) / (Float)height);
                for (int u = 0; u < width; ++u) {
                    Float up = ((Float)u + 0.5f) / (Float)width;
                    img[v].push_back(lum_y(s->inf_map[i].lookup_triangle(P2(up, vp), fwidth)) * sin_t);
                }
            }
            s->inf_distr[i].init(img);
        }
    }
    // PathIntegrator::preprocess (path.rs:81) -> create_light_sample_distribution
    // (light_distrib/mod.rs:59-70): a single light forces the uniform strategy.
    int strat = s->lights.size() == 1 ? B200PT_LIGHTS_UNIFORM : s->integ.light_strategy;
    if (strat == B200PT_LIGHTS_SPATIAL && s->integ.type == B200PT_INTEGRATOR_PATH) {  // SpatialLightDistribution::new(scene, 64), spatial.rs:57-88
        s->spatial = true;
        V3 diag = bdiagonal(s->world_bound);
        Float bmax = diag[maximum_extent(s->world_bound)];
        for (int i = 0; i < 3; ++i) {
            Float r = std::round(diag[i] / bmax * 64.0f);  // f32::round: half away from zero
            long long v = (!(r == r) || r <= 0.0f) ? 0 : (long long)r;
            s->n_voxels[i] = (int)pmax<long long>(1, v);
        }
    }
    if (!s->lights.empty()) {
        std::vector<Float> f;
        for (size_t i = 0; i < s->lights.size(); ++i) f.push_back(strat == B200PT_LIGHTS_UNIFORM ? 1.0f : lum_y(light_power(*s, (int)i)));
        s->light_distr = Distribution1D(f);
    }
    // Film::get_sample_bounds, film/mod.rs:150-159
    s->sample_bounds[0] = (int)std::floor((Float)s->film.crop[0] + 0.5f - s->film.filter_radius[0]);
    s->sample_bounds[1] = (int)std::floor((Float)s->film.crop[1] + 0.5f - s->film.filter_radius[1]);
    s->sample_bounds[2] = (int)std::ceil((Float)s->film.crop[2] - 0.5f + s->film.filter_radius[0]);
    s->sample_bounds[3] = (int)std::ceil((Float)s->film.crop[3] - 0.5f + s->film.filter_radius[1]);
    return s;
}
inline void scene_destroy(RenderScene* s) { delete s; }

// Scene::intersect -> SurfaceInteraction (scene.rs:88-91, triangle.rs:547-629,
// surface_interaction.rs:56-99, interaction/mod.rs:117-136).
// Transform::transform_point_with_abs_error, transform.rs:338-368 (affine: wp == 1)
inline V3 xf_point_abs_err(const M4& M, V3 p, V3 pe, V3* err) {
    const Float (*m)[4] = M.m;
    Float x = p.x, y = p.y, z = p.z;
    Float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
    Float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
    Float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
    Float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
    Float g3 = gamma(3);
    *err = V3((g3 + 1.0f) * (pabs(m[0][0]) * pe.x + pabs(m[0][1]) * pe.y + pabs(m[0][2]) * pe.z) +
                  g3 * (pabs(m[0][0] * x) + pabs(m[0][1] * y) + pabs(m[0][2] * z) + pabs(m[0][3])),
              (g3 + 1.0f) * (pabs(m[1][0]) * pe.x + pabs(m[1][1]) * pe.y + pabs(m[1][2]) * pe.z) +
                  g3 * (pabs(m[1][0] * x) + pabs(m[1][1] * y) + pabs(m[1][2] * z) + pabs(m[1][3])),
              (g3 + 1.0f) * (pabs(m[2][0]) * pe.x + pabs(m[2][1]) * pe.y + pabs(m[2][2]) * pe.z) +
                  g3 * (pabs(m[2][0] * x) + pabs(m[2][1] * y) + pabs(m[2][2] * z) + pabs(m[2][3])));
    if (wp == 1.0f) return V3(xp, yp, zp);
    return V3(xp, yp, zp) / wp;
}
// Transform::transform_normal, transform.rs:439-446 (inverse transpose)
inline V3 xf_normal(const M4& Minv, V3 n) {
    const Float (*mi)[4] = Minv.m;
    return V3(mi[0][0] * n.x + mi[1][0] * n.y + mi[2][0] * n.z, mi[0][1] * n.x + mi[1][1] * n.y + mi[2][1] * n.z,
              mi[0][2] * n.x + mi[1][2] * n.y + mi[2][2] * n.z);
}

inline bool scene_intersect(RenderScene& sc, Ray& ray, SurfHit* sh) {
    sc.n_closest.fetch_add(1, std::memory_order_relaxed);
    HitRecord h;
    V3 d_in = ray.d;
    int inst = -1;
    if (sc.has_instances) { if (!top_intersect(sc.top, ray, &h, &inst)) return false; }
    else if (!bvh_intersect(sc.accel, ray, &h)) return false;
    V3 p0 = sc.accel.vert(h.prim, 0), p1 = sc.accel.vert(h.prim, 1), p2 = sc.accel.vert(h.prim, 2);
    TriGeom g;
    triangle_geometry(p0, p1, p2, h.b0, h.b1, h.b2, sc.accel.attr(h.prim), &g);
    sh->prim = h.prim;
    sh->time = ray.time;
    sh->uv = g.uv; sh->dpdu_g = g.dpdu; sh->dpdv_g = g.dpdv; sh->dndu = g.dndu; sh->dndv = g.dndv;
    if (inst < 0) {
        sh->p = g.p; sh->p_error = g.p_error; sh->n = g.n; sh->shading_n = g.shading_n; sh->dpdu = g.shading_dpdu;
        V3 wo = -d_in;
        Float l2 = length_squared(wo);
        sh->wo = (l2 == 0.0f) ? wo : wo / std::sqrt(l2);
        return true;
    }
    // Hit built in instance space from the instance-space ray (Hit::new normalises wo), then
    // Transform::transform_surface_interaction (transform.rs:566-590) with primitive_to_world.
    const Instance& I = sc.top.instances[(size_t)inst];
    V3 wo_i = -xf_vector(I.w2i, d_in);
    Float l2 = length_squared(wo_i);
    wo_i = (l2 == 0.0f) ? wo_i : wo_i / std::sqrt(l2);
    bool identity = true;
    { M4 id; for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) if (I.i2w.m[a][b] != id.m[a][b]) identity = false; }
    if (identity) {  // transformed_primitive.rs:57-59
        sh->p = g.p; sh->p_error = g.p_error; sh->n = g.n; sh->shading_n = g.shading_n; sh->dpdu = g.shading_dpdu; sh->wo = wo_i;
        return true;
    }
    sh->p = xf_point_abs_err(I.i2w, g.p, g.p_error, &sh->p_error);
    sh->wo = normalize(xf_vector(I.i2w, wo_i));
    sh->n = normalize(xf_normal(I.w2i, g.n));
    V3 sn = normalize(xf_normal(I.w2i, g.shading_n));
    sh->shading_n = face_forward(sn, sh->n);
    sh->dpdu = xf_vector(I.i2w, g.shading_dpdu);
    sh->dpdu_g = xf_vector(I.i2w, g.dpdu); sh->dpdv_g = xf_vector(I.i2w, g.dpdv);  // transform.rs:577-578
    sh->dndu = xf_normal(I.w2i, g.dndu); sh->dndv = xf_normal(I.w2i, g.dndv);      // shading.dndu / dndv, transform.rs:587-588
    return true;
}
inline bool scene_intersect_p(RenderScene& sc, const Ray& ray) {
    sc.n_shadow.fetch_add(1, std::memory_order_relaxed);
    if (sc.has_instances) return top_intersect_p(sc.top, ray);
    return bvh_intersect_p(sc.accel, ray);
}

// materials/src/{matte,plastic,glass,metal}.rs compute_scattering_functions
// with constant textures, no bump map, allow_multiple_lobes = true (path.rs:145).
// BSDF::new is called with eta = None in all four, so bsdf.eta = 1.0.
// matrix4x4.rs:305-318
inline bool solve_linear_system_2x2(const Float a[2][2], const Float b[2], Float* x0, Float* x1) {
    Float det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
    if (pabs(det) < 1e-10f) return false;
    *x0 = (a[1][1] * b[0] - a[0][1] * b[1]) / det;
    *x1 = (a[0][0] * b[1] - a[1][0] * b[0]) / det;
    return !(std::isnan(*x0) || std::isnan(*x1));
}
// SurfaceInteraction::compute_differentials (surface_interaction.rs:203-277): the uv footprint of one pixel step, from
// the ray's differentials and the tangent plane at the hit.  Rays without differentials leave all of it zero.
inline UVDerivs compute_differentials(const SurfHit& sh, const Ray* ray, V3* dpdx = nullptr, V3* dpdy = nullptr) {
    UVDerivs der;
    if (dpdx) { *dpdx = V3(0.0f, 0.0f, 0.0f); *dpdy = V3(0.0f, 0.0f, 0.0f); }
    if (!ray || !ray->has_diff) return der;
    V3 n = sh.n, p = sh.p;
    Float d = dot(n, p);
    Float tx = -(dot(n, ray->rx_o) - d) / dot(n, ray->rx_d);
    if (std::isinf(tx) || std::isnan(tx)) return der;
    V3 px = ray->rx_o + tx * ray->rx_d;
    Float ty = -(dot(n, ray->ry_o) - d) / dot(n, ray->ry_d);
    if (std::isinf(ty) || std::isnan(ty)) return der;
    V3 py = ray->ry_o + ty * ray->ry_d;
    if (dpdx) { *dpdx = px - p; *dpdy = py - p; }
    int dim[2];
    if (pabs(n.x) > pabs(n.y) && pabs(n.x) > pabs(n.z)) { dim[0] = 1; dim[1] = 2; }
    else if (pabs(n.y) > pabs(n.z)) { dim[0] = 0; dim[1] = 2; }
    else { dim[0] = 0; dim[1] = 1; }
    const Float a[2][2] = {{sh.dpdu_g[dim[0]], sh.dpdv_g[dim[0]]}, {sh.dpdu_g[dim[1]], sh.dpdv_g[dim[1]]}};
    const Float bx[2] = {px[dim[0]] - p[dim[0]], px[dim[1]] - p[dim[1]]};
    const Float by[2] = {py[dim[0]] - p[dim[0]], py[dim[1]] - p[dim[1]]};
    if (!solve_linear_system_2x2(a, bx, &der.dudx, &der.dvdx)) { der.dudx = 0.0f; der.dvdx = 0.0f; }
    if (!solve_linear_system_2x2(a, by, &der.dudy, &der.dvdy)) { der.dudy = 0.0f; der.dvdy = 0.0f; }
    return der;
}

// `ray` = the ray that found the hit (isect.compute_scattering_functions(&ray, ..), path.rs:145, whitted.rs:76): its
// differentials, if any, size the footprint a textured "Kd" is filtered over.
inline BSDF make_bsdf(const RenderScene& sc, const SurfHit& sh, bool allow_multiple_lobes = true, const Ray* ray = nullptr) {
    BSDF b;
    b.ns = sh.shading_n; b.ng = sh.n;
    b.ss = normalize(sh.dpdu);
    b.ts = cross(b.ns, b.ss);
    b.eta = 1.0f;
    const int mat_index = sc.prim_material[sh.prim];
    b200pt_material m = sc.materials[(size_t)mat_index];
    if (!sc.material_kd_tex.empty() && sc.material_kd_tex[(size_t)mat_index] >= 0 && (m.type == B200PT_MAT_MATTE || m.type == B200PT_MAT_PLASTIC))
        spectrum_texture_evaluate(sc.spectrum_textures[(size_t)sc.material_kd_tex[(size_t)mat_index]], sh.uv.x, sh.uv.y, compute_differentials(sh, ray), m.kd);
    auto rgb = [](const float* c) { return RGB(c[0], c[1], c[2]); };
    switch (m.type) {
        case B200PT_MAT_MATTE: {
            RGB r = rgb_clamp0(rgb(m.kd));
            Float sig = pclamp(m.sigma, 0.0f, 90.0f);
            if (!is_black(r)) {
                BxDF x; x.type = BSDF_REFLECTION | BSDF_DIFFUSE; x.r = r;
                if (sig == 0.0f) x.kind = BX_LAMBERT;
                else {  // oren_nayar.rs:20-31
                    x.kind = BX_OREN_NAYAR;
                    Float sg = to_radians(sig), s2 = sg * sg;
                    x.on_a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
                    x.on_b = 0.45f * s2 / (s2 + 0.09f);
                }
                b.add(x);
            }
            break;
        }
        case B200PT_MAT_PLASTIC: {
            RGB kd = rgb_clamp0(rgb(m.kd));
            if (!is_black(kd)) { BxDF x; x.kind = BX_LAMBERT; x.type = BSDF_REFLECTION | BSDF_DIFFUSE; x.r = kd; b.add(x); }
            RGB ks = rgb_clamp0(rgb(m.ks));
            if (!is_black(ks)) {
                BxDF x; x.kind = BX_MF_REFL; x.type = BSDF_REFLECTION | BSDF_GLOSSY; x.r = ks;
                x.fr = FR_DIELECTRIC; x.fr_eta_i = 1.5f; x.fr_eta_t = 1.0f;
                Float rough = m.urough;
                if (m.remap_roughness) rough = tr_roughness_to_alpha(rough);
                x.dist = TRDist{pmax(0.001f, rough), pmax(0.001f, rough)};
                b.add(x);
            }
            break;
        }
        case B200PT_MAT_GLASS: {
            Float eta = m.eta[0], ur = m.urough, vr = m.vrough;
            RGB r = rgb_clamp0(rgb(m.ks)), t = rgb_clamp0(rgb(m.kt));
            if (!(is_black(r) && is_black(t))) {
                bool is_spec = ur == 0.0f && vr == 0.0f;
                if (is_spec && !allow_multiple_lobes) {  // glass.rs:112-120
                    if (!is_black(r)) {
                        BxDF x; x.kind = BX_SPEC_REFL; x.type = BSDF_REFLECTION | BSDF_SPECULAR; x.r = r;
                        x.fr = FR_DIELECTRIC; x.fr_eta_i = 1.0f; x.fr_eta_t = eta;
                        b.add(x);
                    }
                    if (!is_black(t)) {
                        BxDF x; x.kind = BX_SPEC_TRANS; x.type = BSDF_TRANSMISSION | BSDF_SPECULAR; x.t = t; x.eta_a = 1.0f; x.eta_b = eta;
                        b.add(x);
                    }
                } else if (is_spec) {
                    BxDF x; x.kind = BX_FRESNEL_SPECULAR; x.type = BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_SPECULAR;
                    x.r = r; x.t = t; x.eta_a = 1.0f; x.eta_b = eta;
                    b.add(x);
                } else {
                    if (m.remap_roughness) { ur = tr_roughness_to_alpha(ur); vr = tr_roughness_to_alpha(vr); }
                    TRDist dist{pmax(0.001f, ur), pmax(0.001f, vr)};
                    if (!is_black(r)) {
                        BxDF x; x.kind = BX_MF_REFL; x.type = BSDF_REFLECTION | BSDF_GLOSSY; x.r = r;
                        x.fr = FR_DIELECTRIC; x.fr_eta_i = 1.0f; x.fr_eta_t = eta; x.dist = dist;
                        b.add(x);
                    }
                    if (!is_black(t)) {
                        BxDF x; x.kind = BX_MF_TRANS; x.type = BSDF_TRANSMISSION | BSDF_GLOSSY; x.t = t;
                        x.eta_a = 1.0f; x.eta_b = eta; x.dist = dist;
                        b.add(x);
                    }
                }
            }
            break;
        }
        case B200PT_MAT_MIRROR: {  // mirror.rs:45-56
            RGB r = rgb_clamp0(rgb(m.ks));
            if (!is_black(r)) { BxDF x; x.kind = BX_SPEC_REFL; x.type = BSDF_REFLECTION | BSDF_SPECULAR; x.r = r; x.fr = FR_NOOP; b.add(x); }
            break;
        }
        case B200PT_MAT_METAL: {
            Float ur = m.urough, vr = m.vrough;
            if (m.remap_roughness) { ur = tr_roughness_to_alpha(ur); vr = tr_roughness_to_alpha(vr); }
            BxDF x; x.kind = BX_MF_REFL; x.type = BSDF_REFLECTION | BSDF_GLOSSY; x.r = RGB(1.0f);
            x.fr = FR_CONDUCTOR; x.c_eta_i = RGB(1.0f); x.c_eta_t = rgb(m.eta); x.c_k = rgb(m.k);
            x.dist = TRDist{pmax(0.001f, ur), pmax(0.001f, vr)};
            b.add(x);
            break;
        }
    }
    return b;
}

// Result of Light::sample_li plus the VisibilityTester endpoints.
struct LiSample {
    bool valid = false;
    V3 wi;
    Float pdf = 0;
    RGB value;
    V3 p1, p1_err, p1_n;  // light-side Hit
};

// DiffuseAreaLight::l, diffuse.rs:220-226
inline RGB area_l(const b200pt_light& l, V3 n, V3 w) { return (l.two_sided || dot(n, w) > 0.0f) ? light_L(l) : RGB(); }

// ProjectionLight::projection (projection.rs:115-139).  light_projection = Transform::perspective(fov, 1e-3, 1e30): its x / y rows are
// (inv_tan, 0, 0, 0) / (0, inv_tan, 0, 0) and w = z, so transform_point is (inv_tan x, inv_tan y, ..) * (1 / z).
inline RGB projection_light_scale(const RenderScene& sc, int li, V3 w) {
    const b200pt_light& l = sc.lights[(size_t)li];
    V3 wl = xf_vector(m4_from(l.world_to_light), w);
    if (wl.z < 1e-3f) return RGB();
    const Float inv_tan = 1.0f / std::tan(to_radians(l.fov) / 2.0f);
    const Float xp = inv_tan * wl.x, yp = inv_tan * wl.y, wp = wl.z;
    Float px = xp, py = yp;
    if (wp != 1.0f) { Float inv = 1.0f / wp; px = xp * inv; py = yp * inv; }
    const bool has_map = !sc.inf_map[(size_t)li].pyramid.empty();
    const Float aspect = (l.map_rgb && l.map_width > 0 && l.map_height > 0) ? (Float)l.map_width / (Float)l.map_height : 1.0f;
    Float x0, y0, x1, y1;
    if (aspect > 1.0f) { x0 = -aspect; y0 = -1.0f; x1 = aspect; y1 = 1.0f; }
    else { x0 = -1.0f; y0 = -1.0f / aspect; x1 = 1.0f; y1 = 1.0f / aspect; }
    if (!(px >= x0 && px <= x1 && py >= y0 && py <= y1)) return RGB();
    if (!has_map) return RGB(1.0f);
    Float ox = px - x0, oy = py - y0;  // Bounds2::offset, bounds2.rs:161-173
    if (x1 > x0) ox /= x1 - x0;
    if (y1 > y0) oy /= y1 - y0;
    return sc.inf_map[(size_t)li].lookup_triangle(P2(ox, oy), 0.0f);
}

inline LiSample light_sample_li(const RenderScene& sc, int li, const SurfHit& hit, P2 u) {
    const b200pt_light& l = sc.lights[li];
    LiSample r;
    if (l.type == B200PT_LIGHT_POINT) {  // point.rs:83-94
        V3 pl(l.pos[0], l.pos[1], l.pos[2]);
        r.wi = normalize(pl - hit.p);
        r.pdf = 1.0f;
        r.p1 = pl;
        r.value = light_L(l) / distance_squared(pl, hit.p);
        r.valid = true;
        return r;
    }
    if (l.type == B200PT_LIGHT_PROJECTION) {  // projection.rs:160-171 with projection(), :115-139
        V3 pl(l.pos[0], l.pos[1], l.pos[2]);
        r.wi = normalize(pl - hit.p);
        r.pdf = 1.0f;
        r.p1 = pl;
        r.value = light_L(l) * projection_light_scale(sc, li, -r.wi) / distance_squared(pl, hit.p);
        r.valid = true;
        return r;
    }
    if (l.type == B200PT_LIGHT_GONIOMETRIC) {  // goniometric.rs:127-138 with scale(), :101-115
        V3 pl(l.pos[0], l.pos[1], l.pos[2]);
        r.wi = normalize(pl - hit.p);
        r.pdf = 1.0f;
        r.p1 = pl;
        V3 wp = normalize(xf_vector(m4_from(l.world_to_light), -r.wi));
        std::swap(wp.y, wp.z);
        RGB scale(1.0f);
        if (!sc.inf_map[(size_t)li].pyramid.empty()) scale = sc.inf_map[(size_t)li].lookup_triangle(P2(spherical_phi(wp) * kInvTwoPi, spherical_theta(wp) * kInvPi), 0.0f);
        r.value = light_L(l) * scale / distance_squared(pl, hit.p);
        r.valid = true;
        return r;
    }
    if (l.type == B200PT_LIGHT_SPOT) {  // spot.rs:97-107 with falloff(), :62-76
        V3 pl(l.pos[0], l.pos[1], l.pos[2]);
        r.wi = normalize(pl - hit.p);
        r.pdf = 1.0f;
        r.p1 = pl;
        V3 wl = normalize(xf_vector(m4_from(l.world_to_light), -r.wi));
        Float cos_theta = wl.z, falloff;
        if (cos_theta < l.cos_total_width) falloff = 0.0f;
        else if (cos_theta >= l.cos_falloff_start) falloff = 1.0f;
        else {
            Float delta = (cos_theta - l.cos_total_width) / (l.cos_falloff_start - l.cos_total_width);
            falloff = (delta * delta) * (delta * delta);
        }
        r.value = light_L(l) * falloff / distance_squared(pl, hit.p);
        r.valid = true;
        return r;
    }
    if (l.type == B200PT_LIGHT_DISTANT) {  // distant.rs:81-90
        r.wi = V3(l.pos[0], l.pos[1], l.pos[2]);
        r.pdf = 1.0f;
        r.p1 = hit.p + r.wi * (2.0f * sc.world_radius);
        r.value = light_L(l);
        r.valid = true;
        return r;
    }
    if (l.type == B200PT_LIGHT_AREA) {
        // Triangle::sample, triangle.rs:918-949
        V3 p0 = sc.accel.vert(l.prim, 0), p1 = sc.accel.vert(l.prim, 1), p2 = sc.accel.vert(l.prim, 2);
        P2 b = uniform_sample_triangle(u);
        V3 p = b.x * p0 + b.y * p1 + (1.0f - b.x - b.y) * p2;
        V3 n = normalize(cross(p1 - p0, p2 - p0));
        const TriAttr at = sc.accel.attr(l.prim);
        if (at.n) {  // triangle.rs:931-937: orient like intersect() does
            V3 ns = b.x * V3(at.n[0], at.n[1], at.n[2]) + b.y * V3(at.n[3], at.n[4], at.n[5]) + (1.0f - b.x - b.y) * V3(at.n[6], at.n[7], at.n[8]);
            n = face_forward(n, ns);
        } else if (at.flip) n = -1.0f * n;
        V3 pas = vabs(b.x * p0) + vabs(b.y * p1) + vabs((1.0f - b.x - b.y) * p2);
        V3 p_err = gamma(6) * V3(pas.x, pas.y, pas.z);
        Float pdf = 1.0f / sc.light_area[li];
        // Shape::sample_solid_angle, shape.rs:64-79
        V3 wi = p - hit.p;
        if (length_squared(wi) == 0.0f) pdf = 0.0f;
        else {
            wi = normalize(wi);
            pdf *= distance_squared(hit.p, p) / abs_dot(n, -wi);
            if (std::isinf(pdf)) pdf = 0.0f;
        }
        // DiffuseAreaLight::sample_li, diffuse.rs:114-129
        V3 wi2 = p - hit.p;
        Float l2 = length_squared(wi2);
        if (pdf == 0.0f || l2 == 0.0f) return r;
        wi2 = wi2 / std::sqrt(l2);
        r.wi = wi2; r.pdf = pdf;
        r.value = area_l(l, n, -wi2);
        r.p1 = p; r.p1_err = p_err; r.p1_n = n;
        r.valid = true;
        return r;
    }
    // InfiniteAreaLight::sample_li, infinite.rs:133-175
    Float map_pdf;
    P2 uv = sc.inf_distr[li].sample_continuous(u, &map_pdf);
    if (map_pdf == 0.0f) return r;
    Float theta = uv.y * kPi, phi = uv.x * kTwoPi;
    Float cos_t = std::cos(theta), sin_t = std::sin(theta);
    Float sin_p = std::sin(phi), cos_p = std::cos(phi);
    V3 wi = xf_vector(m4_from(l.light_to_world), V3(sin_t * cos_p, sin_t * sin_p, cos_t));
    Float pdf = map_pdf / (kTwoPi * kPi * sin_t);
    if (sin_t == 0.0f) pdf = 0.0f;
    r.wi = wi; r.pdf = pdf;
    r.p1 = hit.p + wi * (2.0f * sc.world_radius);
    r.value = sc.inf_map[(size_t)li].lookup_triangle(uv, 0.0f);
    r.valid = true;
    return r;
}

// Light::pdf_li
inline Float light_pdf_li(const RenderScene& sc, int li, const SurfHit& hit, V3 wi) {
    const b200pt_light& l = sc.lights[li];
    if (l.type == B200PT_LIGHT_POINT || l.type == B200PT_LIGHT_DISTANT || l.type == B200PT_LIGHT_SPOT || l.type == B200PT_LIGHT_GONIOMETRIC || l.type == B200PT_LIGHT_PROJECTION) return 0.0f;
    if (l.type == B200PT_LIGHT_AREA) {  // Shape::pdf_solid_angle, shape.rs:81-107
        Ray ray = spawn_ray(hit, wi);
        V3 p0 = sc.accel.vert(l.prim, 0), p1 = sc.accel.vert(l.prim, 1), p2 = sc.accel.vert(l.prim, 2);
        TriHit th;
        if (!triangle_test(ray, p0, p1, p2, &th)) return 0.0f;
        TriGeom g;
        if (!triangle_geometry(p0, p1, p2, th.b0, th.b1, th.b2, sc.accel.attr(l.prim), &g)) return 0.0f;
        Float pdf = distance_squared(hit.p, g.p) / (abs_dot(g.n, -wi) * sc.light_area[li]);
        return std::isinf(pdf) ? 0.0f : pdf;
    }
    // infinite.rs:201-211
    V3 w = xf_vector(m4_from(l.world_to_light), wi);
    Float theta = spherical_theta(w), phi = spherical_phi(w);
    Float sin_t = std::sin(theta);
    if (sin_t == 0.0f) return 0.0f;
    return sc.inf_distr[li].pdf(P2(phi * kInvTwoPi, theta * kInvPi)) / (kTwoPi * kPi * sin_t);
}
inline bool light_is_delta(const b200pt_light& l) { return l.type == B200PT_LIGHT_POINT || l.type == B200PT_LIGHT_DISTANT || l.type == B200PT_LIGHT_SPOT || l.type == B200PT_LIGHT_GONIOMETRIC || l.type == B200PT_LIGHT_PROJECTION; }  // DELTA_POSITION | DELTA_DIRECTION

// core/src/integrator/common.rs:146-299 (handle_media = false, specular = false)
inline RGB estimate_direct(RenderScene& sc, const SurfHit& hit, const BSDF& bsdf, P2 u_scatter, int li, P2 u_light) {
    const uint8_t flags = BSDF_ALL & ~BSDF_SPECULAR;
    const b200pt_light& light = sc.lights[li];
    RGB ld;
    Float scattering_pdf = 0.0f;
    LiSample ls = light_sample_li(sc, li, hit, u_light);
    V3 wi = ls.valid ? ls.wi : V3();
    Float light_pdf = ls.valid ? ls.pdf : 0.0f;
    RGB Li = ls.valid ? ls.value : RGB();
    if (light_pdf > 0.0f && !is_black(Li)) {
        RGB f = bsdf.f(hit.wo, wi, flags) * abs_dot(wi, hit.shading_n);
        scattering_pdf = bsdf.pdf(hit.wo, wi, flags);
        if (!is_black(f)) {
            Ray sr = spawn_ray_to(hit.p, hit.p_error, hit.n, ls.p1, ls.p1_err, ls.p1_n, hit.time);
            if (scene_intersect_p(sc, sr)) Li = RGB();
            if (!is_black(Li)) {
                if (light_is_delta(light)) ld += f * Li / light_pdf;
                else {
                    Float w = power_heuristic(1, light_pdf, 1, scattering_pdf);
                    ld += f * Li * w / light_pdf;
                }
            }
        }
    }
    if (!light_is_delta(light)) {
        BxDFSample bs = bsdf.sample_f(hit.wo, u_scatter, flags);
        scattering_pdf = bs.pdf;
        wi = bs.wi;
        RGB f = bs.f * abs_dot(wi, hit.shading_n);
        bool sampled_specular = (bs.type & BSDF_SPECULAR) != 0;
        if (!is_black(f) && scattering_pdf > 0.0f) {
            Float weight = 1.0f;
            if (!sampled_specular) {
                Float lp = light_pdf_li(sc, li, hit, wi);
                if (lp == 0.0f) return ld;
                weight = power_heuristic(1, scattering_pdf, 1, lp);
            }
            Ray ray = spawn_ray(hit, wi);
            SurfHit lh;
            RGB Li2;
            if (scene_intersect(sc, ray, &lh)) {
                if (!sc.prim_light.empty() && sc.prim_light[lh.prim] == li) Li2 = area_l(light, lh.n, -wi);
            } else if (light.type == B200PT_LIGHT_INFINITE) {
                Li2 = infinite_le(sc, li, ray);  // Light::le; zero for area lights (light/mod.rs default)
            }
            if (!is_black(Li2)) ld += f * Li2 * RGB(1.0f) * weight / scattering_pdf;
        }
    }
    return ld;
}

// SpatialLightDistribution::compute_distribution, spatial.rs:91-160
inline Distribution1D* spatial_compute(RenderScene& sc, const int pi[3]) {
    const Bounds3& wb = sc.world_bound;
    V3 p0((Float)pi[0] / (Float)sc.n_voxels[0], (Float)pi[1] / (Float)sc.n_voxels[1], (Float)pi[2] / (Float)sc.n_voxels[2]);
    V3 p1((Float)(pi[0] + 1) / (Float)sc.n_voxels[0], (Float)(pi[1] + 1) / (Float)sc.n_voxels[1], (Float)(pi[2] + 1) / (Float)sc.n_voxels[2]);
    auto blerp = [](const Bounds3& b, V3 t) { return V3(lerpf(t.x, b.pmin.x, b.pmax.x), lerpf(t.y, b.pmin.y, b.pmax.y), lerpf(t.z, b.pmin.z, b.pmax.z)); };
    Bounds3 vb(blerp(wb, p0), blerp(wb, p1));  // Bounds3::new orders the corners; they already are
    const int kSamples = 128;
    size_t n_lights = sc.lights.size();
    std::vector<Float> contrib(n_lights, 0.0f);
    for (int i = 0; i < kSamples; ++i) {
        V3 po = blerp(vb, V3(radical_inverse(0, (uint64_t)i), radical_inverse(1, (uint64_t)i), radical_inverse(2, (uint64_t)i)));
        SurfHit intr;
        intr.p = po;
        intr.wo = V3(1.0f, 0.0f, 0.0f);
        P2 u(radical_inverse(3, (uint64_t)i), radical_inverse(4, (uint64_t)i));
        for (size_t j = 0; j < n_lights; ++j) {
            LiSample ls = light_sample_li(sc, (int)j, intr, u);
            if (ls.valid && ls.pdf > 0.0f) contrib[j] += lum_y(ls.value) / ls.pdf;
        }
    }
    Float sum = 0.0f;
    for (Float c : contrib) sum += c;
    Float avg = sum / (Float)(kSamples * n_lights);
    Float min_contrib = avg > 0.0f ? 0.001f * avg : 1.0f;
    for (Float& c : contrib) c = pmax(c, min_contrib);
    return new Distribution1D(contrib);
}
// LightDistribution::lookup (spatial.rs:163-244 for the spatial strategy; uniform.rs / power.rs return their one distribution)
inline const Distribution1D& light_distr_lookup(RenderScene& sc, V3 p) {
    if (!sc.spatial) return sc.light_distr;
    V3 off = boffset(sc.world_bound, p);
    int pi[3];
    for (int i = 0; i < 3; ++i) {
        Float v = off[i] * (Float)sc.n_voxels[i];
        int q = !(v == v) ? 0 : (v >= 2147483648.0f ? 0x7fffffff : (v <= -2147483648.0f ? (int)0x80000000 : (int)v));  // `as Int` saturates
        pi[i] = pclamp(q, 0, sc.n_voxels[i] - 1);
    }
    uint64_t key = ((uint64_t)pi[0] << 40) | ((uint64_t)pi[1] << 20) | (uint64_t)pi[2];
    std::lock_guard<std::mutex> g(sc.spatial_mu);
    auto it = sc.spatial_cache.find(key);
    if (it != sc.spatial_cache.end()) return *it->second;
    Distribution1D* d = spatial_compute(sc, pi);
    sc.spatial_cache[key] = d;
    return *d;
}

// common.rs:89-133
inline RGB uniform_sample_one_light(RenderScene& sc, const SurfHit& hit, const BSDF& bsdf, Sampler& sampler) {
    size_t n_lights = sc.lights.size();
    if (n_lights == 0) return RGB();
    Float sample = sampler.get_1d();
    Float light_pdf;
    size_t ln = light_distr_lookup(sc, hit.p).sample_discrete(sample, &light_pdf);  // path.rs:156-157
    if (light_pdf == 0.0f) return RGB();
    P2 u_light = sampler.get_2d();
    P2 u_scatter = sampler.get_2d();
    return estimate_direct(sc, hit, bsdf, u_scatter, (int)ln, u_light) / light_pdf;
}

// integrators/src/path.rs:103-284 (BSSRDF branch out of scope).
inline RGB path_li(RenderScene& sc, Ray ray, Sampler& sampler) {
    RGB L, beta(1.0f);
    bool specular_bounce = false;
    Float eta_scale = 1.0f;
    int bounces = 0;
    for (;;) {
        SurfHit isect;
        V3 ray_d = ray.d;
        bool found = scene_intersect(sc, ray, &isect);
        if (bounces == 0 || specular_bounce) {
            if (found) {
                int al = sc.prim_light.empty() ? -1 : sc.prim_light[isect.prim];
                if (al >= 0) L += beta * area_l(sc.lights[al], isect.n, -ray_d);
            } else {
                for (int li : sc.infinite_lights) L += beta * infinite_le(sc, li, ray);
            }
        }
        if (!found || bounces >= sc.integ.max_depth) break;
        if (sc.prim_material[isect.prim] < 0) {  // path.rs:141-150: no BSDF (Material "" / "none"): spawn_ray(ray.d), bounces not updated
            ray = spawn_ray(isect, ray_d);
            continue;
        }
        BSDF bsdf = make_bsdf(sc, isect, true, &ray);
        if (bsdf.num_components(BSDF_ALL & ~BSDF_SPECULAR) > 0) {
            RGB ld = beta * uniform_sample_one_light(sc, isect, bsdf, sampler);
            L += ld;
        }
        P2 u = sampler.get_2d();
        V3 wo = -ray_d;
        BxDFSample bs = bsdf.sample_f(wo, u, BSDF_ALL);
        if (is_black(bs.f) || bs.pdf == 0.0f) break;
        beta *= bs.f * abs_dot(bs.wi, isect.shading_n) / bs.pdf;
        specular_bounce = (bs.type & BSDF_SPECULAR) != 0;
        if ((bs.type & BSDF_SPECULAR) && (bs.type & BSDF_TRANSMISSION)) {
            Float eta = bsdf.eta;
            eta_scale *= dot(wo, isect.n) > 0.0f ? eta * eta : 1.0f / (eta * eta);
        }
        ray = spawn_ray(isect, bs.wi);
        RGB rr_beta = beta * eta_scale;
        if (max_component_value(rr_beta) < sc.integ.rr_threshold && bounces > 3) {
            Float q = pmax(0.05f, 1.0f - max_component_value(rr_beta));
            if (sampler.get_1d() < q) break;
            beta = beta / (1.0f - q);
        }
        bounces += 1;
    }
    return L;
}

// The child ray of specular_reflect / specular_transmit (sampler_integrator.rs:104-127, 161-232): Hit::spawn_ray plus, when
// the parent ray carries differentials, the reflected / refracted differentials.  `eta_bsdf` = bsdf.eta (1.0 for every
// material of this path: BSDF::new(.., None)).
inline Ray specular_child(const SurfHit& isect, const Ray& ray, V3 wi, bool transmit, Float eta_bsdf) {
    Ray rd = spawn_ray(isect, wi);
    if (!ray.has_diff) return rd;
    V3 dpdx, dpdy;
    UVDerivs uv = compute_differentials(isect, &ray, &dpdx, &dpdy);
    const V3 wo = isect.wo, p = isect.p;
    V3 ns = isect.shading_n;
    rd.has_diff = true;
    rd.rx_o = p + dpdx;
    rd.ry_o = p + dpdy;
    V3 dndx = isect.dndu * uv.dudx + isect.dndv * uv.dvdx;
    V3 dndy = isect.dndu * uv.dudy + isect.dndv * uv.dvdy;
    if (!transmit) {
        V3 dwodx = -ray.rx_d - wo, dwody = -ray.ry_d - wo;
        Float ddndx = dot(dwodx, ns) + dot(wo, dndx);
        Float ddndy = dot(dwody, ns) + dot(wo, dndy);
        rd.rx_d = wi - dwodx + 2.0f * (dot(wo, ns) * dndx + ddndx * ns);
        rd.ry_d = wi - dwody + 2.0f * (dot(wo, ns) * dndy + ddndy * ns);
        return rd;
    }
    Float eta = 1.0f / eta_bsdf;
    if (dot(wo, ns) < 0.0f) {
        eta = 1.0f / eta;
        ns = -ns; dndx = -dndx; dndy = -dndy;
    }
    V3 dwodx = -ray.rx_d - wo, dwody = -ray.ry_d - wo;
    Float ddndx = dot(dwodx, ns) + dot(wo, dndx);
    Float ddndy = dot(dwody, ns) + dot(wo, dndy);
    Float mu = eta * dot(wo, ns) - abs_dot(wi, ns);
    Float dmudx = (eta - (eta * eta * dot(wo, ns)) / abs_dot(wi, ns)) * ddndx;
    Float dmudy = (eta - (eta * eta * dot(wo, ns)) / abs_dot(wi, ns)) * ddndy;
    rd.rx_d = wi - eta * dwodx + (mu * dndx + dmudx * ns);
    rd.ry_d = wi - eta * dwody + (mu * dndy + dmudy * ns);
    return rd;
}

// integrators/src/whitted.rs:60-126 with specular_reflect / specular_transmit of
// core/src/integrator/sampler_integrator.rs:79-238 (ray differentials - the camera ray's, then those specular_child
// derives for the reflected / refracted rays - size the filter footprint of a textured "Kd").
inline RGB whitted_li(RenderScene& sc, Ray ray, Sampler& sampler, int depth) {
    RGB l;
    SurfHit isect;
    if (!scene_intersect(sc, ray, &isect)) {
        for (int li : sc.infinite_lights) l += infinite_le(sc, li, ray);  // Light::le is zero for the other kinds
        return l;
    }
    BSDF bsdf = make_bsdf(sc, isect, false, &ray);
    const V3 n = isect.shading_n, wo = isect.wo;
    int al = sc.prim_light.empty() ? -1 : sc.prim_light[isect.prim];
    if (al >= 0) l += area_l(sc.lights[al], isect.n, wo);
    for (size_t li = 0; li < sc.lights.size(); ++li) {
        P2 u = sampler.get_2d();
        LiSample ls = light_sample_li(sc, (int)li, isect, u);
        if (!ls.valid) continue;
        if (is_black(ls.value) || ls.pdf == 0.0f) continue;
        RGB f = bsdf.f(wo, ls.wi, BSDF_ALL);
        if (!is_black(f)) {
            Ray sr = spawn_ray_to(isect.p, isect.p_error, isect.n, ls.p1, ls.p1_err, ls.p1_n, isect.time);
            if (!scene_intersect_p(sc, sr)) l += f * ls.value * abs_dot(ls.wi, n) / ls.pdf;
        }
    }
    if (depth + 1 < sc.integ.max_depth) {
        RGB refl, trans;
        {
            P2 u = sampler.get_2d();
            BxDFSample bs = bsdf.sample_f(wo, u, BSDF_REFLECTION | BSDF_SPECULAR);
            if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(bs.wi, n) != 0.0f)
                refl = bs.f * whitted_li(sc, specular_child(isect, ray, bs.wi, false, bsdf.eta), sampler, depth + 1) * abs_dot(bs.wi, n) / bs.pdf;
        }
        {
            P2 u = sampler.get_2d();
            BxDFSample bs = bsdf.sample_f(wo, u, BSDF_TRANSMISSION | BSDF_SPECULAR);
            if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(bs.wi, n) != 0.0f)
                trans = bs.f * whitted_li(sc, specular_child(isect, ray, bs.wi, true, bsdf.eta), sampler, depth + 1) * abs_dot(bs.wi, n) / bs.pdf;
        }
        l += refl + trans;
    }
    return l;
}

// integrators/src/direct_lighting.rs:82-146.  The tile samplers come from clone_sampler(), which does not carry the
// sample arrays requested in preprocess() (samplers/src/halton.rs:176-182): get_2d_array() returns empty arrays and
// uniform_sample_all_lights always takes its single-sample branch (integrator/common.rs:38-52).
inline RGB direct_li(RenderScene& sc, Ray ray, Sampler& sampler, int depth) {
    RGB l;
    SurfHit isect;
    if (!scene_intersect(sc, ray, &isect)) {
        for (int li : sc.infinite_lights) l += infinite_le(sc, li, ray);
        return l;
    }
    BSDF bsdf = make_bsdf(sc, isect, false, &ray);
    const V3 n = isect.shading_n, wo = isect.wo;
    int al = sc.prim_light.empty() ? -1 : sc.prim_light[isect.prim];
    if (al >= 0) l += area_l(sc.lights[al], isect.n, wo);
    if (!sc.lights.empty()) {
        if (sc.integ.direct_strategy == B200PT_DIRECT_ALL) {  // uniform_sample_all_lights, common.rs:25-87
            RGB sum;
            for (size_t j = 0; j < sc.lights.size(); ++j) {
                P2 u_light = sampler.get_2d();
                P2 u_scatter = sampler.get_2d();
                sum += estimate_direct(sc, isect, bsdf, u_scatter, (int)j, u_light);
            }
            l += sum;
        } else {  // uniform_sample_one_light without a distribution, common.rs:89-133
            Float n_lights = (Float)sc.lights.size();
            Float u = sampler.get_1d();
            size_t ln = (size_t)pmin(u * n_lights, n_lights - 1.0f);
            Float light_pdf = 1.0f / n_lights;
            P2 u_light = sampler.get_2d();
            P2 u_scatter = sampler.get_2d();
            l += estimate_direct(sc, isect, bsdf, u_scatter, (int)ln, u_light) / light_pdf;
        }
    }
    if (depth + 1 < sc.integ.max_depth) {
        RGB refl, trans;
        {
            P2 u = sampler.get_2d();
            BxDFSample bs = bsdf.sample_f(wo, u, BSDF_REFLECTION | BSDF_SPECULAR);
            if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(bs.wi, n) != 0.0f)
                refl = bs.f * direct_li(sc, specular_child(isect, ray, bs.wi, false, bsdf.eta), sampler, depth + 1) * abs_dot(bs.wi, n) / bs.pdf;
        }
        {
            P2 u = sampler.get_2d();
            BxDFSample bs = bsdf.sample_f(wo, u, BSDF_TRANSMISSION | BSDF_SPECULAR);
            if (bs.pdf > 0.0f && !is_black(bs.f) && abs_dot(bs.wi, n) != 0.0f)
                trans = bs.f * direct_li(sc, specular_child(isect, ray, bs.wi, true, bsdf.eta), sampler, depth + 1) * abs_dot(bs.wi, n) / bs.pdf;
        }
        l += refl + trans;
    }
    return l;
}

// Integrator::li of the scene's integrator
inline RGB integrator_li(RenderScene& sc, Ray ray, Sampler& sampler) {
    if (sc.integ.type == B200PT_INTEGRATOR_DIRECT) return direct_li(sc, ray, sampler, 0);
    return sc.integ.type == B200PT_INTEGRATOR_WHITTED ? whitted_li(sc, ray, sampler, 0) : path_li(sc, ray, sampler);
}

// cameras/src/perspective_camera.rs:144-204 (generate_ray_differential) + core/src/sampler/mod.rs:43-51; the
// differentials are scaled by 1 / sqrt(spp) as render_tile does right after (sampler_integrator.rs:357-358).
// EnvironmentCamera::generate_ray (environment_camera.rs:38-53), world space
inline Ray environment_ray(const RenderScene& sc, P2 p_film, Float time) {
    Float theta = kPi * p_film.y / (Float)sc.film.yres;
    Float phi = kTwoPi * p_film.x / (Float)sc.film.xres;
    V3 dir(std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi));
    return xf_ray(sc.camera_to_world, Ray(V3(0.0f, 0.0f, 0.0f), dir, kInfinity, time));
}

inline Ray camera_ray(const RenderScene& sc, int px, int py, Sampler& sampler, P2* p_film_out) {
    P2 fs = sampler.get_2d();
    P2 p_film((Float)px + fs.x, (Float)py + fs.y);
    Float time_u = sampler.get_1d();
    P2 p_lens = sampler.get_2d();
    *p_film_out = p_film;
    const Float time = lerpf(time_u, sc.camera.shutter_open, sc.camera.shutter_close);
    const bool want_diff = !sc.spectrum_textures.empty();
    const Float diff_scale = 1.0f / std::sqrt((Float)sampler.spp);
    if (sc.camera.type == B200PT_CAMERA_ENVIRONMENT) {
        Ray ray = environment_ray(sc, p_film, time);
        if (want_diff) {  // Camera::generate_ray_differential (core/src/camera.rs:29-78): the weight is 1, so eps = 0.05 is always taken
            const Float eps = 0.05f;
            Ray rx = environment_ray(sc, P2(p_film.x + eps, p_film.y), time);
            Ray ry = environment_ray(sc, P2(p_film.x, p_film.y + eps), time);
            ray.has_diff = true;
            ray.rx_o = ray.o + (rx.o - ray.o) / eps;
            ray.rx_d = ray.d + (rx.d - ray.d) / eps;
            ray.ry_o = ray.o + (ry.o - ray.o) / eps;
            ray.ry_d = ray.d + (ry.d - ray.d) / eps;
            ray.scale_differentials(diff_scale);
        }
        return ray;
    }
    const bool ortho = sc.camera.type == B200PT_CAMERA_ORTHOGRAPHIC;
    V3 p_camera = xf_point(sc.raster_to_camera, V3(p_film.x, p_film.y, 0.0f));
    // perspective_camera.rs:108-136 / orthographic_camera.rs:64-93
    Ray ray = ortho ? Ray(p_camera, V3(0.0f, 0.0f, 1.0f), kInfinity, time) : Ray(V3(0.0f, 0.0f, 0.0f), normalize(p_camera), kInfinity, time);
    if (sc.camera.lens_radius > 0.0f) {
        P2 cd = concentric_sample_disk(p_lens);
        P2 pl(sc.camera.lens_radius * cd.x, sc.camera.lens_radius * cd.y);
        Float ft = sc.camera.focal_distance / ray.d.z;
        V3 p_focus = ray.o + ray.d * ft;
        ray.o = V3(pl.x, pl.y, 0.0f);
        ray.d = normalize(p_focus - ray.o);
    }
    if (want_diff && ortho) {  // orthographic_camera.rs:120-146; dx_camera / dy_camera = raster_to_camera.transform_vector, :44-49
        V3 dx_camera = xf_vector(sc.raster_to_camera, V3(1.0f, 0.0f, 0.0f));
        V3 dy_camera = xf_vector(sc.raster_to_camera, V3(0.0f, 1.0f, 0.0f));
        ray.has_diff = true;
        if (sc.camera.lens_radius > 0.0f) {
            P2 cd = concentric_sample_disk(p_lens);
            P2 pl(sc.camera.lens_radius * cd.x, sc.camera.lens_radius * cd.y);
            Float ft = sc.camera.focal_distance / ray.d.z;
            V3 p_focus_x = p_camera + dx_camera + (ft * V3(0.0f, 0.0f, 1.0f));
            ray.rx_o = V3(pl.x, pl.y, 0.0f);
            ray.rx_d = normalize(p_focus_x - ray.rx_o);
            V3 p_focus_y = p_camera + dy_camera + (ft * V3(0.0f, 0.0f, 1.0f));
            ray.ry_o = V3(pl.x, pl.y, 0.0f);
            ray.ry_d = normalize(p_focus_y - ray.ry_o);
        } else {
            ray.rx_o = ray.o + dx_camera; ray.ry_o = ray.o + dy_camera;
            ray.rx_d = ray.d; ray.ry_d = ray.d;
        }
    } else if (want_diff) {
        // dx_camera / dy_camera, perspective_camera.rs:71-74
        V3 c00 = xf_point(sc.raster_to_camera, V3(0.0f, 0.0f, 0.0f));
        V3 dx_camera = xf_point(sc.raster_to_camera, V3(1.0f, 0.0f, 0.0f)) - c00;
        V3 dy_camera = xf_point(sc.raster_to_camera, V3(0.0f, 1.0f, 0.0f)) - c00;
        ray.has_diff = true;
        if (sc.camera.lens_radius > 0.0f) {  // :175-193
            P2 cd = concentric_sample_disk(p_lens);
            P2 pl(sc.camera.lens_radius * cd.x, sc.camera.lens_radius * cd.y);
            V3 dx = normalize(p_camera + dx_camera);
            Float ftx = sc.camera.focal_distance / dx.z;
            V3 p_focus_x = V3(0.0f, 0.0f, 0.0f) + (ftx * dx);
            ray.rx_o = V3(pl.x, pl.y, 0.0f);
            ray.rx_d = normalize(p_focus_x - ray.rx_o);
            V3 dy = normalize(p_camera + dy_camera);
            Float fty = sc.camera.focal_distance / dy.z;
            V3 p_focus_y = V3(0.0f, 0.0f, 0.0f) + (fty * dy);
            ray.ry_o = V3(pl.x, pl.y, 0.0f);
            ray.ry_d = normalize(p_focus_y - ray.ry_o);
        } else {  // :194-199
            ray.rx_o = ray.o; ray.ry_o = ray.o;
            ray.rx_d = normalize(p_camera + dx_camera);
            ray.ry_d = normalize(p_camera + dy_camera);
        }
    }
    Ray world = xf_ray(sc.camera_to_world, ray);
    world.scale_differentials(diff_scale);
    return world;
}

inline Sampler* make_sampler(const RenderScene& sc, uint64_t seed) {
    // SamplerIntegrator::render_tile: data.sampler.clone_sampler(tile_idx) (sampler_integrator.rs:323).
    // Halton ignores the seed (halton.rs:176-182); the (0,2) sampler seeds its PCG32 with it (zero_two_sequence.rs:45-51).
    if (sc.sampler.type == B200PT_SAMPLER_ZEROTWO) return new ZeroTwoSequenceSampler(sc.sampler.spp, sc.sampler.dimensions, seed);
    if (sc.sampler.type == B200PT_SAMPLER_SOBOL) return new SobolSampler(sc.sampler.spp, sc.sample_bounds);  // sobol.rs:98-100: the seed is ignored
    return new HaltonSampler(sc.sampler.spp, sc.sample_bounds[2] - sc.sample_bounds[0], sc.sample_bounds[3] - sc.sample_bounds[1],
                             sc.sampler.sample_at_center != 0);
}

// Positions a fresh tile sampler on sample `s` of pixel (px, py) exactly as render_tile would have reached it:
// for the (0,2) sampler every earlier pixel of the tile advances the tile's RNG in start_pixel.
inline Sampler* sampler_at(const RenderScene& sc, int px, int py, int s) {
    const int tile_size = 16;
    int ex = sc.sample_bounds[2] - sc.sample_bounds[0];
    int ntx = (ex + tile_size - 1) / tile_size;
    int tx = (px - sc.sample_bounds[0]) / tile_size, ty = (py - sc.sample_bounds[1]) / tile_size;
    Sampler* smp = make_sampler(sc, (uint64_t)(ty * ntx + tx));
    if (sc.sampler.type == B200PT_SAMPLER_ZEROTWO) {
        int x0 = sc.sample_bounds[0] + tx * tile_size, x1 = pmin(x0 + tile_size, sc.sample_bounds[2]);
        int y0 = sc.sample_bounds[1] + ty * tile_size;
        for (int y = y0; y <= py; ++y)
            for (int x = x0; x < x1; ++x) {
                if (y == py && x == px) break;
                smp->start_pixel(x, y);
            }
        smp->start_pixel(px, py);
        ((ZeroTwoSequenceSampler*)smp)->set_sample_number(s);
    } else if (sc.sampler.type == B200PT_SAMPLER_SOBOL) {
        smp->start_pixel(px, py);
        ((SobolSampler*)smp)->set_sample_number(s);
    } else {
        HaltonSampler* h = (HaltonSampler*)smp;
        h->start_pixel(px, py);
        h->cur_sample = s;
        h->dimension = 0;
        h->interval_sample_index = h->get_index_for_sample((uint64_t)s);
    }
    return smp;
}

// sampler_integrator.rs:374-401 radiance sanitisation.
inline RGB sanitize_radiance(RGB l) {
    if (has_nans(l)) return RGB(0.0f);
    if (lum_y(l) < -1e-5f) return RGB(0.0f);
    if (std::isinf(lum_y(l))) return RGB(0.0f);
    return l;
}

// core/src/film/film_tile.rs
struct FilmTile {
    int x0, y0, x1, y1;  // pixel_bounds
    std::vector<RGB> contrib;
    std::vector<Float> wsum;
};
inline FilmTile get_film_tile(const RenderScene& sc, int sx0, int sy0, int sx1, int sy1) {  // film/mod.rs:182-198
    const b200pt_film& f = sc.film;
    FilmTile t;
    int p0x = (int)std::ceil((Float)sx0 - 0.5f - f.filter_radius[0]);
    int p0y = (int)std::ceil((Float)sy0 - 0.5f - f.filter_radius[1]);
    int p1x = (int)std::floor((Float)sx1 - 0.5f + f.filter_radius[0]) + 1;
    int p1y = (int)std::floor((Float)sy1 - 0.5f + f.filter_radius[1]) + 1;
    t.x0 = pmax(p0x, f.crop[0]); t.y0 = pmax(p0y, f.crop[1]);
    t.x1 = pmin(p1x, f.crop[2]); t.y1 = pmin(p1y, f.crop[3]);
    size_t area = (size_t)pmax(0, t.x1 - t.x0) * (size_t)pmax(0, t.y1 - t.y0);
    t.contrib.assign(area, RGB());
    t.wsum.assign(area, 0.0f);
    return t;
}
inline void tile_add_sample(const RenderScene& sc, FilmTile& t, P2 p_film, RGB l, Float sample_weight) {  // film_tile.rs:62-108
    const b200pt_film& f = sc.film;
    Float ly = lum_y(l);
    if (ly > f.max_sample_luminance) l = l * f.max_sample_luminance / ly;
    Float dx = p_film.x - 0.5f, dy = p_film.y - 0.5f;
    int p0x = (int)std::ceil(dx - f.filter_radius[0]), p0y = (int)std::ceil(dy - f.filter_radius[1]);
    int p1x = (int)std::floor(dx + f.filter_radius[0]) + 1, p1y = (int)std::floor(dy + f.filter_radius[1]) + 1;
    p0x = pmax(p0x, t.x0); p0y = pmax(p0y, t.y0);
    p1x = pmin(p1x, t.x1); p1y = pmin(p1y, t.y1);
    Float inv_rx = 1.0f / f.filter_radius[0], inv_ry = 1.0f / f.filter_radius[1];
    const Float tw = 16.0f;
    int width = t.x1 - t.x0;
    for (int y = p0y; y < p1y; ++y) {
        Float fy = pabs(((Float)y - dy) * inv_ry * tw);
        int iy = (int)pmin(std::floor(fy), tw - 1.0f);
        for (int x = p0x; x < p1x; ++x) {
            Float fx = pabs(((Float)x - dx) * inv_rx * tw);
            int ix = (int)pmin(std::floor(fx), tw - 1.0f);
            Float fw = f.filter_table[iy * 16 + ix];
            size_t off = (size_t)(x - t.x0) + (size_t)(y - t.y0) * (size_t)width;
            t.contrib[off] += l * sample_weight * fw;
            t.wsum[off] += fw;
        }
    }
}

// SamplerIntegrator::render_tile, sampler_integrator.rs:312-415
inline FilmTile render_tile(RenderScene& sc, int tile_idx, int n_tiles_x, int tile_size) {
    int tx = tile_idx % n_tiles_x, ty = tile_idx / n_tiles_x;
    Sampler* sampler = make_sampler(sc, (uint64_t)tile_idx);
    int x0 = sc.sample_bounds[0] + tx * tile_size, x1 = pmin(x0 + tile_size, sc.sample_bounds[2]);
    int y0 = sc.sample_bounds[1] + ty * tile_size, y1 = pmin(y0 + tile_size, sc.sample_bounds[3]);
    FilmTile tile = get_film_tile(sc, x0, y0, x1, y1);
    const int* pb = sc.integ.pixel_bounds;
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            sampler->start_pixel(x, y);
            if (!(x >= pb[0] && x < pb[2] && y >= pb[1] && y < pb[3])) continue;
            do {
                P2 p_film;
                Ray ray = camera_ray(sc, x, y, *sampler, &p_film);
                sc.n_camera.fetch_add(1, std::memory_order_relaxed);
                RGB l = sanitize_radiance(integrator_li(sc, ray, *sampler));
                tile_add_sample(sc, tile, p_film, l, 1.0f);
            } while (sampler->start_next_sample());
        }
    delete sampler;
    return tile;
}

// SamplerIntegrator::render (sampler_integrator.rs:243-304) + Film::merge_film_tile
// (film/mod.rs:220-279) + write_image/get_pixel_rgb (:356-417).  Returns the
// wall-clock seconds of the tile loop.
inline double render(RenderScene* scp, float* rgb_out, uint64_t* stats_out, int nthreads) {
    RenderScene& sc = *scp;
    const int tile_size = 16;  // core/src/app/options.rs:58-66
    int ex = sc.sample_bounds[2] - sc.sample_bounds[0], ey = sc.sample_bounds[3] - sc.sample_bounds[1];
    int ntx = (ex + tile_size - 1) / tile_size, nty = (ey + tile_size - 1) / tile_size;
    int tile_count = ntx * nty;
    int cw = sc.film.crop[2] - sc.film.crop[0], ch = sc.film.crop[3] - sc.film.crop[1];
    std::vector<Float> xyz((size_t)cw * ch * 3, 0.0f), wsum((size_t)cw * ch, 0.0f);
    sc.n_camera = 0; sc.n_closest = 0; sc.n_shadow = 0;
    std::mutex film_lock;
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&] {
        for (;;) {
            int ti = next.fetch_add(1);
            if (ti >= tile_count) break;
            FilmTile tile = render_tile(sc, ti, ntx, tile_size);
            std::lock_guard<std::mutex> g(film_lock);
            int tw = tile.x1 - tile.x0;
            for (int y = tile.y0; y < tile.y1; ++y)
                for (int x = tile.x0; x < tile.x1; ++x) {
                    size_t to = (size_t)(x - tile.x0) + (size_t)(y - tile.y0) * tw;
                    size_t fo = (size_t)(x - sc.film.crop[0]) + (size_t)(y - sc.film.crop[1]) * cw;
                    Float c[3], z[3];
                    c[0] = tile.contrib[to].c[0]; c[1] = tile.contrib[to].c[1]; c[2] = tile.contrib[to].c[2];
                    rgb_to_xyz(c, z);
                    xyz[3 * fo] += z[0]; xyz[3 * fo + 1] += z[1]; xyz[3 * fo + 2] += z[2];
                    wsum[fo] += tile.wsum[to];
                }
        }
    };
    if (nthreads <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (size_t i = 0; i < (size_t)cw * ch; ++i) {  // get_pixel_rgb, no splats
        Float rgb[3];
        xyz_to_rgb(&xyz[3 * i], rgb);
        for (int c = 0; c < 3; ++c) {
            Float v = rgb[c];
            if (wsum[i] != 0.0f) { Float inv = 1.0f / wsum[i]; v = pmax(0.0f, v * inv); }
            v += 1.0f * 0.0f;  // splat_scale * splat_rgb[0], no splats on this path
            v *= sc.film.scale;
            rgb_out[3 * i + c] = v;
        }
    }
    if (stats_out) { stats_out[0] = sc.n_camera; stats_out[1] = sc.n_closest; stats_out[2] = sc.n_shadow; stats_out[3] = 0; }
    return secs;
}

// Li and camera ray of explicit (pixel, sample) pairs — per-sample parity checks.
inline void li_batch(RenderScene* scp, const int32_t* ps, int64_t n, float* out, int nthreads) {
    RenderScene& sc = *scp;
    auto body = [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            Sampler* s = sampler_at(sc, ps[3 * i], ps[3 * i + 1], ps[3 * i + 2]);
            P2 pf;
            Ray ray = camera_ray(sc, ps[3 * i], ps[3 * i + 1], *s, &pf);
            RGB l = sanitize_radiance(integrator_li(sc, ray, *s));
            out[3 * i] = l.c[0]; out[3 * i + 1] = l.c[1]; out[3 * i + 2] = l.c[2];
            delete s;
        }
    };
    if (nthreads <= 1 || n < 64) { body(0, (size_t)n); return; }
    std::vector<std::thread> th;
    size_t chunk = ((size_t)n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min((size_t)n, chunk * t), e = std::min((size_t)n, b + chunk);
        if (b < e) th.emplace_back(body, b, e);
    }
    for (auto& t : th) t.join();
}
inline void camera_rays(RenderScene* scp, const int32_t* ps, int64_t n, float* out8) {
    RenderScene& sc = *scp;
    for (int64_t i = 0; i < n; ++i) {
        Sampler* s = sampler_at(sc, ps[3 * i], ps[3 * i + 1], ps[3 * i + 2]);
        P2 pf;
        Ray r = camera_ray(sc, ps[3 * i], ps[3 * i + 1], *s, &pf);
        float* o = out8 + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.t_max;
        o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.time;
        delete s;
    }
}

}  // namespace orc
