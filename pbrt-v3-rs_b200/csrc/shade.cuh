// Device-side shading: BSDFs, materials, lights, samplers, camera — the parts of
// PathIntegrator::li (integrators/src/path.rs:103-284) that run between ray
// casts in the wavefront loop.  Operation order follows the reference files
// cited per function (f32, no FMA contraction: this TU is built with
// -fmad=false).  sin/cos — the only transcendentals that feed ray geometry — go
// through libm_exact.cuh, which reproduces the host libm's sinf/cosf bit for bit,
// so sampled directions (and therefore every later hit) match the CPU path.
// acosf/atan2f (constant-environment lookups only) are CUDA's IEEE-accurate
// versions and may differ from the host by an ulp in the radiance VALUE.
#pragma once
#include "libm_exact.cuh"
#include "pt_math.cuh"
#include "traverse.cuh"

namespace b2 {

struct RGB {
    float r, g, b;
};
B2_HD RGB rgb(float r, float g, float b) { RGB c; c.r = r; c.g = g; c.b = b; return c; }
B2_HD RGB rgb1(float v) { return rgb(v, v, v); }
B2_HD RGB operator+(RGB a, RGB b) { return rgb(a.r + b.r, a.g + b.g, a.b + b.b); }
B2_HD RGB operator-(RGB a, RGB b) { return rgb(a.r - b.r, a.g - b.g, a.b - b.b); }
B2_HD RGB operator*(RGB a, RGB b) { return rgb(a.r * b.r, a.g * b.g, a.b * b.b); }
B2_HD RGB operator/(RGB a, RGB b) { return rgb(a.r / b.r, a.g / b.g, a.b / b.b); }
B2_HD RGB operator*(RGB a, float f) { return rgb(a.r * f, a.g * f, a.b * f); }  // spectrum/common.rs:206-211
B2_HD RGB operator*(float f, RGB a) { return a * f; }
B2_HD RGB operator/(RGB a, float f) { return a * (1.0f / f); }                  // rgb_spectrum.rs:329-358
B2_HD bool is_black(RGB a) { return !(a.r != 0.0f) && !(a.g != 0.0f) && !(a.b != 0.0f); }
B2_HD float lum_y(RGB a) { return 0.212671f * a.r + 0.715160f * a.g + 0.072169f * a.b; }
B2_HD float max_component_value(RGB a) { return pmax(pmax(a.r, a.g), a.b); }
B2_HD RGB rgb_sqrt(RGB a) { return rgb(sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)); }

struct P2 {
    float x, y;
};
B2_HD P2 mk2(float x, float y) { P2 p; p.x = x; p.y = y; return p; }

// ---- core/src/sampling/common.rs ---------------------------------------------
B2_D P2 concentric_sample_disk(P2 u) {  // :138-155
    float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return mk2(0.0f, 0.0f);
    float r, theta;
    if (pabs(ox) > pabs(oy)) { r = ox; theta = kPiOver4 * (oy / ox); }
    else { r = oy; theta = kPiOver2 - kPiOver4 * (ox / oy); }
    return mk2(r * lmx::cosf_glibc(theta), r * lmx::sinf_glibc(theta));
}
B2_D V3 cosine_sample_hemisphere(P2 u) {  // :207-211
    P2 d = concentric_sample_disk(u);
    float z = sqrtf(pmax(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return mk(d.x, d.y, z);
}
B2_D float power_heuristic(float f_pdf, float g_pdf) {  // :239-243 with nf = ng = 1
    float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
    return (f * f) / (f * f + g * g);
}

// ---- core/src/reflection/common.rs -------------------------------------------
B2_D float cos_theta(V3 w) { return w.z; }
B2_D float cos2_theta(V3 w) { return w.z * w.z; }
B2_D float abs_cos_theta(V3 w) { return pabs(w.z); }
B2_D float sin2_theta(V3 w) { return pmax(0.0f, 1.0f - cos2_theta(w)); }
B2_D float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
B2_D float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
B2_D float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
B2_D float cos_phi(V3 w) { float s = sin_theta(w); return s == 0.0f ? 1.0f : pclamp(w.x / s, -1.0f, 1.0f); }
B2_D float sin_phi(V3 w) { float s = sin_theta(w); return s == 0.0f ? 0.0f : pclamp(w.y / s, -1.0f, 1.0f); }
B2_D float cos2_phi(V3 w) { float c = cos_phi(w); return c * c; }
B2_D float sin2_phi(V3 w) { float c = sin_phi(w); return c * c; }
B2_D bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
B2_D bool refract(V3 wi, V3 n, float eta, V3* wt) {  // :136-152
    float cos_i = dot(n, wi);
    float sin2_i = pmax(0.0f, 1.0f - cos_i * cos_i);
    float sin2_t = eta * eta * sin2_i;
    if (sin2_t >= 1.0f) return false;
    float cos_t = sqrtf(1.0f - sin2_t);
    *wt = eta * -wi + (eta * cos_i - cos_t) * n;
    return true;
}
B2_D V3 reflect(V3 wo, V3 n) { return -wo + (2.0f * dot(wo, n)) * n; }  // :155-158

// ---- core/src/reflection/fresnel.rs --------------------------------------------
B2_D float fr_dielectric(float cos_i, float eta_i, float eta_t) {  // :152-185
    cos_i = pclamp(cos_i, -1.0f, 1.0f);
    bool entering = cos_i > 0.0f;
    if (!entering) { float t = eta_i; eta_i = eta_t; eta_t = t; cos_i = pabs(cos_i); }
    float sin_i = sqrtf(fmaxf(0.0f, 1.0f - cos_i * cos_i));
    float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1.0f) return 1.0f;
    float cos_t = sqrtf(fmaxf(0.0f, 1.0f - sin_t * sin_t));
    float r_parl = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    float r_perp = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
// :187-210 — QUIRK kept: sin^2(theta) is computed as 1 - cos(theta).
B2_D RGB fr_conductor(float cos_i, RGB eta_i, RGB eta_t, RGB k) {
    cos_i = pclamp(cos_i, -1.0f, 1.0f);
    RGB eta = eta_t / eta_i;
    RGB eta_k = k / eta_i;
    float cos2 = cos_i * cos_i;
    float sin2 = 1.0f - cos_i;
    RGB eta2 = eta * eta;
    RGB etak2 = eta_k * eta_k;
    RGB t0 = eta2 - etak2 - rgb1(sin2);
    RGB a2pb2 = rgb_sqrt(t0 * t0 + 4.0f * eta2 * etak2);
    RGB t1 = a2pb2 + rgb1(cos2);
    RGB a = rgb_sqrt(0.5f * (a2pb2 + t0));
    RGB t2 = 2.0f * cos_i * a;
    RGB rs = (t1 - t2) / (t1 + t2);
    RGB t3 = cos2 * a2pb2 + rgb1(sin2 * sin2);
    RGB t4 = t2 * sin2;
    RGB rp = rs * (t3 - t4) / (t3 + t4);
    return 0.5f * (rp + rs);
}

// ---- core/src/microfacet/trowbridge_reitz.rs (sample_visible_area = true) --------
struct TRDist {
    float ax, ay;
};
B2_D float tr_d(TRDist d, V3 wh) {  // :64-78
    float t2 = tan2_theta(wh);
    if (isinf(t2)) return 0.0f;
    float cos4 = cos2_theta(wh) * cos2_theta(wh);
    float e = (cos2_phi(wh) / (d.ax * d.ax) + sin2_phi(wh) / (d.ay * d.ay)) * t2;
    return 1.0f / (kPi * d.ax * d.ay * cos4 * (1.0f + e) * (1.0f + e));
}
B2_D float tr_lambda(TRDist d, V3 w) {  // :82-96
    float att = pabs(tan_theta(w));
    if (isinf(att)) return 0.0f;
    float alpha = sqrtf(cos2_phi(w) * d.ax * d.ax + sin2_phi(w) * d.ay * d.ay);
    float a2t2 = (alpha * att) * (alpha * att);
    return (-1.0f + sqrtf(1.0f + a2t2)) / 2.0f;
}
B2_D float tr_g1(TRDist d, V3 w) { return 1.0f / (1.0f + tr_lambda(d, w)); }                              // microfacet/mod.rs:55-57
B2_D float tr_g(TRDist d, V3 wo, V3 wi) { return 1.0f / (1.0f + tr_lambda(d, wo) + tr_lambda(d, wi)); }  // :59-61
B2_D float tr_pdf(TRDist d, V3 wo, V3 wh) { return tr_d(d, wh) * tr_g1(d, wo) * abs_dot(wo, wh) / abs_cos_theta(wo); }  // :80-86
B2_D void tr_sample11(float cos_t, float u1, float u2, float* sx, float* sy) {  // trowbridge_reitz.rs:144-200
    if (cos_t > 0.9999f) {
        float r = sqrtf(u1 / (1.0f - u1));
        float phi = kTwoPi * u2;
        *sx = r * lmx::cosf_glibc(phi);
        *sy = r * lmx::sinf_glibc(phi);
        return;
    }
    float sin_t = sqrtf(pmax(0.0f, 1.0f - cos_t * cos_t));
    float tan_t = sin_t / cos_t;
    float a = 1.0f / tan_t;
    float g1 = 2.0f / (1.0f + sqrtf(1.0f + 1.0f / (a * a)));
    a = 2.0f * u1 / g1 - 1.0f;
    float tmp = 1.0f / (a * a - 1.0f);
    if (tmp > 1e10f) tmp = 1e10f;
    float b = tan_t;
    float dd = sqrtf(pmax(b * b * tmp * tmp - (a * a - b * b) * tmp, 0.0f));
    float sx1 = b * tmp - dd, sx2 = b * tmp + dd;
    *sx = (a < 0.0f || sx2 > 1.0f / tan_t) ? sx1 : sx2;
    float s;
    if (u2 > 0.5f) { s = 1.0f; u2 = 2.0f * (u2 - 0.5f); }
    else { s = -1.0f; u2 = 2.0f * (0.5f - u2); }
    float z = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f)) / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *sy = s * z * sqrtf(1.0f + *sx * *sx);
}
B2_D V3 tr_sample_wh(TRDist d, V3 wo, P2 u) {  // :100-141 (visible-area branch) + :202-220
    bool flip = wo.z < 0.0f;
    V3 wi = flip ? -wo : wo;
    V3 ws = normalize(mk(d.ax * wi.x, d.ay * wi.y, wi.z));
    float sx, sy;
    tr_sample11(cos_theta(ws), u.x, u.y, &sx, &sy);
    float tmp = cos_phi(ws) * sx - sin_phi(ws) * sy;
    sy = sin_phi(ws) * sx + cos_phi(ws) * sy;
    sx = tmp;
    sx *= d.ax;
    sy *= d.ay;
    V3 wh = normalize(mk(-sx, -sy, 1.0f));
    return flip ? -wh : wh;
}

// ---- BxDFs (core/src/reflection/*.rs) ----------------------------------------------
enum : uint32_t { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16, BSDF_ALL = 31 };
enum : int { BX_LAMBERT = 0, BX_OREN_NAYAR = 1, BX_MF_REFL = 2, BX_MF_TRANS = 3, BX_FRESNEL_SPECULAR = 4, BX_SPEC_REFL = 5, BX_SPEC_TRANS = 6 };

// Compile-time lobe masks: the shade stage runs one kernel per material class after the sort, and each only carries the
// code of the lobes its material can produce (bit k = BxDF kind k may occur; KM_COND / KM_DIEL = which Fresnel term a
// microfacet reflection may use).  KM_ALL keeps every branch (tree integrators, explicit li batches).
// KM_TEX: the material's "Kd" may be a spectrum texture evaluated per intersection (spectrum_tex.cuh).
enum : uint32_t { KM_COND = 1u << 8, KM_DIEL = 1u << 9, KM_TEX = 1u << 10, KM_ALL = 0x7fu | KM_COND | KM_DIEL | KM_TEX };
template <uint32_t KM> B2_D bool km_is(int kind, int k) { return ((KM >> k) & 1u) && ((KM & 0x7fu) == (1u << k) || kind == k); }

// One lobe with every per-material constant already evaluated on the host
// (constant textures; roughness remap via ln() done once instead of per hit).
struct DBxDF {
    int kind;
    uint32_t type;
    float r[3], t[3];
    float on_a, on_b;      // Oren-Nayar
    int conductor;         // Fresnel term of a reflection lobe: 0 dielectric, 1 conductor, 2 none (FresnelNoOp: the mirror material)
    float fr_eta_i, fr_eta_t;
    float c_eta_t[3], c_k[3];  // conductor (eta_i = 1)
    float ax, ay;          // Trowbridge-Reitz alpha (already max(1e-3, .))
    float eta_a, eta_b;    // transmission / Fresnel specular
};
struct DMaterial {
    int n_bxdf;
    int type;  // B200PT_MAT_* (sort key of the shade stage)
    int pad[2];
    DBxDF bx[2];
};

struct BxDFSample {
    RGB f;
    float pdf;
    V3 wi;
    uint32_t type;
};

B2_D RGB ldrgb(const float* c) { return rgb(c[0], c[1], c[2]); }
B2_D bool bx_matches(const DBxDF& b, uint32_t flags) { return (b.type & flags) == b.type; }  // reflection/mod.rs:82-85

template <uint32_t KM = KM_ALL> B2_D RGB bx_fresnel(const DBxDF& b, float cos_i) {
    const bool conductor = (KM & KM_COND) && (!(KM & KM_DIEL) || b.conductor);
    if (!conductor) return rgb1(fr_dielectric(cos_i, b.fr_eta_i, b.fr_eta_t));
    return fr_conductor(pabs(cos_i), rgb1(1.0f), ldrgb(b.c_eta_t), ldrgb(b.c_k));
}

template <uint32_t KM = KM_ALL> B2_D RGB bx_f(const DBxDF& b, V3 wo, V3 wi) {
    {
        if (km_is<KM>(b.kind, BX_LAMBERT)) return ldrgb(b.r) * kInvPi;  // lambertian_reflection.rs:38
        if (km_is<KM>(b.kind, BX_OREN_NAYAR)) {                        // oren_nayar.rs:36-57
            float sin_i = sin_theta(wi), sin_o = sin_theta(wo);
            float max_cos = 0.0f;
            if (sin_i > 1e-4f && sin_o > 1e-4f) {
                float sp_i = sin_phi(wi), cp_i = cos_phi(wi), sp_o = sin_phi(wo), cp_o = cos_phi(wo);
                float d_cos = cp_i * cp_o + sp_i * sp_o;
                max_cos = pmax(0.0f, d_cos);
            }
            float aco = abs_cos_theta(wo), aci = abs_cos_theta(wi);
            float sin_alpha, tan_beta;
            if (aci > aco) { sin_alpha = sin_o; tan_beta = sin_i / aci; }
            else { sin_alpha = sin_i; tan_beta = sin_o / aco; }
            return ldrgb(b.r) * kInvPi * (b.on_a + b.on_b * max_cos * sin_alpha * tan_beta);
        }
        if (km_is<KM>(b.kind, BX_MF_REFL)) {  // microfacet_reflection.rs:48-66
            float cos_o = abs_cos_theta(wo), cos_i = abs_cos_theta(wi);
            V3 wh = wi + wo;
            if ((cos_i == 0.0f || cos_o == 0.0f) || (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f)) return rgb1(0.0f);
            wh = normalize(wh);
            TRDist d{b.ax, b.ay};
            RGB F = bx_fresnel<KM>(b, dot(wi, face_forward(wh, mk(0.0f, 0.0f, 1.0f))));
            return ldrgb(b.r) * tr_d(d, wh) * tr_g(d, wo, wi) * F / (4.0f * cos_i * cos_o);
        }
        if (km_is<KM>(b.kind, BX_MF_TRANS)) {  // microfacet_transmission.rs:70-123
            if (same_hemisphere(wo, wi)) return rgb1(0.0f);
            float cos_o = cos_theta(wo), cos_i = cos_theta(wi);
            if (cos_i == 0.0f || cos_o == 0.0f) return rgb1(0.0f);
            float eta = cos_theta(wo) > 0.0f ? b.eta_b / b.eta_a : b.eta_a / b.eta_b;
            V3 wh = normalize(wo + wi * eta);
            if (wh.z < 0.0f) wh = -wh;
            if (dot(wo, wh) * dot(wi, wh) > 0.0f) return rgb1(0.0f);
            TRDist d{b.ax, b.ay};
            RGB F = rgb1(fr_dielectric(dot(wo, wh), b.eta_a, b.eta_b));
            float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
            float factor = 1.0f / eta;  // TransportMode::Radiance
            return (rgb1(1.0f) - F) * ldrgb(b.t) *
                   pabs(tr_d(d, wh) * tr_g(d, wo, wi) * eta * eta * abs_dot(wi, wh) * abs_dot(wo, wh) * factor * factor /
                        (cos_i * cos_o * sqrt_denom * sqrt_denom));
        }
        return rgb1(0.0f);  // FresnelSpecular::f, fresnel_specular.rs:63-66
    }
}

template <uint32_t KM = KM_ALL> B2_D float bx_pdf(const DBxDF& b, V3 wo, V3 wi) {
    {
        if (km_is<KM>(b.kind, BX_LAMBERT) || km_is<KM>(b.kind, BX_OREN_NAYAR)) return same_hemisphere(wo, wi) ? abs_cos_theta(wi) * kInvPi : 0.0f;  // reflection/mod.rs:160-167
        if (km_is<KM>(b.kind, BX_MF_REFL)) {                                                     // microfacet_reflection.rs:96-103
            if (!same_hemisphere(wo, wi)) return 0.0f;
            V3 wh = normalize(wo + wi);
            TRDist d{b.ax, b.ay};
            return tr_pdf(d, wo, wh) / (4.0f * dot(wo, wh));
        }
        if (km_is<KM>(b.kind, BX_MF_TRANS)) {  // microfacet_transmission.rs:151-172
            if (same_hemisphere(wo, wi)) return 0.0f;
            float eta = cos_theta(wo) > 0.0f ? b.eta_b / b.eta_a : b.eta_a / b.eta_b;
            V3 wh = normalize(wo + wi * eta);
            if (dot(wo, wh) * dot(wi, wh) > 0.0f) return 0.0f;
            float sqrt_denom = dot(wo, wh) + eta * dot(wi, wh);
            float dwh_dwi = pabs((eta * eta * dot(wi, wh)) / (sqrt_denom * sqrt_denom));
            TRDist d{b.ax, b.ay};
            return tr_pdf(d, wo, wh) * dwh_dwi;
        }
        return 0.0f;
    }
}

template <uint32_t KM = KM_ALL> B2_D BxDFSample bx_sample_f(const DBxDF& b, V3 wo, P2 u) {
    BxDFSample s;
    s.f = rgb1(0.0f); s.pdf = 0.0f; s.wi = mk(0.0f, 0.0f, 0.0f); s.type = b.type;
    {
        if (km_is<KM>(b.kind, BX_LAMBERT) || km_is<KM>(b.kind, BX_OREN_NAYAR)) {  // reflection/mod.rs:132-141
            V3 wi = cosine_sample_hemisphere(u);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            s.pdf = bx_pdf<KM>(b, wo, wi);
            s.f = bx_f<KM>(b, wo, wi);
            s.wi = wi;
            return s;
        }
        if (km_is<KM>(b.kind, BX_MF_REFL)) {  // microfacet_reflection.rs:68-94
            if (wo.z == 0.0f) return s;
            TRDist d{b.ax, b.ay};
            V3 wh = tr_sample_wh(d, wo, u);
            if (dot(wo, wh) < 0.0f) return s;
            V3 wi = reflect(wo, wh);
            if (!same_hemisphere(wo, wi)) { s.wi = wi; return s; }
            s.pdf = tr_pdf(d, wo, wh) / (4.0f * dot(wo, wh));
            s.f = bx_f<KM>(b, wo, wi);
            s.wi = wi;
            return s;
        }
        if (km_is<KM>(b.kind, BX_MF_TRANS)) {  // microfacet_transmission.rs:125-149
            if (wo.z == 0.0f) return s;
            TRDist d{b.ax, b.ay};
            V3 wh = tr_sample_wh(d, wo, u);
            if (dot(wo, wh) < 0.0f) return s;
            float eta = cos_theta(wo) > 0.0f ? b.eta_a / b.eta_b : b.eta_b / b.eta_a;
            V3 wi;
            if (!refract(wo, wh, eta, &wi)) return s;
            s.pdf = bx_pdf<KM>(b, wo, wi);
            s.f = bx_f<KM>(b, wo, wi);
            s.wi = wi;
            return s;
        }
        if (km_is<KM>(b.kind, BX_SPEC_REFL)) {  // SpecularReflection::sample_f, specular_reflection.rs:45-51 (the path integrator reaches it through the mirror material)
            V3 wi = mk(-wo.x, -wo.y, wo.z);
            s.pdf = 1.0f;
            s.f = rgb1(b.conductor == 2 ? 1.0f : fr_dielectric(cos_theta(wi), b.fr_eta_i, b.fr_eta_t)) * ldrgb(b.r) / abs_cos_theta(wi);
            s.wi = wi;
            return s;
        }
        if (!((KM >> BX_FRESNEL_SPECULAR) & 1u)) return s;
        {  // FresnelSpecular::sample_f, fresnel_specular.rs:68-103
            float F = fr_dielectric(cos_theta(wo), b.eta_a, b.eta_b);
            if (u.x < F) {
                V3 wi = mk(-wo.x, -wo.y, wo.z);
                s.type = BSDF_SPECULAR | BSDF_REFLECTION;
                s.pdf = F;
                s.f = F * ldrgb(b.r) / abs_cos_theta(wi);
                s.wi = wi;
                return s;
            }
            bool entering = cos_theta(wo) > 0.0f;
            float eta_i = entering ? b.eta_a : b.eta_b, eta_t = entering ? b.eta_b : b.eta_a;
            s.type = BSDF_SPECULAR | BSDF_TRANSMISSION;
            V3 wi;
            if (!refract(wo, face_forward(mk(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, &wi)) return s;
            RGB ft = ldrgb(b.t) * (1.0f - F);
            ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));
            s.pdf = 1.0f - F;
            s.f = ft / abs_cos_theta(wi);
            s.wi = wi;
            return s;
        }
    }
}

// SpecularReflection::sample_f (specular_reflection.rs:45-51) and SpecularTransmission::sample_f
// (specular_transmission.rs:60-81, TransportMode::Radiance): the two delta lobes glass gets when the integrator does
// not allow multiple lobes (glass.rs:112-120; WhittedIntegrator).  Their f() and pdf() are zero (bx_f / bx_pdf defaults).
B2_D BxDFSample spec_refl_sample_f(const DBxDF& b, V3 wo) {
    BxDFSample s;
    V3 wi = mk(-wo.x, -wo.y, wo.z);
    s.type = b.type;
    s.pdf = 1.0f;
    s.f = rgb1(b.conductor == 2 ? 1.0f : fr_dielectric(cos_theta(wi), b.fr_eta_i, b.fr_eta_t)) * ldrgb(b.r) / abs_cos_theta(wi);
    s.wi = wi;
    return s;
}
B2_D BxDFSample spec_trans_sample_f(const DBxDF& b, V3 wo) {
    BxDFSample s;
    s.f = rgb1(0.0f); s.pdf = 0.0f; s.wi = mk(0.0f, 0.0f, 0.0f); s.type = b.type;
    bool entering = cos_theta(wo) > 0.0f;
    float eta_i = entering ? b.eta_a : b.eta_b, eta_t = entering ? b.eta_b : b.eta_a;
    V3 wi;
    if (!refract(wo, face_forward(mk(0.0f, 0.0f, 1.0f), wo), eta_i / eta_t, &wi)) return s;
    s.pdf = 1.0f;
    RGB ft = ldrgb(b.t) * (rgb1(1.0f) - rgb1(fr_dielectric(cos_theta(wi), b.eta_a, b.eta_b)));
    ft = ft * ((eta_i * eta_i) / (eta_t * eta_t));
    s.f = ft / abs_cos_theta(wi);
    s.wi = wi;
    return s;
}

// core/src/reflection/bsdf.rs — frame + the material's lobes.
struct BSDF {
    V3 ns, ng, ss, ts;
    const DMaterial* m;
};
B2_D V3 bsdf_to_local(const BSDF& b, V3 v) { return mk(dot(v, b.ss), dot(v, b.ts), dot(v, b.ns)); }
B2_D V3 bsdf_to_world(const BSDF& b, V3 v) {
    return mk(b.ss.x * v.x + b.ts.x * v.y + b.ns.x * v.z, b.ss.y * v.x + b.ts.y * v.y + b.ns.y * v.z, b.ss.z * v.x + b.ts.z * v.y + b.ns.z * v.z);
}
B2_D int bsdf_num_components(const BSDF& b, uint32_t flags) {
    int c = 0;
    for (int i = 0; i < b.m->n_bxdf; ++i) if (bx_matches(b.m->bx[i], flags)) ++c;
    return c;
}
template <uint32_t KM = KM_ALL> B2_D RGB bsdf_f(const BSDF& b, V3 wo_w, V3 wi_w, uint32_t flags) {  // :166-192
    V3 wi = bsdf_to_local(b, wi_w), wo = bsdf_to_local(b, wo_w);
    if (wo.z == 0.0f) return rgb1(0.0f);
    bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
    RGB f = rgb1(0.0f);
    for (int i = 0; i < b.m->n_bxdf; ++i) {
        const DBxDF& x = b.m->bx[i];
        if (bx_matches(x, flags) && ((refl && (x.type & BSDF_REFLECTION)) || (!refl && (x.type & BSDF_TRANSMISSION)))) f = f + bx_f<KM>(x, wo, wi);
    }
    return f;
}
template <uint32_t KM = KM_ALL> B2_D float bsdf_pdf(const BSDF& b, V3 wo_w, V3 wi_w, uint32_t flags) {  // :331-356
    if (b.m->n_bxdf == 0) return 0.0f;
    V3 wo = bsdf_to_local(b, wo_w), wi = bsdf_to_local(b, wi_w);
    if (wo.z == 0.0f) return 0.0f;
    int m = 0;
    float p = 0.0f;
    for (int i = 0; i < b.m->n_bxdf; ++i)
        if (bx_matches(b.m->bx[i], flags)) { ++m; p += bx_pdf<KM>(b.m->bx[i], wo, wi); }
    return m > 0 ? p / (float)m : 0.0f;
}
template <uint32_t KM = KM_ALL> B2_D BxDFSample bsdf_sample_f(const BSDF& b, V3 wo_w, P2 u, uint32_t flags) {  // :194-292
    BxDFSample none;
    none.f = rgb1(0.0f); none.pdf = 0.0f; none.wi = mk(0.0f, 0.0f, 0.0f); none.type = 0;
    int m = bsdf_num_components(b, flags);
    if (m == 0) return none;
    int comp = (int)floorf(u.x * (float)m);
    if (comp > m - 1) comp = m - 1;
    int count = comp, idx = -1;
    for (int i = 0; i < b.m->n_bxdf; ++i)
        if (bx_matches(b.m->bx[i], flags)) { if (count == 0) { idx = i; break; } --count; }
    P2 ur = mk2(pmin(u.x * (float)m - (float)comp, kOneMinusEps), u.y);
    V3 wo = bsdf_to_local(b, wo_w);
    if (wo.z == 0.0f) return none;
    BxDFSample s = bx_sample_f<KM>(b.m->bx[idx], wo, ur);
    if (s.pdf == 0.0f) return none;
    V3 wi_w = bsdf_to_world(b, s.wi);
    if (!(s.type & BSDF_SPECULAR) && m > 1)
        for (int i = 0; i < b.m->n_bxdf; ++i)
            if (i != idx && bx_matches(b.m->bx[i], flags)) s.pdf += bx_pdf<KM>(b.m->bx[i], wo, s.wi);
    if (m > 1) s.pdf /= (float)m;
    if (!(s.type & BSDF_SPECULAR)) {
        bool refl = dot(wi_w, b.ng) * dot(wo_w, b.ng) > 0.0f;
        s.f = rgb1(0.0f);
        for (int i = 0; i < b.m->n_bxdf; ++i) {
            const DBxDF& x = b.m->bx[i];
            if (bx_matches(x, flags) && ((refl && (x.type & BSDF_REFLECTION)) || (!refl && (x.type & BSDF_TRANSMISSION)))) s.f = s.f + bx_f<KM>(x, wo, s.wi);
        }
    }
    s.wi = wi_w;
    return s;
}

// ---- lights -------------------------------------------------------------------------
enum : int { LT_POINT = 0, LT_AREA = 1, LT_INFINITE = 2, LT_DISTANT = 3, LT_SPOT = 4, LT_GONIO = 5, LT_PROJECTION = 6 };
B2_D bool light_is_delta(int type) { return type == LT_POINT || type == LT_DISTANT || type == LT_SPOT || type == LT_GONIO || type == LT_PROJECTION; }  // light.rs: DELTA_POSITION | DELTA_DIRECTION
struct DLight {
    int type;
    int prim;        // area: original primitive index
    int two_sided;
    int inf_slot;    // infinite, goniometric: index into the DInfDistr table (goniometric: texels only)
    float pos[3];
    float area;      // area: Triangle::area (host, f32); spot: cos_total_width
    float L[3];
    float cos_falloff_start;  // spot
    float l2w[9];    // infinite: upper 3x3 of light_to_world (row-major); projection: inv_tan, then the screen window x0 y0 x1 y1
    float w2l[9];    // infinite, spot: upper 3x3 of world_to_light
};
// Environment map of an InfiniteAreaLight (host_envmap.cpp): level 0 of its MIPMap (float4 texels already multiplied by
// L; 1x1 without a "mapname") and the Distribution2D over the (2w x 2h) importance image (infinite.rs:326-369).
struct DInfDistr {
    const float4* texels;
    int width, height;
    int nu, nv;
    const float* func;      // nv x nu
    const float* cdf;       // nv x (nu + 1)
    const float* func_int;  // nv
    const float* mfunc;     // nv
    const float* mcdf;      // nv + 1
    float mfunc_int;
};

// core/src/pbrt/common.rs:251-276 over a cdf array of `size` entries
B2_D int find_interval_cdf(const float* cdf, int size, float u) {
    int first = 0, len = size;
    while (len > 0) {
        int half = len >> 1, middle = first + half;
        if (cdf[middle] <= u) { first = middle + 1; len -= half + 1; }
        else len = half;
    }
    if (first == 0) return 0;
    int v = first - 1;
    return v < 0 ? 0 : (v > size - 2 ? size - 2 : v);
}
// Distribution1D::sample_continuous, distribution_1d.rs:56-76
B2_D float distr_sample_continuous(const float* func, const float* cdf, float func_int, int n, float u, float* pdf, int* off) {
    int offset = find_interval_cdf(cdf, n + 1, u);
    float du = u - cdf[offset];
    if (cdf[offset + 1] - cdf[offset] > 0.0f) du /= cdf[offset + 1] - cdf[offset];
    *pdf = func_int > 0.0f ? func[offset] / func_int : 0.0f;
    *off = offset;
    return ((float)offset + du) / (float)n;
}

// l_map.lookup_triangle(st, 0.0): width 0 always selects MIPMap::triangle(0, st) (core/src/mipmap/mod.rs:226-247,
// 280-311): bilinear blend of four level-0 texels, ImageWrap::Repeat (texel(), :577-608).
B2_D int wrap_index(int a, int n) {  // pbrt::rem, common.rs:116-126
    int r = a - (a / n) * n;
    return r < 0 ? r + n : r;
}
B2_D RGB env_texel(const DInfDistr& D, int s, int t) {
    float4 q = __ldg(D.texels + (long long)wrap_index(t, D.height) * D.width + wrap_index(s, D.width));
    return rgb(q.x, q.y, q.z);
}
B2_D RGB inf_lookup(const DInfDistr& D, P2 st) {
    float s = st.x * (float)D.width - 0.5f, t = st.y * (float)D.height - 0.5f;
    float fs = floorf(s), ft = floorf(t);
    int s0 = (int)fs, t0 = (int)ft;
    float ds = s - (float)s0, dt = t - (float)t0;
    return env_texel(D, s0, t0) * (1.0f - ds) * (1.0f - dt) + env_texel(D, s0, t0 + 1) * (1.0f - ds) * dt + env_texel(D, s0 + 1, t0) * ds * (1.0f - dt) +
           env_texel(D, s0 + 1, t0 + 1) * ds * dt;
}
// Distribution2D::pdf, sampling/distribution_2d.rs (`as usize` saturates, then clamp)
B2_D float distr2d_pdf(const DInfDistr& D, float u, float v) {
    float fu = u * (float)D.nu, fv = v * (float)D.nv;
    int iu = (!(fu == fu) || fu <= 0.0f) ? 0 : (fu >= 2147483520.0f ? D.nu - 1 : (int)fu);
    int iv = (!(fv == fv) || fv <= 0.0f) ? 0 : (fv >= 2147483520.0f ? D.nv - 1 : (int)fv);
    iu = iu > D.nu - 1 ? D.nu - 1 : iu;
    iv = iv > D.nv - 1 ? D.nv - 1 : iv;
    return D.func[(long long)iv * D.nu + iu] / D.mfunc_int;
}
B2_D V3 xf3(const float* m, V3 v) {  // Transform::transform_vector, transform.rs:373-380
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
B2_D float spherical_theta(V3 v) { return acosf(pclamp(v.z, -1.0f, 1.0f)); }  // geometry/util.rs:41-43
B2_D float spherical_phi(V3 v) { float p = atan2f(v.y, v.x); return p < 0.0f ? p + kTwoPi : p; }  // :49-56
// InfiniteAreaLight::le, infinite.rs:188-199
B2_D RGB infinite_le(const DLight& l, const DInfDistr& D, V3 ray_d) {
    V3 w = normalize(xf3(l.w2l, ray_d));
    P2 st = mk2(spherical_phi(w) * kInvTwoPi, spherical_theta(w) * kInvPi);
    return inf_lookup(D, st);
}
// GonioPhotometricLight::scale (goniometric.rs:101-115): the map at the spherical coordinates of the light-space direction with y and z swapped
// (out of line, like projection_scale below: rare lights must not grow every shade kernel's instruction footprint)
__device__ __noinline__ RGB gonio_scale(const DLight& l, const DInfDistr& D, V3 w_world) {
    V3 wp = normalize(xf3(l.w2l, w_world));
    const float t = wp.y; wp.y = wp.z; wp.z = t;
    P2 st = mk2(spherical_phi(wp) * kInvTwoPi, spherical_theta(wp) * kInvPi);
    return inf_lookup(D, st);
}
// ProjectionLight::projection (projection.rs:115-139): Transform::perspective(fov, 1e-3, 1e30).transform_point is
// (inv_tan x, inv_tan y, ..) * (1 / z); outside the screen window or behind the near plane the light is black
__device__ __noinline__ RGB projection_scale(const DLight& l, const DInfDistr& D, V3 w_world) {
    const V3 wl = xf3(l.w2l, w_world);
    if (wl.z < 1e-3f) return rgb1(0.0f);
    const float inv_tan = l.l2w[0], x0 = l.l2w[1], y0 = l.l2w[2], x1 = l.l2w[3], y1 = l.l2w[4];
    float px = inv_tan * wl.x, py = inv_tan * wl.y;
    if (wl.z != 1.0f) { const float inv = 1.0f / wl.z; px = px * inv; py = py * inv; }
    if (!(px >= x0 && px <= x1 && py >= y0 && py <= y1)) return rgb1(0.0f);
    float ox = px - x0, oy = py - y0;  // Bounds2::offset
    if (x1 > x0) ox /= x1 - x0;
    if (y1 > y0) oy /= y1 - y0;
    return inf_lookup(D, mk2(ox, oy));
}
// DiffuseAreaLight::l, diffuse.rs:220-226
B2_D RGB area_l(const DLight& l, V3 n, V3 w) { return (l.two_sided || dot(n, w) > 0.0f) ? ldrgb(l.L) : rgb1(0.0f); }

// ---- sampler: HaltonSampler (samplers/src/halton.rs, core/src/low_discrepency.rs) ----
struct DHalton {
    const uint16_t* perms;      // compute_radical_inverse_permutations(RNG::default())
    const int* primes;          // first 1000 primes
    const int* prime_sums;
    const uint32_t* div_m;      // per prime: magic multiplier / shifts for exact u32 division (host_sampler.cpp)
    const uint32_t* div_sh;
    unsigned long long base_scale[2], base_exp[2], stride;
    long long mult_inv[2];
    int sample_at_center;
};
B2_D uint32_t reverse_bits_32(uint32_t n) { return __brev(n); }
B2_D float radical_inverse_base2(unsigned long long a) {  // low_discrepency.rs:454-460
    unsigned long long n0 = reverse_bits_32((uint32_t)a), n1 = reverse_bits_32((uint32_t)(a >> 32));
    unsigned long long rev = (n0 << 32) | n1;
    return __ull2float_rn(rev) * 0x1.0p-64f;
}
B2_D float radical_inverse_specialized(int base, unsigned long long a) {  // :401-420
    float inv_base = 1.0f / (float)base;
    unsigned long long rev = 0;
    float inv_base_n = 1.0f;
    if (a <= 0xffffffffull) {  // 32-bit fast path: same integer results
        uint32_t x = (uint32_t)a, b = (uint32_t)base;
        while (x != 0) { uint32_t next = x / b, digit = x - next * b; rev = rev * b + digit; inv_base_n *= inv_base; x = next; }
    } else {
        unsigned long long b = (unsigned long long)base;
        while (a != 0) { unsigned long long next = a / b, digit = a - next * b; rev = rev * b + digit; inv_base_n *= inv_base; a = next; }
    }
    return pmin(__ull2float_rn(rev) * inv_base_n, kOneMinusEps);
}
B2_D float scrambled_radical_inverse(int base, unsigned long long a, const uint16_t* perm, uint32_t dm, uint32_t dsh) {  // :428-448
    float inv_base = 1.0f / (float)base;
    unsigned long long rev = 0;
    float inv_base_n = 1.0f;
    if (a <= 0xffffffffull) {
        // same integer digits as the reference's u64 division; the quotient comes from a multiply-high (exact for every u32)
        uint32_t x = (uint32_t)a, b = (uint32_t)base, sh1 = dsh & 0xffu, sh2 = dsh >> 8;
        while (x != 0) {
            uint32_t t = __umulhi(dm, x);
            uint32_t next = (t + ((x - t) >> sh1)) >> sh2, digit = x - next * b;
            rev = rev * b + perm[digit]; inv_base_n *= inv_base; x = next;
        }
    } else {
        unsigned long long b = (unsigned long long)base;
        while (a != 0) { unsigned long long next = a / b, digit = a - next * b; rev = rev * b + perm[digit]; inv_base_n *= inv_base; a = next; }
    }
    float r = inv_base_n * (__ull2float_rn(rev) + inv_base * (float)perm[0] / (1.0f - inv_base));
    return pmin(r, kOneMinusEps);
}
B2_D unsigned long long inverse_radical_inverse(unsigned long long base, unsigned long long inverse, unsigned long long n_digits) {  // :1535-1545
    unsigned long long index = 0;
    for (unsigned long long i = 0; i < n_digits; ++i) { unsigned long long digit = inverse % base; inverse /= base; index = index * base + digit; }
    return index;
}
B2_D int rem_i(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }  // pbrt/common.rs:111-124
// HaltonSampler::get_index_for_sample, halton.rs:118-144
B2_D unsigned long long halton_index(const DHalton& h, int px, int py, unsigned long long sample_num) {
    unsigned long long offset = 0;
    if (h.stride > 1) {
        int pm[2] = {rem_i(px, 128), rem_i(py, 128)};
        for (int i = 0; i < 2; ++i) {
            unsigned long long dim_offset = inverse_radical_inverse(i == 0 ? 2 : 3, (unsigned long long)pm[i], h.base_exp[i]);
            offset += dim_offset * (h.stride / h.base_scale[i]) * (unsigned long long)h.mult_inv[i];
        }
        offset %= h.stride;
    }
    return offset + sample_num * h.stride;
}
// HaltonSampler::sample_dimension, halton.rs:150-160
B2_D float halton_dim(const DHalton& h, unsigned long long index, int dim) {
    if (h.sample_at_center && (dim == 0 || dim == 1)) return 0.5f;
    if (dim == 0) return radical_inverse_base2(index >> h.base_exp[0]);
    if (dim == 1) return radical_inverse_specialized(3, index / h.base_scale[1]);
    return scrambled_radical_inverse(h.primes[dim], index, h.perms + h.prime_sums[dim], h.div_m[dim], h.div_sh[dim]);
}

// ---- sampler: SobolSampler (samplers/src/sobol.rs, core/src/low_discrepency.rs:1770-1845) -------------------------
struct DSobol {
    const uint32_t* m32;                 // SOBOL_MATRICES_32: 1024 dimensions x 52
    const unsigned long long* vdc;       // VD_C_SOBOL_MATRICES[m - 1] (52), derived on the host
    const unsigned long long* vdc_inv;   // VD_C_SOBOL_MATRICES_INV[m - 1] (52)
    int log2_res, res;
    int sb_min[2];
};
B2_D float sobol_sample_f32(const DSobol& S, unsigned long long a, int dim) {  // sobol_sample_f32, scramble = 0
    uint32_t v = 0;
    const uint32_t* m = S.m32 + dim * 52;
    for (; a != 0; a >>= 1, ++m)
        if (a & 1ull) v ^= __ldg(m);
    return pmin(__uint2float_rn(v) * 0x1.0p-32f, kOneMinusEps);
}
B2_D unsigned long long sobol_interval_to_index(const DSobol& S, unsigned long long frame, int px, int py) {
    const int m = S.log2_res;
    if (m == 0) return 0ull;  // low_discrepency.rs:1771-1773
    unsigned long long index = frame << (2 * m), delta = 0ull;
    for (int c = 0; frame > 0; frame >>= 1, ++c)
        if (frame & 1ull) delta ^= S.vdc[c];
    unsigned long long b = ((((unsigned long long)(uint32_t)px) << m) | (unsigned long long)(uint32_t)py) ^ delta;
    for (int c = 0; b > 0; b >>= 1, ++c)
        if (b & 1ull) index ^= S.vdc_inv[c];
    return index;
}
// SobolSampler::sample_dimension for dimensions 0 / 1 (the film position inside pixel (px, py)), sobol.rs:64-80
B2_D float sobol_film_dim(const DSobol& S, unsigned long long index, int dim, int pixel) {
    float s = sobol_sample_f32(S, index, dim);
    s = s * (float)S.res + (float)S.sb_min[dim];
    return pclamp(s - (float)pixel, 0.0f, kOneMinusEps);
}

// ---- sampler: ZeroTwoSequenceSampler (samplers/src/zero_two_sequence.rs, core/src/sampler/pixel_sampler.rs) ------
// The reference generates, per pixel and per 1-D / 2-D slot, spp gray-code samples of a scrambled van der Corput /
// Sobol' (0,2) sequence and shuffles them with the TILE's PCG32 stream (low_discrepency.rs:1676-1763).  A prepass
// (k_zerotwo_tiles, one thread per 16x16 tile, replaying that stream in pixel order) stores per (pixel, slot) the
// scramble(s) and the shuffle as a source-index permutation; a sample is then scramble ^ G * gray(perm[s]).
struct DPcg32;
struct DZeroTwo {
    const uint32_t* scr1;   // [pixel][n1]
    const uint16_t* perm1;  // [pixel][n1][spp]
    const uint32_t* scr2;   // [pixel][n2][2]
    const uint16_t* perm2;  // [pixel][n2][spp]
    int n1, n2, spp;
    // Tile-sequential mode ("dimensions" below the path's worst case, e.g. the reference's default 4): `pixel` above is
    // the path's slot (one path in flight per tile, tables of the tile's current pixel only) and requests past the
    // n1 / n2 pre-generated slots draw from the tile's PCG32 stream (pixel_sampler.rs:88-110).  Null otherwise.
    DPcg32* rng;
};
__device__ __constant__ uint32_t kCSobol1[32] = {
    0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000, 0x80800000, 0xc0c00000, 0xa0a00000,
    0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000, 0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00,
    0xaa00aa00, 0xff00ff00, 0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff};
B2_D float u32_to_unit(uint32_t v) { return pmin(__uint2float_rn(v) * 0x1.0p-32f, kOneMinusEps); }  // low_discrepency.rs:1603-1607
// value after i gray-code steps from `scramble`: scramble ^ (G * gray(i))
B2_D uint32_t gray_vdc(uint32_t i) { return __brev(i ^ (i >> 1)); }  // C_VANDER_CORPUT[b] = 1 << (31 - b)
B2_D uint32_t gray_sobol1(uint32_t i) {
    uint32_t g = i ^ (i >> 1), v = 0;
    while (g) { int b = __ffs(g) - 1; v ^= kCSobol1[b]; g &= g - 1; }
    return v;
}
B2_D float zt_1d(const DZeroTwo& z, long long pix, int slot, int s) {
    uint32_t scr = z.scr1[pix * z.n1 + slot];
    uint32_t j = z.perm1[(pix * z.n1 + slot) * z.spp + s];
    return u32_to_unit(scr ^ gray_vdc(j));
}
B2_D P2 zt_2d(const DZeroTwo& z, long long pix, int slot, int s) {
    const uint32_t* scr = z.scr2 + (pix * z.n2 + slot) * 2;
    uint32_t j = z.perm2[(pix * z.n2 + slot) * z.spp + s];
    return mk2(u32_to_unit(scr[0] ^ gray_vdc(j)), u32_to_unit(scr[1] ^ gray_sobol1(j)));
}

// core/src/rng.rs PCG32 (device copy for the (0,2) prepass)
struct DPcg32 {
    unsigned long long state, inc;
};
B2_D uint32_t pcg_next(DPcg32& r) {
    unsigned long long old = r.state;
    r.state = old * 0x5851f42d4c957f2dULL + r.inc;
    uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27), rot = (uint32_t)(old >> 59);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
}
B2_D void pcg_set_sequence(DPcg32& r, unsigned long long seq) {  // rng.rs:39-59
    r.state = 0; r.inc = (seq << 1) | 1ULL;
    (void)pcg_next(r);
    r.state += 0x853c49e6748fea9bULL;
    (void)pcg_next(r);
}
B2_D uint32_t pcg_bounded(DPcg32& r, uint32_t b) {  // rng.rs:86-96, lower bound 0
    uint32_t threshold = (~b + 1u) % b;
    for (;;) { uint32_t v = pcg_next(r); if (v >= threshold) return v % b; }
}

// ---- camera: PerspectiveCamera::generate_ray_differential without differentials ------
struct DCamera {
    float r2c[16], c2w[16];
    float lens_radius, focal_distance, shutter_open, shutter_close;
    int type;        // B200PT_CAMERA_*
    int xres, yres;  // film.full_resolution (environment camera)
};
B2_D V3 xf_point(const float* m, V3 p) {  // transform.rs:288-302
    float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1.0f) return mk(xp, yp, zp);
    return mk(xp, yp, zp) / wp;
}
// cameras/src/perspective_camera.rs:144-204 + Transform::transform_ray (transform.rs:451-476)
// + cameras/src/orthographic_camera.rs:64-93 (origin on the film plane, direction +z) and environment_camera.rs:38-53
// (direction from the film position in spherical coordinates; sin / cos are the exact glibc ones, libm_exact.cuh)
B2_D Ray32 camera_ray(const DCamera& c, P2 p_film, float time_u, P2 p_lens) {
    V3 p_camera = xf_point(c.r2c, mk(p_film.x, p_film.y, 0.0f));
    V3 o = mk(0.0f, 0.0f, 0.0f), d;
    if (c.type == B200PT_CAMERA_ENVIRONMENT) {
        const float theta = kPi * p_film.y / (float)c.yres;
        const float phi = (kPi * 2.0f) * p_film.x / (float)c.xres;
        const float st = lmx::sinf_glibc(theta), ct = lmx::cosf_glibc(theta), sp = lmx::sinf_glibc(phi), cp = lmx::cosf_glibc(phi);
        d = mk(st * cp, ct, st * sp);
    } else if (c.type == B200PT_CAMERA_ORTHOGRAPHIC) { o = p_camera; d = mk(0.0f, 0.0f, 1.0f); }
    else d = normalize(p_camera);
    float time = (1.0f - time_u) * c.shutter_open + time_u * c.shutter_close;
    if (c.lens_radius > 0.0f) {
        P2 cd = concentric_sample_disk(p_lens);
        P2 pl = mk2(c.lens_radius * cd.x, c.lens_radius * cd.y);
        float ft = c.focal_distance / d.z;
        V3 p_focus = o + d * ft;
        o = mk(pl.x, pl.y, 0.0f);
        d = normalize(p_focus - o);
    }
    // transform_point_with_error (transform.rs:307-331) + transform_vector
    const float* m = c.c2w;
    float x = o.x, y = o.y, z = o.z;
    float xp = (m[0] * x + m[1] * y) + (m[2] * z + m[3]);
    float yp = (m[4] * x + m[5] * y) + (m[6] * z + m[7]);
    float zp = (m[8] * x + m[9] * y) + (m[10] * z + m[11]);
    float wp = (m[12] * x + m[13] * y) + (m[14] * z + m[15]);
    float xs = pabs(m[0] * x) + pabs(m[1] * y) + pabs(m[2] * z) + pabs(m[3]);
    float ys = pabs(m[4] * x) + pabs(m[5] * y) + pabs(m[6] * z) + pabs(m[7]);
    float zs = pabs(m[8] * x) + pabs(m[9] * y) + pabs(m[10] * z) + pabs(m[11]);
    V3 o_err = kGamma3 * mk(xs, ys, zs);
    V3 ow = (wp == 1.0f) ? mk(xp, yp, zp) : mk(xp, yp, zp) / wp;
    V3 dw = mk(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z, m[8] * d.x + m[9] * d.y + m[10] * d.z);
    float l2 = length_squared(dw);
    float t_max = __int_as_float(0x7f800000);
    if (l2 > 0.0f) {
        float dt = dot(vabs(dw), o_err) / l2;
        ow = ow + dw * dt;
        t_max -= dt;
    }
    Ray32 r;
    r.ox = ow.x; r.oy = ow.y; r.oz = ow.z; r.tmax = t_max;
    r.dx = dw.x; r.dy = dw.y; r.dz = dw.z; r.time = time;
    return r;
}

// ---- hit geometry: Triangle::intersect tail (shapes/src/triangle.rs:547-725) -----------
struct SurfHit {
    V3 p, p_error, n;   // Hit::{p, p_error, n}
    V3 ns, dpdu;        // Shading::{n, dpdu} (== n and der.dpdu unless the mesh has N / S)
};
// duv = {uv0 - uv2, uv1 - uv2}; nrm / tan: the three vertex normals / tangents (9 floats) or null.
// flags: B200PT_PRIM_FLIP_NORMAL, B200PT_PRIM_REVERSE_ORIENTATION.
B2_D SurfHit triangle_surface(V3 p0, V3 p1, V3 p2, float b0, float b1, float b2, uint32_t flags, float4 duv, const float* nrm, const float* tan) {
    SurfHit s;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float duv02x = duv.x, duv02y = duv.y, duv12x = duv.z, duv12y = duv.w;
    float determinant = duv02x * duv12y - duv02y * duv12x;
    bool degenerate_uv = pabs(determinant) < 1e-8f;
    V3 dpdu = mk(0.0f, 0.0f, 0.0f), dpdv = dpdu;
    if (!degenerate_uv) {
        float invdet = 1.0f / determinant;
        dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        dpdv = (-duv12x * dp02 + duv02x * dp12) * invdet;
    }
    if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
        V3 ng = cross(p2 - p0, p1 - p0);
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    float xs = pabs(b0 * p0.x) + pabs(b1 * p1.x) + pabs(b2 * p2.x);
    float ys = pabs(b0 * p0.y) + pabs(b1 * p1.y) + pabs(b2 * p2.y);
    float zs = pabs(b0 * p0.z) + pabs(b1 * p1.z) + pabs(b2 * p2.z);
    s.p_error = kGamma7 * mk(xs, ys, zs);
    s.p = b0 * p0 + b1 * p1 + b2 * p2;
    V3 n = normalize(cross(dp02, dp12));
    if (flags & 1u) n = -n;
    s.n = n;
    s.ns = n;
    s.dpdu = dpdu;
    if (nrm || tan) {  // triangle.rs:631-721
        V3 ns = n;
        if (nrm) {
            V3 ns2 = b0 * mk(nrm[0], nrm[1], nrm[2]) + b1 * mk(nrm[3], nrm[4], nrm[5]) + b2 * mk(nrm[6], nrm[7], nrm[8]);
            if (length_squared(ns2) > 0.0f) ns = normalize(ns2);
        }
        V3 ss = normalize(dpdu);
        if (tan) {
            V3 ss2 = b0 * mk(tan[0], tan[1], tan[2]) + b1 * mk(tan[3], tan[4], tan[5]) + b2 * mk(tan[6], tan[7], tan[8]);
            if (length_squared(ss2) > 0.0f) ss = normalize(ss2);
        }
        V3 ts = cross(ss, ns);
        if (length_squared(ts) > 0.0f) {
            ts = normalize(ts);
            ss = cross(ts, ns);
        } else {
            coordinate_system(ns, &ss, &ts);
        }
        if (flags & 8u) ts = -ts;
        // set_shading_geometry(ss, ts, .., orientation_is_authoritative = true), surface_interaction.rs:152-173
        s.ns = normalize(cross(ss, ts));
        s.n = face_forward(s.n, s.ns);
        s.dpdu = ss;
    }
    return s;
}
B2_D float4 default_duv() { return make_float4(0.0f - 1.0f, 0.0f - 1.0f, 1.0f - 1.0f, 0.0f - 1.0f); }

}  // namespace b2
