// See host_envmap.h.  f32 throughout, built with -ffp-contract=off like the rest of the host code: the pyramid, the
// importance image and the CDFs have to come out bit-identical to what the reference computes at scene load.
#include "host_envmap.h"

#include <cmath>
#include <cstddef>

namespace b2host {
namespace {

struct Px { float r, g, b; };
inline Px operator+(Px a, Px b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline Px operator*(Px a, float f) { return {a.r * f, a.g * f, a.b * f}; }
inline float lum(Px a) { return 0.212671f * a.r + 0.715160f * a.g + 0.072169f * a.b; }  // rgb_spectrum.rs:131

struct Level {
    int w, h;
    std::vector<Px> px;
};

// pbrt::rem (common.rs:116-126): remainder that is never negative
inline long wrap(long a, long n) {
    long r = a - (a / n) * n;
    return r < 0 ? r + n : r;
}

struct Pyramid {
    std::vector<Level> lv;
    // texel(), mipmap/mod.rs:577-608 with ImageWrap::Repeat
    Px texel(int level, long s, long t) const {
        const Level& l = lv[(size_t)level];
        return l.px[(size_t)wrap(t, l.h) * (size_t)l.w + (size_t)wrap(s, l.w)];
    }
    // triangle(), mod.rs:280-311
    Px bilinear(int level, float u, float v) const {
        const int n = (int)lv.size();
        level = level < 0 ? 0 : (level > n - 1 ? n - 1 : level);
        const Level& l = lv[(size_t)level];
        float s = u * (float)l.w - 0.5f, t = v * (float)l.h - 0.5f;
        long s0 = (long)std::floor(s), t0 = (long)std::floor(t);
        float ds = s - (float)s0, dt = t - (float)t0;
        return texel(level, s0, t0) * (1.0f - ds) * (1.0f - dt) + texel(level, s0, t0 + 1) * (1.0f - ds) * dt + texel(level, s0 + 1, t0) * ds * (1.0f - dt) +
               texel(level, s0 + 1, t0 + 1) * ds * dt;
    }
    // lookup_triangle(), mod.rs:226-247
    Px trilinear(float u, float v, float width) const {
        const int n = (int)lv.size();
        float w = width > 1e-8f ? width : 1e-8f;
        float level = (float)n - 1.0f + std::log2(w);
        if (level < 0.0f) return bilinear(0, u, v);
        if (level >= (float)(n - 1)) return texel(n - 1, 0, 0);
        int il = (int)std::floor(level);
        float d = level - (float)il;
        return bilinear(il, u, v) * (1.0f - d) + bilinear(il + 1, u, v) * d;
    }
};

// texture/common.rs:216-228
float lanczos2(float x) {
    const float tau = 2.0f, pi = 3.14159265358979323846f;
    x = x < 0.0f ? -x : x;
    if (x < 1e-5f) return 1.0f;
    if (x > 1.0f) return 0.0f;
    x *= pi;
    float s = std::sin(x * tau) / (x * tau);
    float l = std::sin(x) / x;
    return s * l;
}

struct Taps {
    size_t first;
    float w[4];
};
// resample_weights(), mod.rs:542-570.  The float -> usize cast saturates at zero.
std::vector<Taps> taps_for(int from, int to) {
    std::vector<Taps> out((size_t)to);
    for (int i = 0; i < to; ++i) {
        Taps& k = out[(size_t)i];
        float center = ((float)i + 0.5f) * (float)from / (float)to;
        float f = std::floor((center - 2.0f) + 0.5f);
        k.first = f > 0.0f ? (size_t)f : 0;
        for (int j = 0; j < 4; ++j) k.w[j] = lanczos2((((float)k.first + (float)j + 0.5f) - center) / 2.0f);
        float inv = 1.0f / (k.w[0] + k.w[1] + k.w[2] + k.w[3]);
        for (int j = 0; j < 4; ++j) k.w[j] *= inv;
    }
    return out;
}

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int up_pow2(int v) { int p = 1; while (p < v) p *= 2; return p; }
inline float clamp0(float v) { return v < 0.0f ? 0.0f : v; }  // clamp_default: [0, inf)

// resample_image(), mod.rs:373-540: zoom in s, then in t (in place, column by column), clamp at zero
Level zoom_to_pow2(const std::vector<Px>& img, int w, int h) {
    Level o;
    o.w = up_pow2(w); o.h = up_pow2(h);
    o.px.assign((size_t)o.w * o.h, Px{0, 0, 0});
    const std::vector<Taps> ks = taps_for(w, o.w);
    for (int t = 0; t < h; ++t)
        for (int s = 0; s < o.w; ++s) {
            Px acc{0, 0, 0};
            for (int j = 0; j < 4; ++j) {
                size_t src = (size_t)wrap((long)(ks[(size_t)s].first + j), w);
                if (src < (size_t)w) acc = acc + img[(size_t)t * w + src] * ks[(size_t)s].w[j];
            }
            Px& dst = o.px[(size_t)t * o.w + s];
            dst = dst + acc;
        }
    const std::vector<Taps> kt = taps_for(h, o.h);
    std::vector<Px> col((size_t)o.h);
    for (int s = 0; s < o.w; ++s) {
        for (int t = 0; t < o.h; ++t) {
            Px acc{0, 0, 0};
            for (int j = 0; j < 4; ++j) {
                size_t src = (size_t)wrap((long)(kt[(size_t)t].first + j), h);
                if (src < (size_t)h) acc = acc + o.px[src * o.w + s] * kt[(size_t)t].w[j];
            }
            col[(size_t)t] = acc;
        }
        for (int t = 0; t < o.h; ++t) o.px[(size_t)t * o.w + s] = Px{clamp0(col[(size_t)t].r), clamp0(col[(size_t)t].g), clamp0(col[(size_t)t].b)};
    }
    return o;
}

// Distribution1D::new, sampling/distribution_1d.rs:22-48
void make_cdf(const float* f, int n, float* cdf, float* integral) {
    cdf[0] = 0.0f;
    for (int i = 1; i <= n; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / (float)n;
    *integral = cdf[n];
    if (*integral == 0.0f) { for (int i = 1; i <= n; ++i) cdf[i] = (float)i / (float)n; }
    else { for (int i = 1; i <= n; ++i) cdf[i] /= *integral; }
}

}  // namespace

void build_envmap(const float* rgb, int mw, int mh, const float L[3], EnvMapTables* out, bool importance) {
    // InfiniteAreaLight::new, infinite.rs:66-81: texels = image * L, or the 1x1 image [L]
    std::vector<Px> img;
    int w = 1, h = 1;
    if (rgb && mw > 0 && mh > 0) {
        w = mw; h = mh;
        img.resize((size_t)w * h);
        for (size_t k = 0; k < img.size(); ++k) img[k] = Px{rgb[3 * k] * L[0], rgb[3 * k + 1] * L[1], rgb[3 * k + 2] * L[2]};
    } else img.push_back(Px{L[0], L[1], L[2]});

    // MIPMap::new, mod.rs:121-186
    Pyramid P;
    if (!pow2(w) || !pow2(h)) P.lv.push_back(zoom_to_pow2(img, w, h));
    else P.lv.push_back(Level{w, h, img});
    const int longest = P.lv[0].w > P.lv[0].h ? P.lv[0].w : P.lv[0].h;
    int n_levels = 1;
    while ((1 << n_levels) <= longest) ++n_levels;
    for (int i = 1; i < n_levels; ++i) {
        const int pw = P.lv[(size_t)i - 1].w, ph = P.lv[(size_t)i - 1].h;
        Level l{pw / 2 > 1 ? pw / 2 : 1, ph / 2 > 1 ? ph / 2 : 1, {}};
        l.px.resize((size_t)l.w * l.h);
        P.lv.push_back(l);
        for (int t = 0; t < l.h; ++t)
            for (int s = 0; s < l.w; ++s)
                P.lv[(size_t)i].px[(size_t)t * l.w + s] =
                    (P.texel(i - 1, 2 * s, 2 * t) + P.texel(i - 1, 2 * s + 1, 2 * t) + P.texel(i - 1, 2 * s, 2 * t + 1) + P.texel(i - 1, 2 * s + 1, 2 * t + 1)) * 0.25f;
    }

    out->width = P.lv[0].w;
    out->height = P.lv[0].h;
    out->texels.resize((size_t)out->width * out->height * 4);
    for (size_t k = 0; k < P.lv[0].px.size(); ++k) {
        out->texels[4 * k] = P.lv[0].px[k].r; out->texels[4 * k + 1] = P.lv[0].px[k].g; out->texels[4 * k + 2] = P.lv[0].px[k].b; out->texels[4 * k + 3] = 0.0f;
    }

    {
        Px pw = P.trilinear(0.5f, 0.5f, 0.5f);
        out->power_lookup[0] = pw.r; out->power_lookup[1] = pw.g; out->power_lookup[2] = pw.b;
    }
    if (!importance) { out->nu = out->nv = 0; return; }
    // compute_scalar_image + Distribution2D::new, infinite.rs:326-369, sampling/distribution_2d.rs
    const int nu = 2 * out->width, nv = 2 * out->height;
    const float fwidth = 0.5f / (float)(nu < nv ? nu : nv);
    const float pi = 3.14159265358979323846f;
    out->nu = nu; out->nv = nv;
    out->cond_func.resize((size_t)nu * nv);
    out->cond_cdf.resize((size_t)(nu + 1) * nv);
    out->cond_int.resize((size_t)nv);
    for (int v = 0; v < nv; ++v) {
        float vp = ((float)v + 0.5f) / (float)nv;
        float sin_theta = std::sin(pi * ((float)v + 0.5f) / (float)nv);
        float* row = &out->cond_func[(size_t)v * nu];
        for (int u = 0; u < nu; ++u) {
            float up = ((float)u + 0.5f) / (float)nu;
            row[u] = lum(P.trilinear(up, vp, fwidth)) * sin_theta;
        }
        make_cdf(row, nu, &out->cond_cdf[(size_t)v * (nu + 1)], &out->cond_int[(size_t)v]);
    }
    out->marg_func = out->cond_int;
    out->marg_cdf.resize((size_t)nv + 1);
    make_cdf(out->marg_func.data(), nv, out->marg_cdf.data(), &out->marg_int);
}

}  // namespace b2host
