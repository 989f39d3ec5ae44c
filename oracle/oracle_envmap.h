// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle_math.h).
//
// Image-mapped InfiniteAreaLight: the MIPMap the reference builds over the environment image and the lookups the
// light makes through it.  Restates
//   MIPMap::new / lookup_triangle / triangle / texel   core/src/mipmap/mod.rs:121-186, 226-311, 577-608
//   resample_image / resample_weights                  core/src/mipmap/mod.rs:373-575
//   lanczos                                            core/src/texture/common.rs:216-228
//   rem                                                core/src/pbrt/common.rs:116-126
// for T = RGBSpectrum, ImageWrap::Repeat (the only instantiation on this path, lights/src/infinite.rs:81).
// The blocked memory layout of BlockedArray is not observable and is not reproduced.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "oracle_math.h"

namespace orc {

struct MipLevel {
    int w = 0, h = 0;
    std::vector<RGB> px;  // row-major, (s, t) -> px[t * w + s]
    RGB at(int s, int t) const { return px[(size_t)t * w + s]; }
};

// pbrt::rem on isize
inline int64_t irem(int64_t a, int64_t b) {
    int64_t r = a - (a / b) * b;
    return r < 0 ? r + b : r;
}
// texture/common.rs:216-228
inline Float lanczos(Float x, Float tau) {
    x = pabs(x);
    if (x < 1e-5f) return 1.0f;
    if (x > 1.0f) return 0.0f;
    x *= kPi;
    Float s = std::sin(x * tau) / (x * tau);
    Float l = std::sin(x) / x;
    return s * l;
}
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct ResampleWeight {
    size_t first_texel;
    Float weight[4];
};
// mod.rs:542-570.  `(.. ).floor() as usize` saturates at 0 for the first texels of a row (Rust float -> usize cast),
// so their four taps start at texel 0 instead of wrapping around.
inline std::vector<ResampleWeight> resample_weights(int old_res, int new_res) {
    std::vector<ResampleWeight> wt((size_t)new_res);
    const Float filterwidth = 2.0f;
    for (int i = 0; i < new_res; ++i) {
        Float center = ((Float)i + 0.5f) * (Float)old_res / (Float)new_res;
        Float f = std::floor((center - filterwidth) + 0.5f);
        wt[(size_t)i].first_texel = f <= 0.0f ? 0 : (size_t)f;
        for (int j = 0; j < 4; ++j) {
            Float pos = (Float)wt[(size_t)i].first_texel + (Float)j + 0.5f;
            wt[(size_t)i].weight[j] = lanczos((pos - center) / filterwidth, 2.0f);
        }
        Float inv = 1.0f / (wt[(size_t)i].weight[0] + wt[(size_t)i].weight[1] + wt[(size_t)i].weight[2] + wt[(size_t)i].weight[3]);
        for (int j = 0; j < 4; ++j) wt[(size_t)i].weight[j] *= inv;
    }
    return wt;
}

struct MipMap {
    std::vector<MipLevel> pyramid;
    int width() const { return pyramid[0].w; }
    int height() const { return pyramid[0].h; }
    int levels() const { return (int)pyramid.size(); }

    // mod.rs:577-608, ImageWrap::Repeat
    RGB texel(int level, int64_t s, int64_t t) const {
        const MipLevel& l = pyramid[(size_t)level];
        return l.at((int)irem(s, l.w), (int)irem(t, l.h));
    }

    // mod.rs:121-186 (+ resample_image :373-540 when a side is not a power of two)
    void build(int w, int h, const std::vector<RGB>& img) {
        MipLevel l0;
        if (!is_pow2(w) || !is_pow2(h)) {
            const int pw = next_pow2(w), ph = next_pow2(h);
            l0.w = pw; l0.h = ph;
            l0.px.assign((size_t)pw * ph, RGB());
            std::vector<ResampleWeight> sw = resample_weights(w, pw);
            for (int t = 0; t < h; ++t)
                for (int s = 0; s < pw; ++s) {
                    RGB pixel;
                    for (int j = 0; j < 4; ++j) {
                        size_t o = (size_t)irem((int64_t)(sw[(size_t)s].first_texel + j), w);
                        if (o < (size_t)w) pixel += img[(size_t)t * w + o] * sw[(size_t)s].weight[j];
                    }
                    l0.px[(size_t)t * pw + s] += pixel;
                }
            std::vector<ResampleWeight> tw = resample_weights(h, ph);
            std::vector<RGB> work((size_t)ph);
            for (int s = 0; s < pw; ++s) {
                for (int t = 0; t < ph; ++t) {
                    work[(size_t)t] = RGB();
                    for (int j = 0; j < 4; ++j) {
                        size_t o = (size_t)irem((int64_t)(tw[(size_t)t].first_texel + j), h);
                        if (o < (size_t)h) work[(size_t)t] += l0.px[o * pw + s] * tw[(size_t)t].weight[j];
                    }
                }
                for (int t = 0; t < ph; ++t) l0.px[(size_t)t * pw + s] = rgb_clamp0(work[(size_t)t]);
            }
        } else {
            l0.w = w; l0.h = h; l0.px = img;
        }
        pyramid.clear();
        pyramid.push_back(l0);
        int mx = l0.w > l0.h ? l0.w : l0.h, n_levels = 1;
        while ((1 << n_levels) <= mx) ++n_levels;  // 1 + log2int(max)
        for (int i = 1; i < n_levels; ++i) {
            MipLevel l;
            l.w = pyramid[(size_t)i - 1].w / 2 > 1 ? pyramid[(size_t)i - 1].w / 2 : 1;
            l.h = pyramid[(size_t)i - 1].h / 2 > 1 ? pyramid[(size_t)i - 1].h / 2 : 1;
            l.px.resize((size_t)l.w * l.h);
            pyramid.push_back(l);
            for (int t = 0; t < l.h; ++t)
                for (int s = 0; s < l.w; ++s)
                    pyramid[(size_t)i].px[(size_t)t * l.w + s] =
                        (texel(i - 1, 2 * s, 2 * t) + texel(i - 1, 2 * s + 1, 2 * t) + texel(i - 1, 2 * s, 2 * t + 1) + texel(i - 1, 2 * s + 1, 2 * t + 1)) * 0.25f;
        }
    }

    // mod.rs:280-311
    RGB triangle(int level, P2 st) const {
        level = level < 0 ? 0 : (level > levels() - 1 ? levels() - 1 : level);
        const MipLevel& l = pyramid[(size_t)level];
        Float s = st.x * (Float)l.w - 0.5f, t = st.y * (Float)l.h - 0.5f;
        Float fs = std::floor(s), ft = std::floor(t);
        int64_t s0 = (int64_t)fs, t0 = (int64_t)ft;
        Float ds = s - (Float)s0, dt = t - (Float)t0;
        return texel(level, s0, t0) * (1.0f - ds) * (1.0f - dt) + texel(level, s0, t0 + 1) * (1.0f - ds) * dt + texel(level, s0 + 1, t0) * ds * (1.0f - dt) +
               texel(level, s0 + 1, t0 + 1) * ds * dt;
    }
    // mod.rs:226-247
    RGB lookup_triangle(P2 st, Float width) const {
        const int n = levels();
        Float level = (Float)n - 1.0f + std::log2(pmax(width, 1e-8f));
        if (level < 0.0f) return triangle(0, st);
        if (level >= (Float)(n - 1)) return texel(n - 1, 0, 0);
        int il = (int)std::floor(level);
        Float delta = level - (Float)il;
        return triangle(il, st) * (1.0f - delta) + triangle(il + 1, st) * delta;
    }
};

}  // namespace orc
