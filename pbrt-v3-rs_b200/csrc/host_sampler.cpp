// Host-side sampler tables for the device HaltonSampler
// (samplers/src/halton.rs:16-19,61-100; core/src/low_discrepency.rs:9-376,1512-1528;
// core/src/rng.rs:14-120).  Built once per process, uploaded per scene.
#include "host_sampler.h"

#include <mutex>

namespace b2host {

namespace {
struct Pcg32 {  // core/src/rng.rs: RNG::default()
    uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
    uint32_t next() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18) ^ old) >> 27), rot = (uint32_t)(old >> 59);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31u));
    }
    uint32_t bounded(uint32_t b) {  // rng.rs:86-96 with lower bound 0
        uint32_t threshold = (~b + 1u) % b;
        for (;;) {
            uint32_t r = next();
            if (r >= threshold) return r % b;
        }
    }
};
}  // namespace

const HaltonTables& halton_tables() {
    static HaltonTables t;
    static std::once_flag once;
    std::call_once(once, [] {
        // first 1000 primes and their running sums
        const int N = 8000;
        std::vector<char> comp(N, 0);
        for (int i = 2; i < N && (int)t.primes.size() < 1000; ++i) {
            if (comp[i]) continue;
            t.primes.push_back(i);
            for (int j = 2 * i; j < N; j += i) comp[j] = 1;
        }
        int s = 0;
        for (int p : t.primes) { t.prime_sums.push_back(s); s += p; }
        for (int p : t.primes) {
            int l = 0;
            while ((1u << l) < (uint32_t)p) ++l;
            uint64_t m = ((uint64_t(1) << 32) * ((uint64_t(1) << l) - (uint64_t)p)) / (uint64_t)p + 1;
            t.div_m.push_back((uint32_t)m);
            t.div_sh.push_back((uint32_t)(l < 1 ? l : 1) | ((uint32_t)(l - 1 > 0 ? l - 1 : 0) << 8));
        }
        // compute_radical_inverse_permutations: identity per base, Fisher-Yates with RNG::shuffle (rng.rs:106-119)
        t.perms.resize((size_t)s);
        Pcg32 rng;
        size_t off = 0;
        for (int p : t.primes) {
            for (int j = 0; j < p; ++j) t.perms[off + j] = (uint16_t)j;
            for (int i = 0; i < p; ++i) {
                int other = i + (int)rng.bounded((uint32_t)(p - i));
                uint16_t tmp = t.perms[off + i]; t.perms[off + i] = t.perms[off + other]; t.perms[off + other] = tmp;
            }
            off += p;
        }
    });
    return t;
}

static void ext_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {  // halton.rs:286-294
    if (b == 0) { *x = 1; *y = 0; return; }
    int64_t d = (int64_t)(a / b), xp, yp;
    ext_gcd(b, a % b, &xp, &yp);
    *x = yp;
    *y = xp - d * yp;
}
static uint64_t mult_inverse(int64_t a, int64_t n) {  // halton.rs:296-299 (pbrt Mod)
    int64_t x, y;
    ext_gcd((uint64_t)a, (uint64_t)n, &x, &y);
    int64_t r = x - (x / n) * n;
    if (r < 0) r += n;
    return (uint64_t)r;
}

HaltonParams halton_params(int res_x, int res_y) {  // HaltonSampler::new, halton.rs:61-100
    HaltonParams h;
    int res[2] = {res_x, res_y};
    for (int i = 0; i < 2; ++i) {
        uint64_t base = i == 0 ? 2 : 3, scale = 1, e = 0;
        int lim = res[i] < 128 ? res[i] : 128;  // K_MAX_RESOLUTION
        while ((int)scale < lim) { scale *= base; e += 1; }
        h.base_scale[i] = scale;
        h.base_exp[i] = e;
    }
    h.stride = h.base_scale[0] * h.base_scale[1];
    h.mult_inv[0] = (int64_t)mult_inverse((int64_t)h.base_scale[1], (int64_t)h.base_scale[0]);
    h.mult_inv[1] = (int64_t)mult_inverse((int64_t)h.base_scale[0], (int64_t)h.base_scale[1]);
    return h;
}

}  // namespace b2host
