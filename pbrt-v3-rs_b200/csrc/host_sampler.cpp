// Host-side sampler tables for the device HaltonSampler
// (samplers/src/halton.rs:16-19,61-100; core/src/low_discrepency.rs:9-376,1512-1528;
// core/src/rng.rs:14-120).  Built once per process, uploaded per scene.
#include <algorithm>
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


// ---- SobolSampler: the pixel <-> index tables of sobol_interval_to_index (core/src/low_discrepency.rs:1770-1808) ----
// The reference reads them from VD_C_SOBOL_MATRICES[m - 1] / VD_C_SOBOL_MATRICES_INV[m - 1]; here they are derived from the
// generator matrices of dimensions 0 and 1 (the first 2 x 52 entries of SOBOL_MATRICES_32).  Bit j of a sample index
// toggles, in the 2m-bit word (pixel x << m | pixel y) of a 2^m x 2^m image, the pattern
//     e(j) = (M0[j] >> (32 - m)) << m  |  M1[j] >> (32 - m).
// vdc[c] = e(2m + c): what bit c of the sample number does to the pixel.  vdc_inv[c] = the low index bits that
// produce pixel-word bit c alone, i.e. column c of E^-1 with E = [e(0) .. e(2m - 1)], found by reducing [E | I] over
// GF(2) with the basis kept as (pattern, combination) pairs.
namespace b2host {

bool sobol_interval_tables(const uint32_t* m32, int m, uint64_t vdc[52], uint64_t vdc_inv[52]) {
    for (int c = 0; c < 52; ++c) vdc[c] = vdc_inv[c] = 0;
    if (m <= 0 || m > 26) return m == 0;
    const uint32_t* M0 = m32;
    const uint32_t* M1 = m32 + 52;
    auto e = [&](int j) -> uint64_t { return j < 52 ? (((uint64_t)(M0[j] >> (32 - m))) << m) | (uint64_t)(M1[j] >> (32 - m)) : 0ull; };
    const int n = 2 * m;
    for (int c = 0; c + n < 52; ++c) vdc[c] = e(n + c);
    // pat[k]: a pixel-word pattern; comb[k]: the set of index bits whose e() XOR to it
    uint64_t pat[52], comb[52];
    for (int j = 0; j < n; ++j) { pat[j] = e(j); comb[j] = 1ull << j; }
    for (int bit = 0; bit < n; ++bit) {  // make pat[bit] the only pattern with `bit` set, and pat[bit] == 1 << bit in the end
        int pivot = -1;
        for (int k = bit; k < n; ++k) if ((pat[k] >> bit) & 1ull) { pivot = k; break; }
        if (pivot < 0) return false;  // singular: cannot happen for a (0, 2)-sequence in base 2
        std::swap(pat[bit], pat[pivot]);
        std::swap(comb[bit], comb[pivot]);
        for (int k = 0; k < n; ++k)
            if (k != bit && ((pat[k] >> bit) & 1ull)) { pat[k] ^= pat[bit]; comb[k] ^= comb[bit]; }
    }
    for (int c = 0; c < n; ++c) vdc_inv[c] = comb[c];
    return true;
}

}  // namespace b2host

// exposed for the tests (include/b200pt.h)
extern "C" int b200pt_sobol_interval_tables(const uint32_t* sobol_matrices_32, int m, uint64_t vdc_out[52], uint64_t vdc_inv_out[52]) {
    if (!sobol_matrices_32 || !vdc_out || !vdc_inv_out || m < 0 || m > 26) return -2;
    return b2host::sobol_interval_tables(sobol_matrices_32, m, vdc_out, vdc_inv_out) ? 0 : -2;
}
