// Host check of pbrt-v3-rs_b200/csrc/libm_exact.cuh against the platform libm (glibc sinf/cosf):
// prints the number of inputs whose bits differ.  Built and run by tests/test_libm_exact.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../pbrt-v3-rs_b200/csrc/libm_exact.cuh"
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 100000000L;
    uint64_t s = 88172645463325252ULL;
    long bad_s = 0, bad_c = 0;
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        float u = (float)((s >> 40) * (1.0 / 16777216.0));
        float x;
        switch (i & 3) {
            case 0: x = u * 6.2831853f; break;              // [0, 2pi)
            case 1: x = (u - 0.5f) * 6.2831853f; break;     // [-pi, pi)
            case 2: x = (u - 0.5f) * 238.0f; break;         // (-119, 119)
            default: x = u * u * u * 1e-2f; break;          // tiny arguments
        }
        float a = sinf(x), b = lmx::sinf_glibc(x);
        if (lmx::fbits(a) != lmx::fbits(b)) ++bad_s;
        a = cosf(x); b = lmx::cosf_glibc(x);
        if (lmx::fbits(a) != lmx::fbits(b)) ++bad_c;
    }
    printf("%ld %ld %ld\n", n, bad_s, bad_c);
    return (bad_s || bad_c) ? 1 : 0;
}
