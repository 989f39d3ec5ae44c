// Host-side sampler tables (internal header).
#pragma once
#include <cstdint>
#include <vector>

namespace b2host {

struct HaltonTables {
    std::vector<int> primes, prime_sums;
    std::vector<uint16_t> perms;
    // exact u32 division by each prime without a divide (Granlund-Montgomery): q = (t + ((n - t) >> sh1)) >> sh2, t = mulhi(m, n)
    std::vector<uint32_t> div_m, div_sh;  // div_sh = sh1 | sh2 << 8
};
const HaltonTables& halton_tables();

struct HaltonParams {
    uint64_t base_scale[2], base_exp[2], stride;
    int64_t mult_inv[2];
};
HaltonParams halton_params(int res_x, int res_y);

// SobolSampler: VD_C_SOBOL_MATRICES[m - 1] / VD_C_SOBOL_MATRICES_INV[m - 1] derived from SOBOL_MATRICES_32 (m = log2 of the
// power-of-two resolution, 0..26).  Returns false if the matrices are not those of a (0, 2)-sequence.
bool sobol_interval_tables(const uint32_t* m32, int m, uint64_t vdc[52], uint64_t vdc_inv[52]);

}  // namespace b2host
