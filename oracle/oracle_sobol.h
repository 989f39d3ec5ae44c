// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.h header).  PARITY UNPINNED for this file: no render of the reference on this
// path uses the Sobol sampler; its tables are checked against the reference's literals (see oracle_math.h for what is pinned by execution).
//
// SobolSampler, restated from samplers/src/sobol.rs:25-196 and core/src/low_discrepency.rs:1770-1845
// (sobol_interval_to_index, sobol_sample_f32).
//
// The 1024 x 52 generator matrices (core/src/sobol_matrices.rs: SOBOL_MATRICES_32) are data the caller supplies
// (orc_set_sobol_matrices; pbrt-v3-rs_b200/data/sobol_matrices_32.bin, written by tools/extract_sobol_matrices.py).
// The van der Corput / inverse tables that sobol_interval_to_index reads (VD_C_SOBOL_MATRICES[m - 1],
// VD_C_SOBOL_MATRICES_INV[m - 1]) are DERIVED here from dimensions 0 and 1: index bit j moves the pixel by
// col(j) = (top m bits of M0[j]) << m | (top m bits of M1[j]); the table row for resolution 2^m holds col(2m + c)
// (how the sample number shifts the pixel) and the columns of the inverse of the 2m x 2m matrix [col(0) .. col(2m-1)]
// over GF(2) (which low index bits land in a given pixel).  tests/test_sobol.py compares the derivation with the
// reference's literal tables whenever /root/reference is mounted.
#pragma once
#include <vector>
#include "oracle_rng.h"

namespace orc {

static const int kSobolDims = 1024, kSobolMatrixSize = 52;

inline std::vector<uint32_t>& sobol_matrices_32() {
    static std::vector<uint32_t> m;
    return m;
}

struct SobolIntervalTables {
    uint64_t vdc[kSobolMatrixSize];      // VD_C_SOBOL_MATRICES[m - 1]
    uint64_t vdc_inv[kSobolMatrixSize];  // VD_C_SOBOL_MATRICES_INV[m - 1]
};
inline SobolIntervalTables sobol_interval_tables(int m) {
    SobolIntervalTables t;
    const uint32_t* M0 = sobol_matrices_32().data();
    const uint32_t* M1 = M0 + kSobolMatrixSize;
    auto col = [&](int j) -> uint64_t {
        if (j >= kSobolMatrixSize) return 0;
        return ((uint64_t)(M0[j] >> (32 - m)) << m) | (uint64_t)(M1[j] >> (32 - m));
    };
    const int m2 = 2 * m;
    for (int c = 0; c < kSobolMatrixSize; ++c) t.vdc[c] = (m2 + c < kSobolMatrixSize) ? col(m2 + c) : 0;
    // Gauss-Jordan over GF(2): rows = output bits, columns = index bits, augmented with the identity
    std::vector<uint64_t> a((size_t)m2), inv((size_t)m2);
    for (int r = 0; r < m2; ++r) {
        uint64_t row = 0;
        for (int j = 0; j < m2; ++j) if ((col(j) >> r) & 1) row |= 1ull << j;
        a[(size_t)r] = row;
        inv[(size_t)r] = 1ull << r;
    }
    for (int j = 0; j < m2; ++j) {
        int p = -1;
        for (int r = j; r < m2; ++r) if ((a[(size_t)r] >> j) & 1) { p = r; break; }
        std::swap(a[(size_t)j], a[(size_t)p]);
        std::swap(inv[(size_t)j], inv[(size_t)p]);
        for (int r = 0; r < m2; ++r)
            if (r != j && ((a[(size_t)r] >> j) & 1)) { a[(size_t)r] ^= a[(size_t)j]; inv[(size_t)r] ^= inv[(size_t)j]; }
    }
    // now index bit j = XOR over output bits c with (inv[j] >> c) & 1
    for (int c = 0; c < kSobolMatrixSize; ++c) {
        uint64_t v = 0;
        if (c < m2) for (int j = 0; j < m2; ++j) if ((inv[(size_t)j] >> c) & 1) v |= 1ull << j;
        t.vdc_inv[c] = v;
    }
    return t;
}

// low_discrepency.rs:1770-1808
inline uint64_t sobol_interval_to_index(const SobolIntervalTables& t, uint32_t m, uint64_t frame, int px, int py) {
    if (m == 0) return 0;
    const uint32_t m2 = m << 1;
    uint64_t index = frame << m2;
    uint64_t delta = 0;
    for (int c = 0; frame > 0; frame >>= 1, ++c)
        if (frame & 1) delta ^= t.vdc[c];
    uint64_t b = ((((uint64_t)(uint32_t)px) << m) | (uint64_t)(uint32_t)py) ^ delta;
    for (int c = 0; b > 0; b >>= 1, ++c)
        if (b & 1) index ^= t.vdc_inv[c];
    return index;
}
// low_discrepency.rs:1821-1845 (scramble = 0)
inline Float sobol_sample_f32(uint64_t a, int dimension) {
    uint32_t v = 0;
    const uint32_t* m = sobol_matrices_32().data() + (size_t)dimension * kSobolMatrixSize;
    for (; a != 0; a >>= 1, ++m)
        if (a & 1) v ^= *m;
    Float s = (Float)v * 0x1.0p-32f;
    return pmin(s, kOneMinusEpsilon);
}

// samplers/src/sobol.rs:25-196 (no sample arrays, as for Halton)
struct SobolSampler : Sampler {
    int sb_min[2];
    int resolution, log2_resolution;
    SobolIntervalTables tables;
    int cur_px = 0, cur_py = 0, cur_sample = 0, dimension = 0;
    uint64_t interval_sample_index = 0;

    SobolSampler(int spp_, const int sample_bounds[4]) {
        int p2 = 1;
        while (p2 < spp_) p2 <<= 1;  // sobol.rs:28-37: rounded up to a power of two
        spp = p2;
        sb_min[0] = sample_bounds[0]; sb_min[1] = sample_bounds[1];
        int ext = pmax(sample_bounds[2] - sample_bounds[0], sample_bounds[3] - sample_bounds[1]);
        uint32_t r = 1;
        while (r < (uint32_t)ext) r <<= 1;  // next_power_of_two
        resolution = (int)r;
        log2_resolution = 0;
        while ((1 << log2_resolution) < resolution) ++log2_resolution;
        if (log2_resolution > 0) tables = sobol_interval_tables(log2_resolution);
    }
    uint64_t get_index_for_sample(uint64_t sample_num) const {
        return sobol_interval_to_index(tables, (uint32_t)log2_resolution, sample_num, cur_px - sb_min[0], cur_py - sb_min[1]);
    }
    Float sample_dimension(uint64_t index, int dim) const {  // sobol.rs:64-80
        Float s = sobol_sample_f32(index, dim);
        if (dim == 0 || dim == 1) {
            s = s * (Float)resolution + (Float)sb_min[dim];
            s = pclamp(s - (Float)(dim == 0 ? cur_px : cur_py), 0.0f, kOneMinusEpsilon);
        }
        return s;
    }
    void start_pixel(int px, int py) override {
        cur_px = px; cur_py = py; cur_sample = 0; dimension = 0;
        interval_sample_index = get_index_for_sample(0);
    }
    Float get_1d() override { Float p = sample_dimension(interval_sample_index, dimension); dimension += 1; return p; }
    P2 get_2d() override {
        P2 p(sample_dimension(interval_sample_index, dimension), sample_dimension(interval_sample_index, dimension + 1));
        dimension += 2;
        return p;
    }
    bool start_next_sample() override {
        dimension = 0;
        interval_sample_index = get_index_for_sample((uint64_t)cur_sample + 1);
        cur_sample += 1;
        return cur_sample < spp;
    }
    void set_sample_number(int s) {
        dimension = 0;
        cur_sample = s;
        interval_sample_index = get_index_for_sample((uint64_t)s);
    }
};

}  // namespace orc
