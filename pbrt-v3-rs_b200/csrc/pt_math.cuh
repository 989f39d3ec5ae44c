// Device-side scalar/vector helpers for the B200 path tracer.
//
// Bit-parity rules (SURVEY.md §7 "hard parts"): the reference is Rust f32 —
// never FMA-contracted, never re-associated, IEEE divide and sqrt.  This
// translation unit is therefore compiled with -fmad=false and the default
// -prec-div=true -prec-sqrt=true -ftz=false, and comparisons keep the
// reference's direction (a < b ? a : b), never fminf/fmaxf, so NaNs fall
// through exactly as in core/src/pbrt/common.rs:83-108.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__

constexpr float kMachineEps = 5.9604644775390625e-08f;  // f32::EPSILON * 0.5, common.rs:46
constexpr float kShadowEps = 0.0001f;
constexpr float kOneMinusEps = 0x1.fffffep-1f;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 1.0f / kPi;
constexpr float kPiOver2 = kPi * 0.5f;
constexpr float kPiOver4 = kPi * 0.25f;
constexpr float kTwoPi = kPi * 2.0f;
constexpr float kInvTwoPi = 1.0f / kTwoPi;
constexpr float kFourPi = kPi * 4.0f;

// core/src/pbrt/common.rs:130-133, evaluated in f32 at compile time.
constexpr float gamma_c(int n) { return ((float)n * kMachineEps) / (1.0f - (float)n * kMachineEps); }
constexpr float kGamma2 = gamma_c(2), kGamma3 = gamma_c(3), kGamma5 = gamma_c(5), kGamma6 = gamma_c(6), kGamma7 = gamma_c(7);
constexpr float kSlabInflate = 1.0f + 2.0f * kGamma3;  // bounds3.rs:304-306

B2_HD float pmin(float a, float b) { return a < b ? a : b; }
B2_HD float pmax(float a, float b) { return a > b ? a : b; }
B2_HD float pabs(float a) { return a < 0.0f ? -a : a; }
B2_HD float pclamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

struct V3 {
    float x, y, z;
};
B2_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
B2_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
B2_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
B2_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
B2_HD V3 operator*(float f, V3 v) { return mk(f * v.x, f * v.y, f * v.z); }
B2_HD V3 operator*(V3 v, float f) { return mk(f * v.x, f * v.y, f * v.z); }
// core/src/geometry/vector3.rs:408-417: division multiplies by the reciprocal
B2_HD V3 operator/(V3 v, float f) { float inv = 1.0f / f; return mk(inv * v.x, inv * v.y, inv * v.z); }
B2_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2_HD float abs_dot(V3 a, V3 b) { return pabs(dot(a, b)); }
B2_HD V3 cross(V3 a, V3 b) { return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)); }
B2_HD float length_squared(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
B2_HD float length(V3 v) { return sqrtf(length_squared(v)); }
B2_HD V3 normalize(V3 v) { return v / length(v); }
B2_HD V3 vabs(V3 v) { return mk(pabs(v.x), pabs(v.y), pabs(v.z)); }
B2_HD float max_component(V3 v) { return v.x > v.y ? (v.x > v.z ? v.x : v.z) : (v.y > v.z ? v.y : v.z); }  // vector3.rs:114-128
B2_HD int max_dimension(V3 v) { return v.x > v.y ? (v.x > v.z ? 0 : 2) : (v.y > v.z ? 1 : 2); }          // vector3.rs:133-148
B2_HD V3 face_forward(V3 n, V3 v) { return dot(n, v) < 0.0f ? -n : n; }
B2_HD float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
B2_HD float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// core/src/geometry/coordinate_system.rs:12-20
B2_HD void coordinate_system(V3 v1, V3* v2, V3* v3) {
    if (pabs(v1.x) > pabs(v1.y)) *v2 = mk(-v1.z, 0.0f, v1.x) / sqrtf(v1.x * v1.x + v1.z * v1.z);
    else *v2 = mk(0.0f, v1.z, -v1.y) / sqrtf(v1.y * v1.y + v1.z * v1.z);
    *v3 = cross(v1, *v2);
}

// core/src/pbrt/common.rs:205-243
B2_D float next_float_up(float v) {
    if (isinf(v) && v > 0.0f) return v;
    float nv = (v == -0.0f) ? 0.0f : v;
    uint32_t ui = __float_as_uint(nv);
    if (nv >= 0.0f) ui += 1; else ui -= 1;
    return __uint_as_float(ui);
}
B2_D float next_float_down(float v) {
    if (isinf(v) && v < 0.0f) return v;
    float nv = (v == 0.0f) ? -0.0f : v;
    uint32_t ui = __float_as_uint(nv);
    if (nv > 0.0f) ui -= 1; else ui += 1;
    return __uint_as_float(ui);
}
// core/src/geometry/ray.rs:107-127
B2_D V3 offset_ray_origin(V3 p, V3 p_error, V3 n, V3 w) {
    float d = dot(vabs(n), p_error);
    V3 offset = d * n;
    if (dot(w, n) < 0.0f) offset = -offset;
    V3 po = p + offset;
    if (offset.x > 0.0f) po.x = next_float_up(po.x); else if (offset.x < 0.0f) po.x = next_float_down(po.x);
    if (offset.y > 0.0f) po.y = next_float_up(po.y); else if (offset.y < 0.0f) po.y = next_float_down(po.y);
    if (offset.z > 0.0f) po.z = next_float_up(po.z); else if (offset.z < 0.0f) po.z = next_float_down(po.z);
    return po;
}

// 16-byte vector loads through the read-only path (ld.global.nc.v4.f32).
B2_D float4 ldg4(const float4* p) { return __ldg(p); }
// 32-byte vector load (sm_100: LDG.E.256).  For scattered per-lane records the L1 tag stage is the limiter
// (one 128-B line per lane per instruction), so fetching a 64-byte node as 2 x 256-bit instead of 4 x 128-bit
// halves the L1 wavefronts per traversal step.  p must be 32-byte aligned.
B2_D void ldg8(const float4* p, float4* a, float4* b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
                 : "l"(p));
}

}  // namespace b2
