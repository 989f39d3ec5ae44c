// Textured "Kd" of matte / plastic materials: kd.evaluate(&si.hit, &si.uv, &si.der) once per intersection
// (materials/src/matte.rs:63, plastic.rs:81) for the spectrum textures of this path - constant and the 2-D checkerboard
// (textures/src/checkerboard_2d.rs) over UVMapping2D (core/src/texture/mapping/uv_2d.rs).
//
// The checkerboard's default "aamode" closedform box-filters the pattern over the footprint (dudx, dvdx, dudy, dvdy) that
// SurfaceInteraction::compute_differentials (core/src/interaction/surface_interaction.rs:203-277) derives from the ray's
// differentials.  Only camera rays carry differentials on this path (PerspectiveCamera::generate_ray_differential,
// cameras/src/perspective_camera.rs:144-204, scaled by 1 / sqrt(spp) in render_tile, sampler_integrator.rs:357-358): the ray
// generation kernels store them per path (Wave::cam_diff, 3 float4) when the scene has such a texture, and the first vertex
// of a path reads them back; every later vertex point-samples (zero derivatives), as the reference's spawned rays do.
#pragma once
#include "shade.cuh"

namespace b2 {

struct DSpecTex {
    int type;            // B200PT_STEX_*
    float su, sv, du, dv;
    float tex1[3], tex2[3];
    int closedform;
};

// generate_ray_differential's rx / ry rays in world space, scaled (Ray::scale_differentials, ray.rs:90-99) around the
// already transformed main ray (o_w is the NUDGED origin transform_ray returns; the differential origins are not nudged,
// transform.rs:464-472).  out: (rx_o, rx_d.x) (rx_d.yz, ry_o.xy) (ry_o.z, ry_d).
B2_D void camera_differentials(const DCamera& c, P2 p_film, P2 p_lens, V3 o_w, V3 d_w, float scale, float4* out) {
    V3 rx_o, ry_o, rx_d, ry_d;
    if (c.type == B200PT_CAMERA_ENVIRONMENT) {
        // Camera::generate_ray_differential (core/src/camera.rs:29-78): finite differences over 0.05 pixel of the WORLD-space
        // rays (the weight is always 1, so the first eps is taken); the time sample does not move the ray
        const float eps = 0.05f;
        const Ray32 rx = camera_ray(c, mk2(p_film.x + eps, p_film.y), 0.0f, p_lens);
        const Ray32 ry = camera_ray(c, mk2(p_film.x, p_film.y + eps), 0.0f, p_lens);
        rx_o = o_w + (mk(rx.ox, rx.oy, rx.oz) - o_w) / eps;
        rx_d = d_w + (mk(rx.dx, rx.dy, rx.dz) - d_w) / eps;
        ry_o = o_w + (mk(ry.ox, ry.oy, ry.oz) - o_w) / eps;
        ry_d = d_w + (mk(ry.dx, ry.dy, ry.dz) - d_w) / eps;
    } else {
        const V3 p_camera = xf_point(c.r2c, mk(p_film.x, p_film.y, 0.0f));
        const float* m = c.c2w;
        auto vec = [&](V3 v) { return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); };
        if (c.type == B200PT_CAMERA_ORTHOGRAPHIC) {  // orthographic_camera.rs:120-146; dx_camera = raster_to_camera.transform_vector((1, 0, 0)), :44-49
            const float* r = c.r2c;
            const V3 dx_camera = mk(r[0] * 1.0f + r[1] * 0.0f + r[2] * 0.0f, r[4] * 1.0f + r[5] * 0.0f + r[6] * 0.0f, r[8] * 1.0f + r[9] * 0.0f + r[10] * 0.0f);
            const V3 dy_camera = mk(r[0] * 0.0f + r[1] * 1.0f + r[2] * 0.0f, r[4] * 0.0f + r[5] * 1.0f + r[6] * 0.0f, r[8] * 0.0f + r[9] * 1.0f + r[10] * 0.0f);
            if (c.lens_radius > 0.0f) {
                const P2 cd = concentric_sample_disk(p_lens);
                const P2 pl = mk2(c.lens_radius * cd.x, c.lens_radius * cd.y);
                // the main ray after the lens (camera space): its direction's z sets ft
                const float ft0 = c.focal_distance / 1.0f;
                const V3 p_focus = p_camera + mk(0.0f, 0.0f, 1.0f) * ft0;
                const V3 d_lens = normalize(p_focus - mk(pl.x, pl.y, 0.0f));
                const float ft = c.focal_distance / d_lens.z;
                rx_o = mk(pl.x, pl.y, 0.0f);
                rx_d = normalize(p_camera + dx_camera + (mk(0.0f, 0.0f, 1.0f) * ft) - rx_o);
                ry_o = rx_o;
                ry_d = normalize(p_camera + dy_camera + (mk(0.0f, 0.0f, 1.0f) * ft) - ry_o);
            } else {
                rx_o = p_camera + dx_camera; ry_o = p_camera + dy_camera;
                rx_d = mk(0.0f, 0.0f, 1.0f); ry_d = rx_d;
            }
        } else {
            const V3 c00 = xf_point(c.r2c, mk(0.0f, 0.0f, 0.0f));
            const V3 dx_camera = xf_point(c.r2c, mk(1.0f, 0.0f, 0.0f)) - c00;  // perspective_camera.rs:71-74
            const V3 dy_camera = xf_point(c.r2c, mk(0.0f, 1.0f, 0.0f)) - c00;
            rx_o = mk(0.0f, 0.0f, 0.0f); ry_o = rx_o;
            if (c.lens_radius > 0.0f) {  // :175-193
                const P2 cd = concentric_sample_disk(p_lens);
                const P2 pl = mk2(c.lens_radius * cd.x, c.lens_radius * cd.y);
                const V3 dx = normalize(p_camera + dx_camera);
                const float ftx = c.focal_distance / dx.z;
                const V3 p_focus_x = mk(0.0f, 0.0f, 0.0f) + (dx * ftx);
                rx_o = mk(pl.x, pl.y, 0.0f);
                rx_d = normalize(p_focus_x - rx_o);
                const V3 dy = normalize(p_camera + dy_camera);
                const float fty = c.focal_distance / dy.z;
                const V3 p_focus_y = mk(0.0f, 0.0f, 0.0f) + (dy * fty);
                ry_o = rx_o;
                ry_d = normalize(p_focus_y - ry_o);
            } else {  // :194-199
                rx_d = normalize(p_camera + dx_camera);
                ry_d = normalize(p_camera + dy_camera);
            }
        }
        rx_o = xf_point(m, rx_o); ry_o = xf_point(m, ry_o); rx_d = vec(rx_d); ry_d = vec(ry_d);  // Transform::transform_point / transform_vector
    }
    rx_o = o_w + (rx_o - o_w) * scale;
    ry_o = o_w + (ry_o - o_w) * scale;
    rx_d = d_w + (rx_d - d_w) * scale;
    ry_d = d_w + (ry_d - d_w) * scale;
    out[0] = make_float4(rx_o.x, rx_o.y, rx_o.z, rx_d.x);
    out[1] = make_float4(rx_d.y, rx_d.z, ry_o.x, ry_o.y);
    out[2] = make_float4(ry_o.z, ry_d.x, ry_d.y, ry_d.z);
}

B2_D float clamp0inf(float v) { return v < 0.0f ? 0.0f : v; }  // clamp(v, 0, INFINITY), pbrt/common.rs
B2_D float vcomp(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
// matrix4x4.rs:305-318
B2_D bool solve_2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x0, float* x1) {
    const float det = a00 * a11 - a01 * a10;
    if (pabs(det) < 1e-10f) return false;
    *x0 = (a11 * b0 - a01 * b1) / det;
    *x1 = (a00 * b1 - a10 * b0) / det;
    return !(isnan(*x0) || isnan(*x1));
}
B2_D float bump_int(float x) {  // checkerboard_2d.rs:101-103
    const float h = floorf(x / 2.0f);
    return h + 2.0f * pmax((x / 2.0f) - h - 0.5f, 0.0f);
}
B2_D int f32_as_i32(float f) {  // Rust `as i32`: saturating, NaN -> 0
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 0x7fffffff;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}

// si.der after SurfaceInteraction::compute_differentials: the screen-space derivatives of uv and of the hit point.
struct HitDerivs {
    float dudx, dvdx, dudy, dvdy;
    V3 dpdx, dpdy;
};
// compute_differentials (surface_interaction.rs:203-277) for a triangle hit.  (p0, p1, p2, duv) = the primitive's vertices and
// uv differences in the space the hit was found in; i2w = the instance's primitive_to_world 4x4 (null: top-level triangle or
// identity) for der.dpdu / der.dpdv (transform.rs:577-578); p, n = Hit::p / Hit::n in world space; diff = the ray's
// differentials (3 float4, camera_differentials' layout).  Out of line: one call site per kernel.
__device__ __noinline__ void hit_differentials(V3 p0, V3 p1, V3 p2, float4 duv, const float* i2w, V3 p, V3 n, const float4* diff, HitDerivs* out) {
    HitDerivs D;
    D.dudx = D.dvdx = D.dudy = D.dvdy = 0.0f;
    D.dpdx = D.dpdy = mk(0.0f, 0.0f, 0.0f);
    // der.dpdu / der.dpdv: the geometric partials of Triangle::intersect (triangle.rs:548-570)
    const V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float determinant = duv.x * duv.w - duv.y * duv.z;
    const bool degenerate_uv = pabs(determinant) < 1e-8f;
    V3 dpdu = mk(0.0f, 0.0f, 0.0f), dpdv = dpdu;
    if (!degenerate_uv) {
        const float invdet = 1.0f / determinant;
        dpdu = (duv.w * dp02 - duv.y * dp12) * invdet;
        dpdv = (-duv.z * dp02 + duv.x * dp12) * invdet;
    }
    if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) coordinate_system(normalize(cross(p2 - p0, p1 - p0)), &dpdu, &dpdv);
    if (i2w) {
        const float* m = i2w;
        dpdu = mk(m[0] * dpdu.x + m[1] * dpdu.y + m[2] * dpdu.z, m[4] * dpdu.x + m[5] * dpdu.y + m[6] * dpdu.z, m[8] * dpdu.x + m[9] * dpdu.y + m[10] * dpdu.z);
        dpdv = mk(m[0] * dpdv.x + m[1] * dpdv.y + m[2] * dpdv.z, m[4] * dpdv.x + m[5] * dpdv.y + m[6] * dpdv.z, m[8] * dpdv.x + m[9] * dpdv.y + m[10] * dpdv.z);
    }
    const float4 d0 = diff[0], d1 = diff[1], d2 = diff[2];
    const V3 rx_o = mk(d0.x, d0.y, d0.z), rx_d = mk(d0.w, d1.x, d1.y), ry_o = mk(d1.z, d1.w, d2.x), ry_d = mk(d2.y, d2.z, d2.w);
    const float d = dot(n, p);
    const float tx = -(dot(n, rx_o) - d) / dot(n, rx_d);
    const float ty = -(dot(n, ry_o) - d) / dot(n, ry_d);
    if (!(isinf(tx) || isnan(tx)) && !(isinf(ty) || isnan(ty))) {
        const V3 px = rx_o + rx_d * tx, py = ry_o + ry_d * ty;
        D.dpdx = px - p; D.dpdy = py - p;
        int a0, a1;
        if (pabs(n.x) > pabs(n.y) && pabs(n.x) > pabs(n.z)) { a0 = 1; a1 = 2; }
        else if (pabs(n.y) > pabs(n.z)) { a0 = 0; a1 = 2; }
        else { a0 = 0; a1 = 1; }
        const float a00 = vcomp(dpdu, a0), a01 = vcomp(dpdv, a0), a10 = vcomp(dpdu, a1), a11 = vcomp(dpdv, a1);
        const float bx0 = vcomp(px, a0) - vcomp(p, a0), bx1 = vcomp(px, a1) - vcomp(p, a1);
        const float by0 = vcomp(py, a0) - vcomp(p, a0), by1 = vcomp(py, a1) - vcomp(p, a1);
        if (!solve_2x2(a00, a01, a10, a11, bx0, bx1, &D.dudx, &D.dvdx)) { D.dudx = 0.0f; D.dvdx = 0.0f; }
        if (!solve_2x2(a00, a01, a10, a11, by0, by1, &D.dudy, &D.dvdy)) { D.dudy = 0.0f; D.dvdy = 0.0f; }
    }
    *out = D;
}

// Texture<Spectrum>::evaluate at (u, v) with the uv derivatives of D (zero = point sampling), before the material's clamp.
B2_D RGB spectrum_texture_eval(const DSpecTex& T, float u, float v, float dudx, float dvdx, float dudy, float dvdy) {
    if (T.type == 0) return rgb(T.tex1[0], T.tex1[1], T.tex1[2]);
    // uv_2d.rs:44-50
    const float dsdx = T.su * dudx, dtdx = T.sv * dvdx, dsdy = T.su * dudy, dtdy = T.sv * dvdy;
    const float s = T.su * u + T.du, t = T.sv * v + T.dv;
    const int sum = (int)((unsigned)f32_as_i32(floorf(s)) + (unsigned)f32_as_i32(floorf(t)));
    const bool first = sum % 2 == 0;
    const RGB point = first ? rgb(T.tex1[0], T.tex1[1], T.tex1[2]) : rgb(T.tex2[0], T.tex2[1], T.tex2[2]);
    if (!T.closedform) return point;
    // checkerboard_2d.rs:72-97
    const float ds = pmax(pabs(dsdx), pabs(dsdy)), dt = pmax(pabs(dtdx), pabs(dtdy));
    const float s0 = s - ds, s1 = s + ds, t0 = t - dt, t1 = t + dt;
    if (floorf(s0) == floorf(s1) && floorf(t0) == floorf(t1)) return point;
    const float sint = (bump_int(s1) - bump_int(s0)) / (2.0f * ds);
    const float tint = (bump_int(t1) - bump_int(t0)) / (2.0f * dt);
    const float area2 = (ds > 1.0f || dt > 1.0f) ? 0.5f : sint + tint - 2.0f * sint * tint;
    return rgb(T.tex1[0] * (1.0f - area2) + T.tex2[0] * area2, T.tex1[1] * (1.0f - area2) + T.tex2[1] * area2, T.tex1[2] * (1.0f - area2) + T.tex2[2] * area2);
}

// Kd of one intersection of the path integrator (uv6 = the primitive's three uvs; the rest as for hit_differentials; diff =
// null for a ray without differentials).  Out of line: one call site per shade kernel.
__device__ __noinline__ RGB kd_texture_eval(const DSpecTex* tp, const float* uv6, float b0, float b1, float b2, V3 p0, V3 p1, V3 p2, float4 duv, const float* i2w, V3 p,
                                            V3 n, const float4* diff) {
    const DSpecTex T = *tp;
    // triangle.rs:573-575: uv = b0 uv0 + b1 uv1 + b2 uv2
    const float u = b0 * uv6[0] + b1 * uv6[2] + b2 * uv6[4];
    const float v = b0 * uv6[1] + b1 * uv6[3] + b2 * uv6[5];
    HitDerivs D;
    D.dudx = D.dvdx = D.dudy = D.dvdy = 0.0f;
    if (T.type != 0 && T.closedform && diff) hit_differentials(p0, p1, p2, duv, i2w, p, n, diff, &D);
    return spectrum_texture_eval(T, u, v, D.dudx, D.dvdx, D.dudy, D.dvdy);
}

// The differentials of the child ray specular_reflect / specular_transmit spawn (sampler_integrator.rs:108-125, 164-227) from
// the parent's (diff), the hit's derivatives D, shading normal ns and its derivatives dndu / dndv, for the sampled direction wi.
// eta_bsdf = bsdf.eta (1 for every material of this path: BSDF::new(.., None)).  Layout of out: camera_differentials'.
B2_D void specular_child_differentials(const float4* diff, const HitDerivs& D, V3 p, V3 wo, V3 ns, V3 dndu, V3 dndv, V3 wi, bool transmit, float eta_bsdf, float4* out) {
    const float4 d0 = diff[0], d1 = diff[1], d2 = diff[2];
    const V3 prx_d = mk(d0.w, d1.x, d1.y), pry_d = mk(d2.y, d2.z, d2.w);
    const V3 rx_o = p + D.dpdx, ry_o = p + D.dpdy;
    V3 dndx = dndu * D.dudx + dndv * D.dvdx;
    V3 dndy = dndu * D.dudy + dndv * D.dvdy;
    V3 rx_d, ry_d;
    if (!transmit) {
        const V3 dwodx = -prx_d - wo, dwody = -pry_d - wo;
        const float ddndx = dot(dwodx, ns) + dot(wo, dndx);
        const float ddndy = dot(dwody, ns) + dot(wo, dndy);
        rx_d = wi - dwodx + 2.0f * (dot(wo, ns) * dndx + ddndx * ns);
        ry_d = wi - dwody + 2.0f * (dot(wo, ns) * dndy + ddndy * ns);
    } else {
        float eta = 1.0f / eta_bsdf;
        if (dot(wo, ns) < 0.0f) {
            eta = 1.0f / eta;
            ns = -ns; dndx = -dndx; dndy = -dndy;
        }
        const V3 dwodx = -prx_d - wo, dwody = -pry_d - wo;
        const float ddndx = dot(dwodx, ns) + dot(wo, dndx);
        const float ddndy = dot(dwody, ns) + dot(wo, dndy);
        const float mu = eta * dot(wo, ns) - abs_dot(wi, ns);
        const float dmudx = (eta - (eta * eta * dot(wo, ns)) / abs_dot(wi, ns)) * ddndx;
        const float dmudy = (eta - (eta * eta * dot(wo, ns)) / abs_dot(wi, ns)) * ddndy;
        rx_d = wi - eta * dwodx + (mu * dndx + dmudx * ns);
        ry_d = wi - eta * dwody + (mu * dndy + dmudy * ns);
    }
    out[0] = make_float4(rx_o.x, rx_o.y, rx_o.z, rx_d.x);
    out[1] = make_float4(rx_d.y, rx_d.z, ry_o.x, ry_o.y);
    out[2] = make_float4(ry_o.z, ry_d.x, ry_d.y, ry_d.z);
}

}  // namespace b2
