// Wavefront state and the device functions every shading stage shares: the scene as the kernels see it, the wave's
// structure-of-arrays buffers, sampler access, hit reconstruction (Triangle::intersect's SurfaceInteraction incl. the
// TransformedPrimitive transform), Light::sample_li, estimate_direct up to the rays it traces, and the voxel lookup of
// the spatial light distribution.  Internal header of b200pt_scene.cu (kernels: b200pt_scene.cu, shade_tree.cuh).
#pragma once
#include "common.cuh"
#include "instancing.cuh"
#include "shade.cuh"
#include "spectrum_tex.cuh"

namespace b2 {

struct DeviceScene {
    DeviceAccel accel;
    const float4* prim_verts;  // 3 float4 per ORIGINAL primitive: (p0, bits material), (p1, bits light or -1), (p2, bits flags)
    const float4* prim_duv;    // optional (meshes with uvs): uv0 - uv2, uv1 - uv2 per ORIGINAL primitive
    const float* prim_n;       // optional (meshes with N): 9 floats per ORIGINAL primitive
    const float* prim_s;       // optional (meshes with S): 9 floats per ORIGINAL primitive
    const DMaterial* materials;
    // textured "Kd" (null / unused when the scene has none): per material the spectrum-texture index or -1, the textures,
    // and 6 floats of uv per ORIGINAL primitive (defaults (0,0) (1,0) (1,1) filled in for meshes without uvs)
    const int* mat_kd_tex;
    const DSpecTex* spec_tex;
    const float* prim_uv6;
    float cam_diff_scale;  // 1 / sqrt(samples per pixel): Ray::scale_differentials in render_tile (sampler_integrator.rs:358)
    const DLight* lights;
    const int* infinite_lights;
    const DInfDistr* inf_distr;
    const DInstance* instances;  // null unless the scene has TransformedPrimitives
    const float* light_func;
    const float* light_cdf;
    float light_func_int;
    int n_lights, n_infinite;
    // SpatialLightDistribution (light_distrib/spatial.rs): voxel rows handed out from a pool on first touch (the
    // reference's hash table is filled lazily too, spatial.rs:170-250).  A row holds func[n_lights], cdf[n_lights + 1],
    // func_int; vox_state[v] = 0 untouched, 1 queued, 2 ready; vox_row[v] = the voxel's row once queued.
    int spatial;
    int n_voxels[3];
    float wb[6];          // scene.world_bound
    float* vox_table;
    int* vox_state;
    int* vox_row;
    int* vox_pool_next;   // rows handed out so far
    int vox_pool_cap;
    int* vox_work;        // voxels queued by the current shade launch
    DHalton halton;
    DZeroTwo zt;
    DSobol sobol;
    int sampler_type;  // B200PT_SAMPLER_*
    DCamera camera;
    int max_depth;
    float rr_threshold;
    float world_radius;
    int pb[4];  // integrator pixel bounds
    int sb[4];  // film sample bounds
};

// Wave buffers.  "slot" arrays are indexed by queue position, "path" arrays by path id within the wave.
struct Wave {
    // ray queues (double buffered) and their path ids
    float4* ray[2];     // 2 float4 per slot
    int* qpid[2];
    float4* hit;        // per slot
    float* hit_b2;
    int* hit_inst;      // instance of the hit (two-level scenes), -1 = top-level triangle
    // shadow / MIS queues
    float4* sh_ray;     // 2 float4 per slot
    uint8_t* sh_occ;
    float4* mis_ray;
    float4* mis_hit;
    float* mis_b2;      // third barycentric of the MIS hits (needed when the light's mesh has vertex normals)
    // per path
    float4* L;          // rgb, -
    float4* beta;       // rgb, eta_scale
    unsigned long long* hidx;
    int* meta;          // dim (16) | bounces (8) | specular flag (bit 24) | the ray is no camera ray any more: no differentials (bit 25)
    float4* cam_diff;   // 3 float4 per path: the camera ray's scaled differentials (null unless a closedform checkerboard needs them)
    // pending direct lighting, per path
    float4* pend_a;     // ld_light rgb, pick pdf
    float4* pend_b;     // mis f rgb, mis weight
    float4* pend_c;     // beta rgb at the vertex, mis scattering pdf
    int4* pend_d;       // light index, shadow slot, mis slot, -
    int* pend_q;        // path ids with a pending record
    int* counters;      // the current control block: [0] next rays, [1] shadow rays, [2] mis rays, [3] pending records, [4] sampler-dimension overflow,
                        // [5] [6] real shadow / MIS rays (tree integrators), [8..15] bin counts, [16..23] bin cursors, [24] voxels queued,
                        // [25] slots parked (spatial light sampling), [26] spatial row pool exhausted, [32..37] traversal work counters
    // sort-by-material: key per slot (0 = miss / dead, 1 + material type otherwise) and the slots grouped by key
    uint8_t* key;
    int* sorted;
    // Whitted integrator only: per-path stack of postponed specular-transmission children (3 float4 per entry,
    // max_depth entries per path) and the per-shadow-ray contribution of the light loop (rgb, valid)
    float4* wstack;
    int wstack_n;       // float4 per stack entry: 3, or 6 when the rays carry differentials (cam_diff)
    float4* sh_c;
    // DirectLighting integrator only: per-slot MIS record (f rgb, weight) and (scattering pdf, light index)
    float4* dp_b;
    float2* dp_c;
    // spatial light sampling: slots whose voxel was not ready in the first shade launch of a bounce
    int* deferred;
};
static const int kBins = 7;     // miss, matte, plastic, glass, metal, mirror, null material
static const int kBinNull = 6;
// Control block: 64 ints per bounce iteration.  Block 0 opens the wave ([0] = its paths); block i + 1 receives the counts
// iteration i produces, and its [0] is the queue size of iteration i + 1.  [32..37] are the three 8-byte work counters of
// the persistent traversal launches that consume the block's queues (closest, shadow, MIS).
static const int kCtl = 64;
static const int kSegIters = 16;  // iterations per control-block segment (deeper paths continue in a new segment)

B2_D int meta_pack(int dim, int bounces, int spec) { return (dim & 0xffff) | ((bounces & 0xff) << 16) | ((spec & 0xff) << 24); }

// Sampler::get_1d / get_2d for the path's current sample.  `key` is the Halton sample index, or for the (0,2)
// sampler (owned pixel index << 16 | sample number); `dim` is the Halton dimension counter, or the 1-D slot counter
// in its low byte and the 2-D slot counter in the next byte (core/src/sampler/pixel_sampler.rs:88-110).
B2_D float smp_1d(const DeviceScene& S, unsigned long long key, int& dim) {
    if (S.sampler_type == B200PT_SAMPLER_HALTON) { float v = halton_dim(S.halton, key, dim); dim += 1; return v; }
    if (S.sampler_type == B200PT_SAMPLER_SOBOL) { float v = sobol_sample_f32(S.sobol, key, dim); dim += 1; return v; }  // dimensions >= 2 only (k_raygen draws 0 / 1)
    int d1 = dim & 0xff;
    if (S.zt.rng && d1 >= S.zt.n1) return u32_to_unit(pcg_next(S.zt.rng[key >> 16]));  // RNG::uniform_float, rng.rs:98-103
    float v = zt_1d(S.zt, (long long)(key >> 16), d1, (int)(key & 0xffff));
    dim += 1;
    return v;
}
B2_D P2 smp_2d(const DeviceScene& S, unsigned long long key, int& dim) {
    if (S.sampler_type == B200PT_SAMPLER_HALTON) { P2 v = mk2(halton_dim(S.halton, key, dim), halton_dim(S.halton, key, dim + 1)); dim += 2; return v; }
    if (S.sampler_type == B200PT_SAMPLER_SOBOL) { P2 v = mk2(sobol_sample_f32(S.sobol, key, dim), sobol_sample_f32(S.sobol, key, dim + 1)); dim += 2; return v; }
    int d2 = (dim >> 8) & 0xff;
    if (S.zt.rng && d2 >= S.zt.n2) {
        DPcg32& r = S.zt.rng[key >> 16];
        const float a = u32_to_unit(pcg_next(r));
        const float b = u32_to_unit(pcg_next(r));
        return mk2(a, b);
    }
    P2 v = zt_2d(S.zt, (long long)(key >> 16), d2, (int)(key & 0xffff));
    dim += 0x100;
    return v;
}

B2_D void load_prim(const DeviceScene& S, uint32_t prim, V3* p0, V3* p1, V3* p2, int* mat, int* light, uint32_t* flags) {
    float4 a = ldg4(S.prim_verts + 3ll * prim), b = ldg4(S.prim_verts + 3ll * prim + 1), c = ldg4(S.prim_verts + 3ll * prim + 2);
    *p0 = mk(a.x, a.y, a.z); *p1 = mk(b.x, b.y, b.z); *p2 = mk(c.x, c.y, c.z);
    *mat = __float_as_int(a.w); *light = __float_as_int(b.w); *flags = __float_as_uint(c.w);
}

B2_D void store_ray(float4* q, int slot, V3 o, V3 d, float tmax, float time) {
    q[2 * slot] = make_float4(o.x, o.y, o.z, tmax);
    q[2 * slot + 1] = make_float4(d.x, d.y, d.z, time);
}

// The SurfaceInteraction of a closest hit: Triangle::intersect's geometry (triangle.rs:548-725), Hit::new's normalised
// wo (interaction/mod.rs:117-136) and, for a hit inside an object instance, transform_surface_interaction.
struct HitCtx {
    SurfHit sh;
    V3 wo;
    int mat, alight;
    uint32_t pflags;
};
B2_D void surface_at(const DeviceScene& S, const Wave& W, int slot, uint32_t prim, float4 hit, float hb2, V3 ray_d, bool found, HitCtx* out) {
    V3 p0, p1, p2;
    int mat = 0, alight = -1;
    uint32_t pflags = 0;
    SurfHit& sh = out->sh;
    // Hit::new normalises wo (interaction/mod.rs:117-136)
    V3 wo_raw = -ray_d;
    float l2 = length_squared(wo_raw);
    V3 hit_wo = (l2 == 0.0f) ? wo_raw : wo_raw / sqrtf(l2);
    if (found) {
        load_prim(S, prim, &p0, &p1, &p2, &mat, &alight, &pflags);
        const float4 duv = (S.prim_duv && (pflags & B200PT_PRIM_HAS_UV)) ? ldg4(S.prim_duv + prim) : default_duv();
        const float* vn = (S.prim_n && (pflags & B200PT_PRIM_HAS_NORMALS)) ? S.prim_n + 9ll * prim : nullptr;
        const float* vs = (S.prim_s && (pflags & B200PT_PRIM_HAS_TANGENTS)) ? S.prim_s + 9ll * prim : nullptr;
        sh = triangle_surface(p0, p1, p2, hit.z, hit.w, hb2, pflags, duv, vn, vs);
        const int inst = S.instances ? W.hit_inst[slot] : -1;
        if (inst >= 0) {
            // The hit was built in instance space from the instance-space ray, then
            // Transform::transform_surface_interaction(primitive_to_world) (transform.rs:566-590).
            const DInstance I = S.instances[inst];
            V3 wo_i = -mk(I.w2i[0] * ray_d.x + I.w2i[1] * ray_d.y + I.w2i[2] * ray_d.z, I.w2i[4] * ray_d.x + I.w2i[5] * ray_d.y + I.w2i[6] * ray_d.z,
                          I.w2i[8] * ray_d.x + I.w2i[9] * ray_d.y + I.w2i[10] * ray_d.z);
            float li2 = length_squared(wo_i);
            wo_i = (li2 == 0.0f) ? wo_i : wo_i / sqrtf(li2);
            if (I.identity) hit_wo = wo_i;
            else {
                const float* m = I.i2w;
                float x = sh.p.x, y = sh.p.y, z = sh.p.z;
                V3 pe = sh.p_error;
                V3 pw = mk((m[0] * x + m[1] * y) + (m[2] * z + m[3]), (m[4] * x + m[5] * y) + (m[6] * z + m[7]), (m[8] * x + m[9] * y) + (m[10] * z + m[11]));
                V3 ew = mk((kGamma3 + 1.0f) * (pabs(m[0]) * pe.x + pabs(m[1]) * pe.y + pabs(m[2]) * pe.z) + kGamma3 * (pabs(m[0] * x) + pabs(m[1] * y) + pabs(m[2] * z) + pabs(m[3])),
                           (kGamma3 + 1.0f) * (pabs(m[4]) * pe.x + pabs(m[5]) * pe.y + pabs(m[6]) * pe.z) + kGamma3 * (pabs(m[4] * x) + pabs(m[5] * y) + pabs(m[6] * z) + pabs(m[7])),
                           (kGamma3 + 1.0f) * (pabs(m[8]) * pe.x + pabs(m[9]) * pe.y + pabs(m[10]) * pe.z) + kGamma3 * (pabs(m[8] * x) + pabs(m[9] * y) + pabs(m[10] * z) + pabs(m[11])));
                const float wp = (m[12] * x + m[13] * y) + (m[14] * z + m[15]);  // transform.rs:338-368
                if (!(wp == 1.0f)) pw = pw / wp;
                sh.p = pw; sh.p_error = ew;
                hit_wo = normalize(mk(m[0] * wo_i.x + m[1] * wo_i.y + m[2] * wo_i.z, m[4] * wo_i.x + m[5] * wo_i.y + m[6] * wo_i.z, m[8] * wo_i.x + m[9] * wo_i.y + m[10] * wo_i.z));
                const float* mi = I.w2i;  // transform_normal: inverse transpose (transform.rs:439-446)
                V3 n = sh.n;
                sh.n = normalize(mk(mi[0] * n.x + mi[4] * n.y + mi[8] * n.z, mi[1] * n.x + mi[5] * n.y + mi[9] * n.z, mi[2] * n.x + mi[6] * n.y + mi[10] * n.z));
                V3 sn = sh.ns;  // si.shading.n = transform_normal(shading.n).normalize().face_forward(hit.n)
                sn = normalize(mk(mi[0] * sn.x + mi[4] * sn.y + mi[8] * sn.z, mi[1] * sn.x + mi[5] * sn.y + mi[9] * sn.z, mi[2] * sn.x + mi[6] * sn.y + mi[10] * sn.z));
                sh.ns = face_forward(sn, sh.n);
                V3 du = sh.dpdu;
                sh.dpdu = mk(m[0] * du.x + m[1] * du.y + m[2] * du.z, m[4] * du.x + m[5] * du.y + m[6] * du.z, m[8] * du.x + m[9] * du.y + m[10] * du.z);
            }
        }
    }
    out->wo = hit_wo;
    out->mat = mat;
    out->alight = alight;
    out->pflags = pflags;
}

// Material::compute_scattering_functions of a matte / plastic material whose "Kd" is spectrum texture `tex`
// (matte.rs:63-72, plastic.rs:81-84): kd = texture(si).clamp(); the diffuse lobe - always lobe 0 of `base`, which the host
// builds with a placeholder Kd - takes it, or is left out when it is black.  diff = the path's camera differentials or null.
B2_D void textured_material(const DeviceScene& S, const Wave& W, int slot, uint32_t prim, float4 hit, float hb2, const SurfHit& sh, int tex, const float4* diff,
                            const DMaterial& base, DMaterial* out) {
    V3 p0, p1, p2;
    int mat_unused, light_unused;
    uint32_t pflags;
    load_prim(S, prim, &p0, &p1, &p2, &mat_unused, &light_unused, &pflags);
    const float4 duv = (S.prim_duv && (pflags & B200PT_PRIM_HAS_UV)) ? ldg4(S.prim_duv + prim) : default_duv();
    const int inst = S.instances ? W.hit_inst[slot] : -1;
    const float* i2w = (inst >= 0 && !S.instances[inst].identity) ? S.instances[inst].i2w : nullptr;
    RGB kd = kd_texture_eval(S.spec_tex + tex, S.prim_uv6 + 6ll * prim, hit.z, hit.w, hb2, p0, p1, p2, duv, i2w, sh.p, sh.n, diff);
    kd = rgb(clamp0inf(kd.r), clamp0inf(kd.g), clamp0inf(kd.b));  // Spectrum::clamp_default
    *out = base;
    if (is_black(kd)) {  // `if !r.is_black()`: no diffuse lobe
        out->n_bxdf = base.n_bxdf - 1;
        out->bx[0] = base.bx[1];
    } else {
        out->bx[0].r[0] = kd.r; out->bx[0].r[1] = kd.g; out->bx[0].r[2] = kd.b;
    }
}

// Tree integrators with ray differentials: si.der (compute_differentials), the interpolated uv and shading.dndu / dndv
// (triangle.rs:679-713; transform_normal for a hit inside an instance, transform.rs:587-588) of the hit in `slot`.
B2_D void tree_hit_derivs(const DeviceScene& S, const Wave& W, int slot, uint32_t prim, float4 hit, float hb2, const SurfHit& sh, const float4* diff, HitDerivs* D,
                          float* u, float* v, V3* dndu, V3* dndv) {
    V3 p0, p1, p2;
    int mat_unused, light_unused;
    uint32_t pflags;
    load_prim(S, prim, &p0, &p1, &p2, &mat_unused, &light_unused, &pflags);
    const float4 duv = (S.prim_duv && (pflags & B200PT_PRIM_HAS_UV)) ? ldg4(S.prim_duv + prim) : default_duv();
    const int inst = S.instances ? W.hit_inst[slot] : -1;
    const bool xf = inst >= 0 && !S.instances[inst].identity;
    hit_differentials(p0, p1, p2, duv, xf ? S.instances[inst].i2w : nullptr, sh.p, sh.n, diff, D);
    const float* uv6 = S.prim_uv6 + 6ll * prim;
    *u = hit.z * uv6[0] + hit.w * uv6[2] + hb2 * uv6[4];
    *v = hit.z * uv6[1] + hit.w * uv6[3] + hb2 * uv6[5];
    V3 du = mk(0.0f, 0.0f, 0.0f), dv = du;
    if (S.prim_n && (pflags & B200PT_PRIM_HAS_NORMALS)) {
        const float* vn = S.prim_n + 9ll * prim;
        const V3 n0 = mk(vn[0], vn[1], vn[2]), n1 = mk(vn[3], vn[4], vn[5]), n2 = mk(vn[6], vn[7], vn[8]);
        const V3 dn1 = n0 - n2, dn2 = n1 - n2;
        const float determinant = duv.x * duv.w - duv.y * duv.z;
        if (pabs(determinant) < 1e-8f) {
            const V3 dn = cross(n2 - n0, n1 - n0);
            if (length_squared(dn) != 0.0f) coordinate_system(dn, &du, &dv);
        } else {
            const float invdet = 1.0f / determinant;
            du = (duv.w * dn1 - duv.y * dn2) * invdet;
            dv = (-duv.z * dn1 + duv.x * dn2) * invdet;
        }
        if (xf) {
            const float* mi = S.instances[inst].w2i;
            du = mk(mi[0] * du.x + mi[4] * du.y + mi[8] * du.z, mi[1] * du.x + mi[5] * du.y + mi[9] * du.z, mi[2] * du.x + mi[6] * du.y + mi[10] * du.z);
            dv = mk(mi[0] * dv.x + mi[4] * dv.y + mi[8] * dv.z, mi[1] * dv.x + mi[5] * dv.y + mi[9] * dv.z, mi[2] * dv.x + mi[6] * dv.y + mi[10] * dv.z);
        }
    }
    *dndu = du; *dndv = dv;
}

// Light::sample_li for the three light kinds of this path (point.rs:83-94, diffuse.rs:114-129 over Triangle::sample
// and Shape::sample_solid_angle, infinite.rs:133-175) plus the light-side endpoint of the VisibilityTester.
struct LightSample {
    bool valid;
    V3 wi, p1, p1_err, p1_n;
    float pdf;
    RGB Li;
    V3 q0, q1, q2;  // area light: its triangle
    uint32_t lflags;
};
B2_D LightSample sample_light(const DeviceScene& S, const DLight& light, const SurfHit& sh, P2 u_light) {
    bool li_valid = false;
    V3 wi = mk(0.0f, 0.0f, 0.0f), lp1 = wi, lp1_err = wi, lp1_n = wi;
    float light_pdf = 0.0f;
    RGB Li = rgb1(0.0f);
    V3 q0 = mk(0, 0, 0), q1 = q0, q2 = q0;  // area light triangle
    bool lflip = false;
    uint32_t lflags = 0;
    if (light.type == LT_POINT) {  // point.rs:83-94
        V3 pl = mk(light.pos[0], light.pos[1], light.pos[2]);
        wi = normalize(pl - sh.p);
        light_pdf = 1.0f;
        lp1 = pl;
        Li = ldrgb(light.L) / distance_squared(pl, sh.p);
        li_valid = true;
    } else if (light.type == LT_PROJECTION) {  // projection.rs:160-171
        V3 pl = mk(light.pos[0], light.pos[1], light.pos[2]);
        wi = normalize(pl - sh.p);
        light_pdf = 1.0f;
        lp1 = pl;
        Li = ldrgb(light.L) * projection_scale(light, S.inf_distr[light.inf_slot], -wi) / distance_squared(pl, sh.p);
        li_valid = true;
    } else if (light.type == LT_GONIO) {  // goniometric.rs:127-138
        V3 pl = mk(light.pos[0], light.pos[1], light.pos[2]);
        wi = normalize(pl - sh.p);
        light_pdf = 1.0f;
        lp1 = pl;
        Li = ldrgb(light.L) * gonio_scale(light, S.inf_distr[light.inf_slot], -wi) / distance_squared(pl, sh.p);
        li_valid = true;
    } else if (light.type == LT_SPOT) {  // spot.rs:97-107, falloff() :62-76
        V3 pl = mk(light.pos[0], light.pos[1], light.pos[2]);
        wi = normalize(pl - sh.p);
        light_pdf = 1.0f;
        lp1 = pl;
        const V3 w = -wi;
        const float* m = light.w2l;
        const V3 wl = normalize(mk(m[0] * w.x + m[1] * w.y + m[2] * w.z, m[3] * w.x + m[4] * w.y + m[5] * w.z, m[6] * w.x + m[7] * w.y + m[8] * w.z));
        const float cos_theta = wl.z, cos_total = light.area, cos_start = light.cos_falloff_start;
        float falloff;
        if (cos_theta < cos_total) falloff = 0.0f;
        else if (cos_theta >= cos_start) falloff = 1.0f;
        else {
            const float delta = (cos_theta - cos_total) / (cos_start - cos_total);
            falloff = (delta * delta) * (delta * delta);
        }
        Li = ldrgb(light.L) * falloff / distance_squared(pl, sh.p);
        li_valid = true;
    } else if (light.type == LT_DISTANT) {  // distant.rs:81-90: p_outside = p + w_light * (2 * world_radius), pdf 1
        wi = mk(light.pos[0], light.pos[1], light.pos[2]);
        light_pdf = 1.0f;
        lp1 = sh.p + wi * (2.0f * S.world_radius);
        Li = ldrgb(light.L);
        li_valid = true;
    } else if (light.type == LT_AREA) {
        int m2, l2i; uint32_t f2;
        load_prim(S, (uint32_t)light.prim, &q0, &q1, &q2, &m2, &l2i, &f2);
        lflip = (f2 & 1u) != 0;
        lflags = f2;
        // Triangle::sample (triangle.rs:918-949) + Shape::sample_solid_angle (shape.rs:64-79)
        float su0 = sqrtf(u_light.x);
        float bx = 1.0f - su0, by = u_light.y * su0;
        V3 p = bx * q0 + by * q1 + (1.0f - bx - by) * q2;
        V3 n = normalize(cross(q1 - q0, q2 - q0));
        if (S.prim_n && (f2 & B200PT_PRIM_HAS_NORMALS)) {  // triangle.rs:931-937: orient like intersect() does
            const float* vn = S.prim_n + 9ll * light.prim;
            V3 ns = bx * mk(vn[0], vn[1], vn[2]) + by * mk(vn[3], vn[4], vn[5]) + (1.0f - bx - by) * mk(vn[6], vn[7], vn[8]);
            n = face_forward(n, ns);
        } else if (lflip) n = -1.0f * n;
        V3 pas = vabs(bx * q0) + vabs(by * q1) + vabs((1.0f - bx - by) * q2);
        V3 p_err = kGamma6 * pas;
        float pdf = 1.0f / light.area;
        V3 w = p - sh.p;
        if (length_squared(w) == 0.0f) pdf = 0.0f;
        else {
            w = normalize(w);
            pdf *= distance_squared(sh.p, p) / abs_dot(n, -w);
            if (isinf(pdf)) pdf = 0.0f;
        }
        V3 w2 = p - sh.p;  // DiffuseAreaLight::sample_li, diffuse.rs:114-129
        float wl2 = length_squared(w2);
        if (!(pdf == 0.0f || wl2 == 0.0f)) {
            w2 = w2 / sqrtf(wl2);
            wi = w2; light_pdf = pdf;
            Li = area_l(light, n, -w2);
            lp1 = p; lp1_err = p_err; lp1_n = n;
            li_valid = true;
        }
    } else {  // InfiniteAreaLight::sample_li, infinite.rs:133-175
        const DInfDistr& D = S.inf_distr[light.inf_slot];
        float pdf1, pdf0; int v, dummy;
        float d1 = distr_sample_continuous(D.mfunc, D.mcdf, D.mfunc_int, D.nv, u_light.y, &pdf1, &v);
        float d0 = distr_sample_continuous(D.func + (long long)v * D.nu, D.cdf + (long long)v * (D.nu + 1), D.func_int[v], D.nu, u_light.x, &pdf0, &dummy);
        float map_pdf = pdf0 * pdf1;
        if (map_pdf != 0.0f) {
            float theta = d1 * kPi, phi = d0 * kTwoPi;
            float cos_t = lmx::cosf_glibc(theta), sin_t = lmx::sinf_glibc(theta);
            float sin_p = lmx::sinf_glibc(phi), cos_p = lmx::cosf_glibc(phi);
            wi = xf3(light.l2w, mk(sin_t * cos_p, sin_t * sin_p, cos_t));
            light_pdf = map_pdf / (kTwoPi * kPi * sin_t);
            if (sin_t == 0.0f) light_pdf = 0.0f;
            lp1 = sh.p + wi * (2.0f * S.world_radius);
            Li = inf_lookup(D, mk2(d0, d1));
            li_valid = true;
        }
    }
    LightSample r;
    r.valid = li_valid; r.wi = wi; r.p1 = lp1; r.p1_err = lp1_err; r.p1_n = lp1_n; r.pdf = light_pdf; r.Li = Li;
    r.q0 = q0; r.q1 = q1; r.q2 = q2; r.lflags = lflags;
    return r;
}

// estimate_direct (core/src/integrator/common.rs:146-299) up to the two rays it traces: the light-sampling half gives
// ld_light (added if the shadow ray is unoccluded), the BSDF-sampling half (non-delta lights) gives f, the MIS weight and
// the pdf of a closest-hit ray whose hit decides whether the light is seen (resolved later, k_resolve).
struct DirectEst {
    RGB ld_light, mis_f;
    float mis_w, mis_pdf;
    bool shadow, mis;
    V3 sh_o, sh_d, mis_o, mis_d;
};
template <uint32_t KM = KM_ALL> B2_D DirectEst estimate_direct_rays(const DeviceScene& S, const DLight& light, const SurfHit& sh, V3 hit_wo, const BSDF& bsdf, P2 u_light, P2 u_scatter) {
    const uint32_t kNoSpec = BSDF_ALL & ~BSDF_SPECULAR;
    // ---- estimate_direct (common.rs:146-299), light-sampling half ----
    DirectEst r;
    r.shadow = false; r.mis = false;
    r.sh_o = r.sh_d = r.mis_o = r.mis_d = mk(0.0f, 0.0f, 0.0f);
    RGB ld_light = rgb1(0.0f);
    const LightSample ls = sample_light(S, light, sh, u_light);
    const bool li_valid = ls.valid;
    const V3 wi = ls.wi, lp1 = ls.p1, lp1_err = ls.p1_err, lp1_n = ls.p1_n;
    const float light_pdf = ls.pdf;
    const RGB Li = ls.Li;
    const V3 q0 = ls.q0, q1 = ls.q1, q2 = ls.q2;
    const uint32_t lflags = ls.lflags;
    float scattering_pdf = 0.0f;
    if (li_valid && light_pdf > 0.0f && !is_black(Li)) {
        RGB f = bsdf_f<KM>(bsdf, hit_wo, wi, kNoSpec) * abs_dot(wi, sh.ns);
        scattering_pdf = bsdf_pdf<KM>(bsdf, hit_wo, wi, kNoSpec);
        if (!is_black(f)) {
            // VisibilityTester -> Hit::spawn_ray_to_hit (interaction/mod.rs:212-223)
            V3 origin = offset_ray_origin(sh.p, sh.p_error, sh.n, lp1 - sh.p);
            V3 target = offset_ray_origin(lp1, lp1_err, lp1_n, origin - lp1);
            r.shadow = true; r.sh_o = origin; r.sh_d = target - origin;
            if (light_is_delta(light.type)) ld_light = f * Li / light_pdf;
            else {
                float wgt = power_heuristic(light_pdf, scattering_pdf);
                ld_light = f * Li * wgt / light_pdf;
            }
        }
    }
    // ---- BSDF-sampling half (non-delta lights only) ----
    RGB mis_f = rgb1(0.0f);
    float mis_w = 1.0f, mis_pdf = 0.0f;
    if (!light_is_delta(light.type)) {
        BxDFSample bs = bsdf_sample_f<KM>(bsdf, hit_wo, u_scatter, kNoSpec);
        V3 wi2 = bs.wi;
        RGB f = bs.f * abs_dot(wi2, sh.ns);
        bool sampled_specular = (bs.type & BSDF_SPECULAR) != 0;
        if (!is_black(f) && bs.pdf > 0.0f) {
            float weight = 1.0f;
            bool ok = true;
            V3 ro = offset_ray_origin(sh.p, sh.p_error, sh.n, wi2);  // Hit::spawn_ray
            if (!sampled_specular) {
                float lp;
                if (light.type == LT_AREA) {  // Shape::pdf_solid_angle, shape.rs:81-107
                    TriCtx tc = make_tri_ctx(wi2.x, wi2.y, wi2.z);
                    float t, c0, c1, c2;
                    lp = 0.0f;
                    const float4 lduv = (S.prim_duv && (lflags & B200PT_PRIM_HAS_UV)) ? ldg4(S.prim_duv + light.prim) : default_duv();
                    if (triangle_test(ro, tc, __int_as_float(0x7f800000), q0, q1, q2, &t, &c0, &c1, &c2) && triangle_nondegenerate(q0, q1, q2, lduv)) {
                        const float* lvn = (S.prim_n && (lflags & B200PT_PRIM_HAS_NORMALS)) ? S.prim_n + 9ll * light.prim : nullptr;
                    const float* lvs = (S.prim_s && (lflags & B200PT_PRIM_HAS_TANGENTS)) ? S.prim_s + 9ll * light.prim : nullptr;
                    SurfHit lh = triangle_surface(q0, q1, q2, c0, c1, c2, lflags, lduv, lvn, lvs);
                        lp = distance_squared(sh.p, lh.p) / (abs_dot(lh.n, -wi2) * light.area);
                        if (isinf(lp)) lp = 0.0f;
                    }
                } else {  // infinite.rs:201-211
                    V3 w = xf3(light.w2l, wi2);
                    float theta = spherical_theta(w), phi = spherical_phi(w);
                    float sin_t = lmx::sinf_glibc(theta);
                    if (sin_t == 0.0f) lp = 0.0f;
                    else {
                        const DInfDistr& D = S.inf_distr[light.inf_slot];
                        lp = distr2d_pdf(D, phi * kInvTwoPi, theta * kInvPi) / (kTwoPi * kPi * sin_t);
                    }
                }
                if (lp == 0.0f) ok = false;  // common.rs:258-260: return ld
                else weight = power_heuristic(bs.pdf, lp);
            }
            if (ok) {
                r.mis = true; r.mis_o = ro; r.mis_d = wi2;
                mis_f = f; mis_w = weight; mis_pdf = bs.pdf;
            }
        }
    }
    r.ld_light = ld_light; r.mis_f = mis_f; r.mis_w = mis_w; r.mis_pdf = mis_pdf;
    return r;
}

// ---- SpatialLightDistribution lookup (core/src/light_distrib/spatial.rs:166-180) -------------------------------------
B2_D int ld_volatile_int(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
// lookup(), spatial.rs:166-180: voxel of a point (Bounds3::offset, `as Int` truncation, clamp)
B2_D int spatial_voxel(const DeviceScene& S, V3 p) {
    int pi[3];
    const float pc[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float o = pc[i] - S.wb[i];
        if (S.wb[3 + i] > S.wb[i]) o /= S.wb[3 + i] - S.wb[i];
        float v = o * (float)S.n_voxels[i];
        int q = !(v == v) ? 0 : (v >= 2147483648.0f ? 0x7fffffff : (v <= -2147483648.0f ? (int)0x80000000 : (int)v));
        pi[i] = q < 0 ? 0 : (q > S.n_voxels[i] - 1 ? S.n_voxels[i] - 1 : q);
    }
    return (pi[0] * S.n_voxels[1] + pi[1]) * S.n_voxels[2] + pi[2];
}
B2_D float lerp_ref(float t, float a, float b) { return (1.0f - t) * a + t * b; }  // pbrt::lerp, common.rs


// Geometric normal of a hit on an emissive triangle as Triangle::intersect leaves it (face-forwarded to the shading normal
// when the mesh has vertex normals / tangents, triangle.rs:625-721), for DiffuseAreaLight::l at the end of a MIS ray.
B2_D V3 emitter_hit_normal(const DeviceScene& S, uint32_t prim, V3 p0, V3 p1, V3 p2, uint32_t fl, float b0, float b1, float b2) {
    if (fl & (B200PT_PRIM_HAS_NORMALS | B200PT_PRIM_HAS_TANGENTS)) {
        const float4 duv = (S.prim_duv && (fl & B200PT_PRIM_HAS_UV)) ? ldg4(S.prim_duv + prim) : default_duv();
        const float* vn = (S.prim_n && (fl & B200PT_PRIM_HAS_NORMALS)) ? S.prim_n + 9ll * prim : nullptr;
        const float* vs = (S.prim_s && (fl & B200PT_PRIM_HAS_TANGENTS)) ? S.prim_s + 9ll * prim : nullptr;
        return triangle_surface(p0, p1, p2, b0, b1, b2, fl, duv, vn, vs).n;
    }
    V3 n = normalize(cross(p0 - p2, p1 - p2));
    return (fl & 1u) ? -n : n;
}

}  // namespace b2
