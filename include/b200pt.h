/*
 * b200pt.h — C ABI of the B200-native replacement for the ray-scene
 * intersection + path-integration hot path of hackmad/pbrt-v3-rs.
 *
 * This is the drop-in boundary: exactly what a `b200pt-sys`
 * FFI crate on the reference side would bind.  Plain pointers and sizes, no
 * C++/torch types, no exceptions across the boundary.  Every call returns 0 on
 * success or a negative b200pt_status; b200pt_last_error() gives the message
 * (thread-local).  There is NO CPU fallback: without an sm_100 device
 * b200pt_init() fails and every compute entry point returns
 * B200PT_ERR_NO_DEVICE.
 *
 * file:line citations are relative to the reference repository root.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b200pt_status {
    B200PT_OK = 0,
    B200PT_ERR_NO_DEVICE = -1,   /* no CUDA device / not sm_100 */
    B200PT_ERR_INVALID = -2,     /* bad argument */
    B200PT_ERR_CUDA = -3,        /* CUDA runtime error, see last_error */
    B200PT_ERR_OOM = -4,
    B200PT_ERR_UNSUPPORTED = -5
} b200pt_status;

#define B200PT_MISS 0xffffffffu

/* core/src/geometry/ray.rs:10-28 (o, d, t_max, time; differentials and medium
 * are not carried on this path). 32 bytes, two 16-byte vectors. */
typedef struct b200pt_ray {
    float o[3];
    float tmax;
    float d[3];
    float time;
} b200pt_ray;

/* Device form of the SurfaceInteraction returned by BVHAccel::intersect
 * (accelerators/src/bvh/mod.rs:173-226): hit distance, ORIGINAL primitive
 * index (position in RenderOptions.primitives; B200PT_MISS when nothing was
 * hit) and the first two barycentrics. 16 bytes. */
typedef struct b200pt_hit {
    float t;
    uint32_t prim;
    float b0;
    float b1;
} b200pt_hit;

/* accelerators/src/bvh/common.rs:163-179 LinearBVHNode, fixed to 32 bytes. */
typedef struct b200pt_bvh_node {
    float bounds[6];       /* p_min.xyz, p_max.xyz */
    uint32_t offset;       /* leaf: first ordered primitive; interior: second child */
    uint16_t n_primitives; /* > 0 => leaf */
    uint8_t axis;
    uint8_t pad;
} b200pt_bvh_node;

/* Per-primitive flags. */
#define B200PT_PRIM_FLIP_NORMAL 1u        /* reverse_orientation ^ transform_swaps_handedness (shapes/src/triangle.rs:625-629) */
#define B200PT_PRIM_ALPHA_ZERO 2u         /* constant "alpha" texture == 0 (triangle.rs:587-607) */
#define B200PT_PRIM_SHADOW_ALPHA_ZERO 4u  /* constant "shadowalpha" texture == 0 (triangle.rs:840-899) */
#define B200PT_PRIM_REVERSE_ORIENTATION 8u /* reverse_orientation alone: flips the shading bitangent (triangle.rs:714-716) */
#define B200PT_PRIM_HAS_UV 16u            /* the primitive's mesh has "uv"/"st": tri_uvs holds its three uvs (triangle.rs:384-394) */
#define B200PT_PRIM_HAS_NORMALS 32u       /* ... has "N": tri_normals (triangle.rs:631-653) */
#define B200PT_PRIM_HAS_TANGENTS 64u      /* ... has "S": tri_tangents (triangle.rs:655-670) */
#define B200PT_PRIM_ALPHA_TEXTURE 128u    /* the mesh's "alpha" or "shadowalpha" is a non-constant float texture: prim_alpha_tex names it (triangle.rs:278-312) */

/* materials/src/{matte,plastic,glass,metal,mirror}.rs with constant textures (matte / plastic "Kd" may be textured, see
 * b200pt_spectrum_texture). */
enum { B200PT_MAT_MATTE = 0, B200PT_MAT_PLASTIC = 1, B200PT_MAT_GLASS = 2, B200PT_MAT_METAL = 3, B200PT_MAT_MIRROR = 4 };
typedef struct b200pt_material {
    int32_t type;
    float kd[3];    /* matte/plastic "Kd" */
    float ks[3];    /* plastic "Ks"; glass "Kr"; mirror "Kr" (default 0.9): SpecularReflection with FresnelNoOp, mirror.rs:47-52 */
    float kt[3];    /* glass "Kt" */
    float eta[3];   /* metal "eta" (RGB); glass "index"/"eta" in eta[0] */
    float k[3];     /* metal "k" */
    float sigma;    /* matte "sigma" (degrees) */
    float urough;   /* plastic/metal "roughness", glass/metal "uroughness" — as written in the scene file */
    float vrough;
    int32_t remap_roughness;
} b200pt_material;

/* lights/src/{point,diffuse,infinite}.rs */
enum { B200PT_LIGHT_POINT = 0, B200PT_LIGHT_AREA = 1, B200PT_LIGHT_INFINITE = 2,
       B200PT_LIGHT_DISTANT = 3, /* lights/src/distant.rs: pos = w_light, the NORMALISED world-space direction TOWARDS the light
                                   * (light_to_world.transform_vector(from - to).normalize(), distant.rs:50-52), L = L * scale */
       B200PT_LIGHT_GONIOMETRIC = 5, /* lights/src/goniometric.rs: pos = p_light, L = I * scale, world_to_light, and the "mapname" image in
                                      * map_rgb (decoded, as for an infinite light): the intensity in direction w is scaled by the image
                                      * at (phi, theta) of world_to_light(w) with y and z swapped - MIPMap::lookup_triangle(st, 0) (:101-115) */
       B200PT_LIGHT_PROJECTION = 6, /* lights/src/projection.rs: pos = p_light, L = I * scale, world_to_light, the image in map_rgb (NULL: 1),
                                     * fov = "fov" in degrees (the projection is Transform::perspective(fov, 1e-3, 1e30); the screen window follows
                                     * from the image's aspect, :84-88), cos_total_width = z of the normalised screen corner (:92-95; power() only) */
       B200PT_LIGHT_SPOT = 4 /* lights/src/spot.rs: pos = p_light = light_to_world(0), L = I * scale, world_to_light = the inverse of
                              * ctm * Translate(from) * dir_to_z^-1 (spot.rs:209-226), cos_total_width / cos_falloff_start =
                              * cos(radians(coneangle)), cos(radians(coneangle - conedeltaangle)) (spot.rs:57-58) */ };
typedef struct b200pt_light {
    int32_t type;
    float pos[3];             /* point: p_light (world) */
    float L[3];               /* point: I*scale; area: L*scale; infinite: L*scale */
    int32_t prim;             /* area: ORIGINAL primitive index of the emitting triangle (api/src/lib.rs:783-803: one light per triangle) */
    int32_t two_sided;        /* area */
    float light_to_world[16]; /* infinite (row-major 4x4) */
    float world_to_light[16];
    /* infinite: the "mapname" image, already decoded (image IO stays with the caller): map_width x map_height RGB
     * texels, row-major from the top row, NOT yet multiplied by L (lights/src/infinite.rs:66-79).  NULL = no map
     * (the reference then uses the 1x1 image [L]). */
    const float* map_rgb;
    int32_t map_width, map_height;
    float cos_total_width, cos_falloff_start; /* spot; projection: cos_total_width */
    float fov;                                /* projection: "fov" (degrees) */
    int32_t pad_;
} b200pt_light;

/* cameras/src/perspective_camera.rs + core/src/camera.rs:276-306: the two
 * matrices the reference's PerspectiveCamera holds (row-major 4x4). */
/* B200PT_CAMERA_ORTHOGRAPHIC: cameras/src/orthographic_camera.rs (same two matrices; camera_to_screen = Transform::orthographic(0, 1)).
 * B200PT_CAMERA_ENVIRONMENT: cameras/src/environment_camera.rs (direction from the film position over film.xres x film.yres;
 * raster_to_camera is not read; differentials by Camera::generate_ray_differential's finite differences, core/src/camera.rs:29-78). */
enum { B200PT_CAMERA_PERSPECTIVE = 0, B200PT_CAMERA_ORTHOGRAPHIC = 1, B200PT_CAMERA_ENVIRONMENT = 2 };
typedef struct b200pt_camera {
    float raster_to_camera[16];
    float camera_to_world[16];
    float lens_radius;
    float focal_distance;
    float shutter_open;
    float shutter_close;
    int32_t type;              /* B200PT_CAMERA_* */
} b200pt_camera;

/* core/src/film/mod.rs:89-146 */
typedef struct b200pt_film {
    int32_t xres, yres;
    int32_t crop[4];           /* cropped_pixel_bounds: x0, y0, x1, y1 */
    float filter_radius[2];
    float filter_table[256];   /* 16x16, film/mod.rs:113-125 */
    float scale;
    float max_sample_luminance; /* INFINITY when unset */
} b200pt_film;

/* B200PT_SAMPLER_SOBOL: SobolSampler (samplers/src/sobol.rs; pixelsamples rounded up to a power of two).  The scene
 * description must then carry the generator matrices (b200pt_scene_desc.sobol_matrices_32). */
enum { B200PT_SAMPLER_HALTON = 0, B200PT_SAMPLER_ZEROTWO = 1, B200PT_SAMPLER_SOBOL = 2 };
typedef struct b200pt_sampler {
    int32_t type;
    int32_t spp;               /* "pixelsamples" */
    int32_t sample_at_center;  /* halton "samplepixelcenter" */
    int32_t dimensions;        /* 02sequence "dimensions" */
} b200pt_sampler;

/* B200PT_LIGHTS_SPATIAL: SpatialLightDistribution (core/src/light_distrib/spatial.rs), the path integrator's default
 * "lightsamplestrategy".  Deterministic here: every lookup gets the voxel's distribution, whereas the reference's
 * lock-free table returns None (-> uniform light sampling) to threads that look a voxel up while another thread is
 * still computing it, i.e. this is the reference's single-thread result. */
enum { B200PT_LIGHTS_UNIFORM = 0, B200PT_LIGHTS_POWER = 1, B200PT_LIGHTS_SPATIAL = 2 };
/* B200PT_INTEGRATOR_PATH: PathIntegrator (integrators/src/path.rs:103-326).
 * B200PT_INTEGRATOR_WHITTED: WhittedIntegrator (integrators/src/whitted.rs:60-158 with specular_reflect /
 * specular_transmit of core/src/integrator/sampler_integrator.rs:79-238); reads max_depth and pixel_bounds only and
 * needs the halton sampler (the number of sampler dimensions a camera sample consumes is not bounded in advance).
 * B200PT_INTEGRATOR_DIRECT: DirectLightingIntegrator (integrators/src/direct_lighting.rs:82-146) with "strategy"
 * all | one; as the reference's tile samplers come from clone_sampler(), which drops the requested sample arrays
 * (samplers/src/halton.rs:176-182), "all" takes one MIS estimate_direct per light from plain get_2d() pairs. */
enum { B200PT_INTEGRATOR_PATH = 0, B200PT_INTEGRATOR_WHITTED = 1, B200PT_INTEGRATOR_DIRECT = 2 };
enum { B200PT_DIRECT_ALL = 0, B200PT_DIRECT_ONE = 1 };
typedef struct b200pt_integrator {
    int32_t max_depth;         /* "maxdepth" (5) */
    float rr_threshold;        /* "rrthreshold" (1.0) */
    int32_t pixel_bounds[4];   /* x0,y0,x1,y1 (sample bounds ∩ "pixelbounds") */
    int32_t light_strategy;    /* "lightsamplestrategy": uniform | power */
    int32_t type;              /* B200PT_INTEGRATOR_* */
    int32_t direct_strategy;   /* B200PT_DIRECT_* (directlighting "strategy") */
} b200pt_integrator;

/* ObjectBegin/ObjectEnd: the triangles [first_prim, first_prim + n_prims) of the scene's per-primitive arrays with their
 * own BVHAccel (nodes / ordered_prims index the object's triangles locally, 0-based) — api/src/lib.rs:952-968. */
typedef struct b200pt_object {
    const b200pt_bvh_node* nodes;
    int64_t n_nodes;
    const uint32_t* ordered_prims;
    int64_t first_prim;
    int64_t n_prims;
} b200pt_object;

/* ObjectInstance -> TransformedPrimitive with a static transform (core/src/primitives/transformed_primitive.rs:16-73):
 * primitive_to_world and its inverse, row-major 4x4, affine (last row 0 0 0 1). */
typedef struct b200pt_instance {
    int32_t object;
    float instance_to_world[16];
    float world_to_instance[16];
} b200pt_instance;

/* Float textures a mesh can name as "alpha" / "shadowalpha" (shapes/src/triangle.rs:278-312), evaluated at the hit's
 * interpolated uv inside Triangle::intersect / intersect_p (triangle.rs:587-607, 840-899) through UVMapping2D
 * (core/src/texture/mapping/uv_2d.rs: st = (su * u + du, sv * v + dv)).  The interaction the reference builds for that
 * test has zero differentials, so a checkerboard point-samples (textures/src/checkerboard_2d.rs:62-84) and an image map is
 * the level-0 bilinear lookup MIPMap::triangle(0, st) for both filters (core/src/mipmap/mod.rs:212-311).  A hit whose
 * texture evaluates to exactly 0 is rejected.  Constant textures are the B200PT_PRIM_*_ALPHA_ZERO flag bits instead. */
#define B200PT_TEX_CONSTANT 0
#define B200PT_TEX_CHECKERBOARD 1 /* 2-D: value = {tex1, tex2} constants (checkerboard_2d.rs) */
#define B200PT_TEX_DOTS 2         /* value = {outside_dot, inside_dot} as DotsTexture STORES them: its from-params hands
                                   * ("inside", "outside") to new(outside_dot, inside_dot) (textures/src/dots.rs:31,86), so
                                   * the scene-file parameter "inside" is value[0] and colours the outside of the dots */
#define B200PT_TEX_IMAGEMAP 3     /* texels = pyramid level 0 of the MIPMap<Float> (after convert_in's scale / gamma / y(),
                                   * the vertical flip and the power-of-two resampling of MIPMap::new), row t, column s */
typedef struct b200pt_float_texture {
    int32_t type;
    float su, sv, du, dv; /* "uscale" "vscale" "udelta" "vdelta" (textures/src/lib.rs:47-52) */
    float value[2];
    int32_t wrap;         /* imagemap: 0 repeat, 1 black, 2 clamp (core/src/mipmap/mod.rs:580-608) */
    int32_t width, height;
    const float* texels;
} b200pt_float_texture;

/* Spectrum textures a matte / plastic material can name as "Kd" (materials/src/matte.rs:63, plastic.rs:81:
 * kd.evaluate(&si.hit, &si.uv, &si.der).clamp_default() once per intersection), through UVMapping2D.  A 2-D
 * checkerboard with "aamode" closedform (the default, textures/src/checkerboard_2d.rs:62-98) box-filters over the
 * footprint (dudx, dvdx, dudy, dvdy) that SurfaceInteraction::compute_differentials (surface_interaction.rs:203-277)
 * derives from the ray's differentials: only CAMERA rays carry them (PerspectiveCamera::generate_ray_differential,
 * perspective_camera.rs:144-204, scaled by 1 / sqrt(spp), sampler_integrator.rs:357-358) and hand them on to the specular
 * children of the Whitted / DirectLighting trees; rays spawned by the path integrator have none and point-sample.  tex1 / tex2 are constants here (nested textures are not supported). */
#define B200PT_STEX_CONSTANT 0
#define B200PT_STEX_CHECKERBOARD 1
typedef struct b200pt_spectrum_texture {
    int32_t type;
    float su, sv, du, dv;  /* "uscale" "vscale" "udelta" "vdelta" */
    float tex1[3];         /* constant: the value; checkerboard: "tex1" (default 1) */
    float tex2[3];         /* checkerboard: "tex2" (default 0) */
    int32_t aa_closedform; /* checkerboard "aamode": 1 = closedform (default), 0 = none */
} b200pt_spectrum_texture;

typedef struct b200pt_scene_desc {
    const b200pt_bvh_node* nodes;
    int64_t n_nodes;
    const uint32_t* ordered_prims; /* BVHAccel.primitives order: ordered position -> original index */
    const float* tri_verts;        /* 9 floats per ORIGINAL primitive, world space (TriangleMesh::new transforms to world, triangle.rs:92) */
    const uint32_t* prim_flags;    /* per original primitive, may be NULL */
    const int32_t* prim_material;  /* per original primitive index into materials; -1 = no material (Material "" / "none":
                                    * the path integrator passes through the surface without counting a bounce, path.rs:146-150) */
    const int32_t* prim_light;     /* per original primitive index into lights or -1, may be NULL */
    int64_t n_prims;
    const b200pt_material* materials;
    int32_t n_materials;
    const b200pt_light* lights;
    int32_t n_lights;
    b200pt_camera camera;
    b200pt_film film;
    b200pt_sampler sampler;
    b200pt_integrator integrator;
    /* Instancing (optional; n_objects == 0 => every primitive is a top-level triangle).  With objects, the top-level
     * `nodes` / `ordered_prims` are built over n_top_tris + n_instances primitives: index p < n_top_tris is triangle p,
     * otherwise instance p - n_top_tris; triangles [n_top_tris, n_prims) belong to the objects. */
    int64_t n_top_tris;
    const b200pt_object* objects;
    int32_t n_objects;
    const b200pt_instance* instances;
    int32_t n_instances;
    /* Optional vertex attributes, de-indexed like tri_verts and already in world space (TriangleMesh::new applies
     * transform_normal / transform_vector and does not renormalise, triangle.rs:92-99).  Each may be NULL; a primitive
     * uses an array only when its flag bit (B200PT_PRIM_HAS_*) is set. */
    const float* tri_uvs;      /* 6 floats per ORIGINAL primitive: uv0 uv1 uv2 */
    const float* tri_normals;  /* 9 floats per ORIGINAL primitive: n0 n1 n2 */
    const float* tri_tangents; /* 9 floats per ORIGINAL primitive: s0 s1 s2 */
    /* Sobol sampler only: SOBOL_MATRICES_32 of core/src/sobol_matrices.rs, 1024 dimensions x 52 u32 (the Rust side passes
     * the constant's address; the Python mirror and the scene-file loader read pbrt-v3-rs_b200/data/sobol_matrices_32.bin).
     * NULL otherwise. */
    const uint32_t* sobol_matrices_32;
    /* Alpha masks (optional): prim_alpha_tex = 2 ints per ORIGINAL primitive, the index into float_textures of its mesh's
     * "alpha" and "shadowalpha" texture or -1 (none / constant, see the flag bits); read for primitives whose flags carry
     * B200PT_PRIM_ALPHA_TEXTURE.  noise_perm = NOISE_PERM[0..256) of core/src/texture/common.rs (Perlin's permutation;
     * needed by "dots" textures only: the Rust side passes the constant, the mirror and the loader read
     * pbrt-v3-rs_b200/data/noise_perm.bin). */
    const b200pt_float_texture* float_textures;
    int32_t n_float_textures;
    const int32_t* prim_alpha_tex;
    const uint8_t* noise_perm;
    /* Textured "Kd" (optional): material_kd_tex = per material the index into spectrum_textures of its "Kd" texture or -1
     * (the constant b200pt_material.kd); read for matte and plastic materials.  NULL = every Kd is constant.  Needs
     * tri_uvs for meshes with "uv" / "st" (others use the default (0,0) (1,0) (1,1), triangle.rs:384-394).  Under Whitted /
     * DirectLighting the specular children carry the reflected / refracted differentials of specular_reflect /
     * specular_transmit (sampler_integrator.rs:108-125, 164-227). */
    const b200pt_spectrum_texture* spectrum_textures;
    int32_t n_spectrum_textures;
    const int32_t* material_kd_tex;
} b200pt_scene_desc;

typedef struct b200pt_accel b200pt_accel; /* opaque: device-resident BVHAccel */
typedef struct b200pt_loaded_scene b200pt_loaded_scene; /* opaque: a parsed scene file, owns the arrays of its scene_desc */
typedef struct b200pt_scene b200pt_scene; /* opaque: device-resident Scene + PathIntegrator state */

/* ---- library ----------------------------------------------------------
 * b200pt_init(device) checks that `device` is an sm_100 GPU, registers it and binds the CALLING THREAD to it; it may
 * be called once per device of the box.  The first device initialised is the default of threads that never called
 * b200pt_init / b200pt_set_device.  Creation calls (b200pt_accel_create*, b200pt_scene_create, the *_gpu / *_device
 * BVH builders) work on the calling thread's device; every handle remembers the device it was created on and each
 * call on it switches to that device, so one process - the reference is one process, bin/src/main.rs:29-85 - can
 * drive all GPUs of the box (b200pt_render_multi). */
int b200pt_init(int device);
int b200pt_set_device(int device);   /* rebinds the calling thread to an initialised device */
int b200pt_current_device(void);     /* the calling thread's device, -1 before any b200pt_init */
const char* b200pt_last_error(void);
int b200pt_version(void);
/* number of SMs / L2 bytes of the bound device (0 before init) */
int b200pt_device_sm_count(void);
int64_t b200pt_device_l2_bytes(void);

/* ---- host-side BVH build: BVHAccel::new with SplitMethod::SAH ----------
 * (accelerators/src/bvh/mod.rs:43-153, sah.rs:26-367).  prim_bounds = 6
 * floats per primitive (Primitive::world_bound).  nodes_out needs room for
 * 2*n-1 nodes, ordered_out for n indices.  *n_nodes_out receives the node
 * count.  max_prims_in_node is the reference's u8 "maxnodeprims". */
int b200pt_bvh_build_sah(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                         int64_t* n_nodes_out, uint32_t* ordered_out);
/* Triangle::world_bound (shapes/src/triangle.rs:427-431) for n triangles. */
int b200pt_triangle_bounds(const float* tri_verts, int64_t n, float* bounds_out);

/* BVHAccel::new with SplitMethod::HLBVH (accelerators/src/bvh/hlbvh.rs:33-449, morton.rs:37-120): same argument
 * meaning as b200pt_bvh_build_sah.  Keeps two reference behaviours that show in the result: encode_morton_3
 * interleaves bits of the float BIT PATTERN of the scaled centroid offset (morton.rs:43-49, float_to_bits is a
 * transmute; release-build semantics), and treelets are emitted in order (the reference's --nthreads 1 primitive order).
 * Returns B200PT_ERR_INVALID where the reference asserts. */
int b200pt_bvh_build_hlbvh(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                           int64_t* n_nodes_out, uint32_t* ordered_out);
/* the reference's Morton code of every primitive (before sorting) */
int b200pt_hlbvh_morton_codes(const float* prim_bounds, int64_t n, uint32_t* codes_out);

/* ---- the same build on the GPU (csrc/bvh_build.cu) ----------------------
 * Results (node bytes, ordered_prims) are identical to b200pt_bvh_build_sah and therefore to the reference's
 * BVHAccel::new(.., SplitMethod::SAH) (mod.rs:43-153, sah.rs:26-367): the level-parallel schedule only reorders
 * min/max reductions and restates itertools::partition (sah.rs:354) through a prefix sum.  n < 2^31.
 * _gpu: host buffers in / out (drop-in for b200pt_bvh_build_sah).  _device: device pointers, `stream` is a
 * cudaStream_t (NULL = default stream); returns after the build has completed on that stream. */
int b200pt_bvh_build_sah_gpu(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                             int64_t* n_nodes_out, uint32_t* ordered_out);
int b200pt_bvh_build_sah_device(const float* d_prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* d_nodes_out,
                                int64_t* n_nodes_out, uint32_t* d_ordered_out, void* stream);
int b200pt_triangle_bounds_device(const float* d_tri_verts, int64_t n, float* d_bounds_out, void* stream);
/* b200pt_bvh_build_hlbvh on the GPU (Morton codes, 5 x 6-bit stable radix passes, one thread per treelet; only the
 * <= 4096-treelet upper SAH layout runs on the host): device pointers in / out, same bytes as the host builder. */
int b200pt_bvh_build_hlbvh_device(const float* d_prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* d_nodes_out,
                                  int64_t* n_nodes_out, uint32_t* d_ordered_out, void* stream);
/* host buffers in / out: the drop-in for b200pt_bvh_build_hlbvh once a device is bound */
int b200pt_bvh_build_hlbvh_gpu(const float* prim_bounds, int64_t n, int max_prims_in_node, b200pt_bvh_node* nodes_out,
                               int64_t* n_nodes_out, uint32_t* ordered_out);
/* The builder keeps its device scratch (about 160 bytes per primitive of the largest build so far) between calls;
 * this frees it. */
int b200pt_bvh_build_release(void);

/* Host-only: what InfiniteAreaLight::new prepares for an environment image (lights/src/infinite.rs:61-92, 326-369;
 * core/src/mipmap/mod.rs): level 0 of the MIPMap (sides rounded up to powers of two by the Lanczos resampler) and the
 * (2w x 2h) importance image y * sin(theta) the light's Distribution2D is built over.  Call with the out pointers
 * NULL to get the sizes: size4 = {level0_width, level0_height, importance_width, importance_height}. */
int b200pt_envmap_prepare(const float* map_rgb, int32_t map_width, int32_t map_height, const float L[3], int32_t size4[4],
                          float* level0_rgb_out, float* importance_out, float power_lookup_out[3]);

/* Host-only: the pixel <-> sample-index tables SobolSampler uses for a 2^m x 2^m image (the reference's
 * VD_C_SOBOL_MATRICES[m - 1] / VD_C_SOBOL_MATRICES_INV[m - 1], core/src/low_discrepency.rs:1770-1808), derived from
 * dimensions 0 and 1 of SOBOL_MATRICES_32.  Exposed for the tests. */
int b200pt_sobol_interval_tables(const uint32_t* sobol_matrices_32, int m, uint64_t vdc_out[52], uint64_t vdc_inv_out[52]);

/* ---- scene ingestion (host only; api/src/lib.rs + api/src/parser, shapes/src/plymesh.rs, core/src/image_io.rs) ----
 * Reads the subset of the pbrt-v3 scene format that reaches this path (perspective / orthographic / environment camera;
 * image film; box / gaussian / triangle / mitchell / sinc filter; halton / 02sequence / sobol sampler; path / whitted /
 * directlighting integrator with uniform / power / spatial light sampling; bvh accelerator with splitmethod sah / hlbvh;
 * trianglemesh / plymesh shapes with P, N, S, uv/st, alpha, shadowalpha (float textures: constant, checkerboard, dots,
 * imagemap); matte / plastic / glass / metal / mirror with constant parameters (a matte / plastic "Kd" may be a constant or
 * checkerboard spectrum texture); point / spot / projection / goniometric / distant / infinite / diffuse area lights (image maps: .pfm and
 * 8-bit .png); rgb / color /
 * blackbody spectra; transforms, attribute and transform stacks, named materials, object instancing, Include) and builds the BVHs
 * with b200pt_bvh_build_sah (on the GPU once a device is bound) / b200pt_bvh_build_hlbvh.  Anything else returns
 * B200PT_ERR_UNSUPPORTED with the offending directive
 * in b200pt_last_error.  The returned desc stays valid until b200pt_loaded_scene_free. */
int b200pt_load_pbrt(const char* path, b200pt_loaded_scene** out);
const b200pt_scene_desc* b200pt_loaded_scene_desc(const b200pt_loaded_scene* s);
const char* b200pt_loaded_scene_output(const b200pt_loaded_scene* s); /* Film "filename" */
void b200pt_loaded_scene_free(b200pt_loaded_scene* s);
/* PFM images, rgb = width x height x 3 floats, top row first. read: call with rgb_out NULL to get size2 = {w, h}. */
int b200pt_write_pfm(const char* path, const float* rgb, int32_t width, int32_t height);
int b200pt_read_pfm(const char* path, float* rgb_out, int32_t size2[2]);
/* 8-bit sRGB PNG exactly as the reference encodes it (core/src/image_io.rs:291-390: clamp(255 * gamma_correct(v) + 0.5) as u8),
 * and write_image's dispatch on the file extension (.png, .pfm on this path). */
int b200pt_write_png(const char* path, const float* rgb, int32_t width, int32_t height);
int b200pt_write_image(const char* path, const float* rgb, int32_t width, int32_t height);

/* ---- accelerator: impl Primitive for BVHAccel --------------------------
 * Copies the arrays to the device; the caller keeps ownership of its own. */
int b200pt_accel_create(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered_prims, const float* tri_verts,
                        const uint32_t* prim_flags, int64_t n_prims, b200pt_accel** out);
/* Same for meshes with "uv"/"st": tri_uvs = 6 floats per ORIGINAL primitive (used by the primitives whose flags carry
 * B200PT_PRIM_HAS_UV).  The uvs reach the traversal through the degenerate-hit rejection of triangle.rs:551-572. */
int b200pt_accel_create_uv(const b200pt_bvh_node* nodes, int64_t n_nodes, const uint32_t* ordered_prims, const float* tri_verts,
                           const float* tri_uvs, const uint32_t* prim_flags, int64_t n_prims, b200pt_accel** out);
/* BVHAccel::new (SplitMethod::SAH) for triangles that already live on the device: bounds, the GPU SAH build and the
 * traversal records are all produced in HBM (csrc/bvh_build.cu); d_tri_verts = 9 floats per triangle, d_prim_flags may be
 * NULL (no per-mesh uvs on this entry point).  The tree is the reference's: b200pt_accel_download returns the same
 * LinearBVHNode array / ordered_prims as b200pt_bvh_build_sah.  Rebuilding a 1 M-triangle accelerator takes a few ms. */
int b200pt_accel_create_device(const float* d_tri_verts, int64_t n_prims, const uint32_t* d_prim_flags, int max_prims_in_node, void* stream,
                               b200pt_accel** out);
/* Alpha masks for a stand-alone accelerator (same arrays as in b200pt_scene_desc; tri_uvs may be NULL when no mesh has
 * uvs: Triangle::get_uvs then supplies (0,0) (1,0) (1,1), triangle.rs:384-394).  Call before tracing. */
int b200pt_accel_set_alpha_textures(b200pt_accel* a, const b200pt_float_texture* float_textures, int32_t n_float_textures,
                                    const int32_t* prim_alpha_tex, const float* tri_uvs, const uint32_t* prim_flags, const uint8_t* noise_perm);
int b200pt_accel_download(const b200pt_accel* a, b200pt_bvh_node* nodes_out, int64_t* n_nodes_out, uint32_t* ordered_out);
void b200pt_accel_destroy(b200pt_accel* a);
/* Primitive::world_bound (mod.rs:159-165): 6 floats. */
int b200pt_accel_world_bound(const b200pt_accel* a, float* bounds6);
/* Primitive::intersect / intersect_p for one ray (correctness path; the ray's
 * tmax is lowered on a hit exactly as the trait requires). */
int b200pt_accel_intersect1(const b200pt_accel* a, b200pt_ray* ray, b200pt_hit* hit);
int b200pt_accel_occluded1(const b200pt_accel* a, const b200pt_ray* ray, uint8_t* occluded);
/* Batched forms with HOST buffers (copies in and out inside the call). */
int b200pt_intersect_batch(const b200pt_accel* a, const b200pt_ray* rays, int64_t n, b200pt_hit* hits);
int b200pt_occluded_batch(const b200pt_accel* a, const b200pt_ray* rays, int64_t n, uint8_t* occluded);
/* Batched forms with DEVICE buffers, enqueued on `stream` (a cudaStream_t; 0 =
 * default stream), asynchronous. `variant` selects the traversal kernel
 * (0 = default; others are kept for A/B measurement). */
int b200pt_intersect_batch_device(const b200pt_accel* a, const void* d_rays, int64_t n, void* d_hits, void* stream, int variant);
int b200pt_occluded_batch_device(const b200pt_accel* a, const void* d_rays, int64_t n, void* d_occluded, void* stream, int variant);
/* Work accounting for the roofline: totals[0] = nodes whose bounds are tested,
 * totals[1] = triangle tests, summed over the batch, obtained by walking the 32-byte LinearBVHNode
 * array in the reference order (BVHAccel::intersect when any_hit == 0, intersect_p otherwise).
 * d_per_ray (optional, device): 2 x uint32 per ray. Synchronous. */
int b200pt_count_work_device(const b200pt_accel* a, const void* d_rays, int64_t n, int any_hit, uint64_t totals[2], void* d_per_ray);
/* Kernel launches issued by this library since init (bench.py's gpu_launches). */
int64_t b200pt_launch_count(void);

/* ---- scene + SamplerIntegrator (path / whitted / directlighting) -------- */
int b200pt_scene_create(const b200pt_scene_desc* desc, b200pt_scene** out);
void b200pt_scene_destroy(b200pt_scene* s);
/* Integrator::render (core/src/integrator/sampler_integrator.rs:243-304) for
 * the pixel rows [row_begin, row_end) of the cropped window (the multi-GPU
 * shard; pass 0, height for the whole image).  film_xyzw: 4 floats per pixel
 * of the FULL cropped window {X, Y, Z, filter_weight_sum}, zero outside the
 * shard, HOST memory. */
int b200pt_render_rows(b200pt_scene* s, int32_t row_begin, int32_t row_end, float* film_xyzw);
/* Same, film stays on the device (d_film_xyzw: device pointer, full window). */
int b200pt_render_rows_device(b200pt_scene* s, int32_t row_begin, int32_t row_end, void* d_film_xyzw, void* stream);
/* Multi-GPU decomposition: the pixel rows are cut into bands of `band_rows` rows dealt in snake order (b200pt_band_owner) to `n_shards`
 * shards (b200pt_band_owner; interleaving balances sky and geometry); this call renders shard `shard` into a zero-initialised film of
 * the full window (device memory).  Summing the shards' films (NCCL all-reduce) gives the whole image. */
/* Which shard renders band `band` (bands of band_rows pixel rows, counted from the top): dealt in snake order
 * 0 1 .. n-1 n-1 .. 1 0 0 1 .., so that a cost gradient down the image (sky above, geometry below) cancels within every
 * pair of passes instead of making the last shard of every pass the slowest. */
int32_t b200pt_band_owner(int32_t band, int32_t n_shards);
int b200pt_render_shard_device(b200pt_scene* s, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_xyzw, void* stream);
/* The same shard as {sum of filter-weighted RGB, sum of filter weights} per pixel - FilmTile's running sums before
 * Film::merge_film_tile converts them to XYZ (core/src/film/mod.rs:243-248).  Shard films are combined in this space
 * (copy the owned bands, add the overlapping rows), then b200pt_film_finish_device converts once, so that the assembled
 * film equals the single-device film bit for bit (the conversion is not additive in floating point).  In place allowed. */
int b200pt_render_shard_device_raw(b200pt_scene* s, int32_t shard, int32_t n_shards, int32_t band_rows, void* d_film_rgbw, void* stream);
int b200pt_film_finish_device(const void* d_film_rgbw, int64_t n_pix, void* d_film_xyzw, void* stream);
/* ---- several GPUs, one process ---------------------------
 * The reference renders from ONE process (bin/src/main.rs:29-85) and deals 16x16 tiles to its thread pool
 * (core/src/integrator/sampler_integrator.rs:252-296).  b200pt_multi_create replicates the scene on every listed device
 * (b200pt_init is called for each); b200pt_multi_render deals bands of `band_rows` pixel rows in snake order to the
 * devices, renders them concurrently (one host thread per device) and gathers the bands on devices[0] over NVLink with
 * NCCL: box-sized filters (radius <= 0.5 px) send each device's own bands into place (ncclSend / ncclRecv, 1 / n of the
 * film per device, no reduction); wider filters overlap by their apron and are summed with one ncclReduce.  Without
 * libnccl.so.2 (or with B200PT_GATHER=peer) the same transfers run as cudaMemcpyPeerAsync.  film_xyzw (HOST, 4 floats
 * per pixel of the cropped window {X, Y, Z, weight}, may be NULL) receives the assembled film; results are identical
 * to b200pt_render_rows on one device for box-sized filters (each sample is taken once, by exactly one device). */
typedef struct b200pt_multi b200pt_multi;
int b200pt_multi_create(const b200pt_scene_desc* desc, const int32_t* devices, int32_t n_devices, b200pt_multi** out);
int b200pt_multi_render(b200pt_multi* m, int32_t band_rows, float* film_xyzw);
/* rays of the last render summed over the devices, the slowest device's gather time, whether NCCL carried it */
int b200pt_multi_info(const b200pt_multi* m, uint64_t rays[3], double* gather_ms, int32_t* uses_nccl);
void* b200pt_multi_film_device(const b200pt_multi* m); /* devices[0]'s copy of the assembled film (device pointer) */
void b200pt_multi_destroy(b200pt_multi* m);
/* create + render + destroy */
int b200pt_render_multi(const b200pt_scene_desc* desc, const int32_t* devices, int32_t n_devices, int32_t band_rows, float* film_xyzw);

/* Film::write_image normalisation (film/mod.rs:356-417): XYZ+weight -> RGB,
 * 3 floats per pixel, HOST memory both sides. */
int b200pt_film_resolve(const b200pt_film* film, const float* film_xyzw, float* rgb_out);
/* Integrator::li for explicit (pixel x, pixel y, sample index) triples:
 * out = 3 floats (RGB radiance) per entry; rays_out (optional) = camera ray. */
int b200pt_li_batch(b200pt_scene* s, const int32_t* pixel_sample, int64_t n, float* li_out, b200pt_ray* rays_out);
/* Bytes of device memory the wave state of a render may take (ray queues, path state, the wave's samples: about 330
 * bytes per path in flight).  The image is rendered in as many waves as that needs; the film keeps running sums
 * between waves, so neither memory nor the result depends on resolution x samples per pixel or on the budget (the
 * reference holds one FilmTile per thread, core/src/film/film_tile.rs:62-108).  0 restores the default: the
 * environment variable B200PT_MEM_BUDGET (bytes, K / M / G suffix), else 60 % of the device's free memory. */
int b200pt_scene_set_memory_budget(b200pt_scene* s, uint64_t bytes);
/* Rays traced by the last render on this scene: [camera, closest-hit, shadow]. */
int b200pt_scene_ray_counts(const b200pt_scene* s, uint64_t counts[3]);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
