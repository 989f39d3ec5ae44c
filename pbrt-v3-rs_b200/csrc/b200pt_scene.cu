// Scene + PathIntegrator entry points (wavefront path tracer) — under construction.
#include "common.cuh"
extern "C" {
int b200pt_scene_create(const b200pt_scene_desc*, b200pt_scene** out) { if (out) *out = nullptr; b200pt_set_error("scene: not implemented yet"); return B200PT_ERR_UNSUPPORTED; }
void b200pt_scene_destroy(b200pt_scene*) {}
int b200pt_render_rows(b200pt_scene*, int32_t, int32_t, float*) { return B200PT_ERR_UNSUPPORTED; }
int b200pt_render_rows_device(b200pt_scene*, int32_t, int32_t, void*, void*) { return B200PT_ERR_UNSUPPORTED; }
int b200pt_film_resolve(const b200pt_film*, const float*, float*) { return B200PT_ERR_UNSUPPORTED; }
int b200pt_li_batch(b200pt_scene*, const int32_t*, int64_t, float*, b200pt_ray*) { return B200PT_ERR_UNSUPPORTED; }
int b200pt_scene_ray_counts(const b200pt_scene*, uint64_t*) { return B200PT_ERR_UNSUPPORTED; }
}
