// Two-level (instanced) accelerator: device structs and launchers.  Internal header.
#pragma once
#include "common.cuh"

struct b200pt_scene_desc;

namespace b2 {

// Shading-side copy of a TransformedPrimitive's static transform (k_shade: transform_surface_interaction).
struct DInstance {
    float w2i[16];  // world_to_instance, row-major (full 4x4: the reference's Gauss-Jordan inverse is not exactly affine)
    float i2w[16];  // instance_to_world
    int object;
    int identity;   // Transform::is_identity(instance_to_world)
    int pad[2];
};
struct DeviceAccel2 {
    DeviceAccel top;  // wide nodes / records of the scene aggregate AND of every object (global indices)
    // Traversal-side instance record, 6 x float4 (96 B): rows 0..3 of world_to_instance, then the object's root
    // bounds {min.xyz, max.x} {max.yz, bits(root_code), bits(object)}.
    const float4* inst_trav;
    const DInstance* instances;
};
struct Accel2Impl {
    DeviceAccel2 dev;
    float4* d_wide = nullptr;
    float4* d_recs = nullptr;
    float4* d_trav = nullptr;
    DInstance* d_insts = nullptr;
};

int accel2_build_device(const b200pt_scene_desc* d, Accel2Impl* out);
void accel2_free_device(Accel2Impl* a);
// variant 0 = loop-free postponed-leaf persistent kernel with the instance entered / left inside one loop (default),
// variant 4 = phase-scheduled persistent kernel without postponement, variant 2 = one thread per ray with a nested walk.
int launch_intersect2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, float* d_b2, int* d_inst, int variant = 0,
                      const TraceLaunch* tl = nullptr);
int launch_occluded2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s, int variant = 0, const TraceLaunch* tl = nullptr);

}  // namespace b2
