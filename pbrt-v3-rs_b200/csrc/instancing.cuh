// Two-level (instanced) accelerator: device structs and launchers.  Internal header.
#pragma once
#include "traverse.cuh"

struct b200pt_scene_desc;

namespace b2 {

struct DInstance {
    float w2i[12];  // rows 0..2 of world_to_instance (affine)
    float i2w[12];  // rows 0..2 of instance_to_world
    int object;
    int identity;   // Transform::is_identity(instance_to_world)
    int pad[2];
};
struct DObject {
    float root_bounds[6];
    int root_code;
    int pad;
};
struct DeviceAccel2 {
    DeviceAccel top;  // wide nodes / records of the scene aggregate AND of every object (global indices)
    const DObject* objects;
    const DInstance* instances;
};
struct Accel2Impl {
    DeviceAccel2 dev;
    float4* d_wide = nullptr;
    float4* d_recs = nullptr;
    DObject* d_objs = nullptr;
    DInstance* d_insts = nullptr;
};

int accel2_build_device(const b200pt_scene_desc* d, Accel2Impl* out);
void accel2_free_device(Accel2Impl* a);
int launch_intersect2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_hits, cudaStream_t s, float* d_b2, int* d_inst);
int launch_occluded2(const DeviceAccel2& A, const void* d_rays, int64_t n, void* d_out, cudaStream_t s);

}  // namespace b2
