// Host-side preparation of an image-mapped InfiniteAreaLight (lights/src/infinite.rs:61-92, 326-369):
// MIP pyramid over the environment image (core/src/mipmap/mod.rs), the (2w x 2h) importance image and its
// Distribution2D tables.  The device only ever looks up level 0 (lookup_triangle(st, 0.0)) and the tables.
#pragma once
#include <cstdint>
#include <vector>

namespace b2host {

struct EnvMapTables {
    int width = 1, height = 1;       // level 0 of the pyramid (power-of-two sides)
    std::vector<float> texels;       // 4 floats per level-0 texel: r g b 0, row-major
    int nu = 2, nv = 2;              // importance image: 2 * width x 2 * height
    std::vector<float> cond_func;    // nv x nu
    std::vector<float> cond_cdf;     // nv x (nu + 1)
    std::vector<float> cond_int;     // nv
    std::vector<float> marg_func;    // nv
    std::vector<float> marg_cdf;     // nv + 1
    float marg_int = 0.0f;
    float power_lookup[3] = {0, 0, 0};  // l_map.lookup_triangle((0.5, 0.5), 0.5), infinite.rs:177-186
};

// rgb: map_width x map_height x 3 floats (NULL => the 1x1 image [L]); L multiplies every texel.
// importance = false (a goniometric light's map): the pyramid, level 0 and power_lookup only.
void build_envmap(const float* rgb, int map_width, int map_height, const float L[3], EnvMapTables* out, bool importance = true);

}  // namespace b2host
