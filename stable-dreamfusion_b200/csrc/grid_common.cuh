// grid_common.cuh — per-level parameters of the multiresolution grid, shared by gridenc.cu and fused_field.cu.
#pragma once
#include "common.cuh"

constexpr uint32_t kMaxLevels = 32;

struct LevelParams {
    uint32_t offset[kMaxLevels];      // first entry of the level
    uint32_t size[kMaxLevels];        // entries in the level (hashmap_size)
    uint32_t res[kMaxLevels];         // ceil(exp2f(level*S)*H)                  reference gridencoder.cu:133
};

// Fills the per-device scratch LevelParams from the offsets tensor on `st` (device-side exp2f, as the
// reference evaluates it per thread) and returns its device address.
int sdf_get_level_params(const int* offsets, uint32_t L, float S, uint32_t H, cudaStream_t st, LevelParams** out);
