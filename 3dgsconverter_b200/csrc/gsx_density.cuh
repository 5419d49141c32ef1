#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t density_workspace_bytes(int64_t n, int64_t cap);
int density_voxel_count(const float* xyz, int64_t n, float voxel, int64_t min_points, int64_t* dense_vox_host,
                        int32_t* dense_cnt_host, int64_t cap, int64_t* n_dense_host, int64_t* n_voxels_host, void* ws,
                        int64_t ws_bytes, cudaStream_t st);
int density_member_mask(const float* xyz, int64_t n, float voxel, const int64_t* keep_vox_host, int64_t n_keep,
                        uint8_t* mask, void* ws, int64_t ws_bytes, cudaStream_t st);
}
