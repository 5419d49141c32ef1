#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t density_workspace_bytes(int64_t n, int64_t cap);
int density_voxel_count(const float* xyz, int64_t n, float voxel, int64_t min_points, int64_t* dense_vox_host,
                        int32_t* dense_cnt_host, int64_t cap, int64_t* n_dense_host, int64_t* n_voxels_host, void* ws,
                        int64_t ws_bytes, cudaStream_t st);
int density_member_mask(const float* xyz, int64_t n, float voxel, const int64_t* keep_vox_host, int64_t n_keep,
                        uint8_t* mask, void* ws, int64_t ws_bytes, cudaStream_t st);
void density_voxel_range(const float* minmax_host, float voxel, int64_t* q0, int64_t* dim);
int density_grid_count(const float* xyz, int64_t n, float voxel, const int64_t* q0, const int64_t* dim, int* grid_dev,
                       unsigned long long* oob_dev, cudaStream_t st);
int density_grid_dense(const int* grid_dev, const int64_t* q0, const int64_t* dim, int64_t min_points,
                       int64_t* dense_vox_host, int32_t* dense_cnt_host, int64_t cap, int64_t* n_dense_host,
                       int64_t* n_voxels_host, void* ws, int64_t ws_bytes, cudaStream_t st);
}
