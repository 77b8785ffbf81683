// vloam_b200 — declarations shared between the C-ABI layer (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

struct vloam_lidar_params;

namespace vb {

// sr_kernels.cu
void launch_scan_registration(cudaStream_t st, int B, int cap, const float* xyz, int stride, size_t slab_floats,
                              const int* n_points_dev, float min_range, int n_scans, SRHeader* hdr, uint8_t* ring8,
                              int* blockHist, float4* cloud, float* curv, int8_t* label, int* featIdx,
                              float4* lessFlatStage, float4* sharp, int* sharpIdx, float4* lessSharp, int* lessSharpIdx,
                              float4* flat, int* flatIdx, float4* lessFlat);

// lo_kernels.cu
void launch_lo_init(cudaStream_t st, LOState* lo, int B);
void launch_lo_pass(cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const SRHeader* hdrLast, LOState* lo,
                    const float4* sharp, const float4* flat, const float4* cornerLast, const float4* surfLast,
                    int4* corr, int pass, int max_iterations, int integrate, const double* prior);
void launch_lo_export(cudaStream_t st, const LOState* lo, double* pose, int B);
void launch_lo_set_motion(cudaStream_t st, LOState* lo, const double* motion, int B);

// lm_kernels.cu — laser mapping state of a batch of streams
struct LMDevice;
cudaError_t lm_create(cudaStream_t st, int B, int cap, const vloam_lidar_params* p, LMDevice** out);
void lm_destroy(LMDevice* lm);
void lm_reset(LMDevice* lm);
cudaError_t lm_run(LMDevice* lm, cudaStream_t st, const SRHeader* hdrCur, const float4* cornerLast, const float4* surfLast,
                   const LOState* lo, bool skip_frame, long long* launches);
cudaError_t lm_get_pose(LMDevice* lm, cudaStream_t st, double* pose_out);
cudaError_t lm_get_cloud(LMDevice* lm, cudaStream_t st, int stream, int which, float* out, int capacity, int* n_out);
cudaError_t lm_set_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, const float* xyzi, int n);
cudaError_t lm_get_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, float* out, int capacity, int* n_out);
cudaError_t lm_get_info(LMDevice* lm, cudaStream_t st, int* info);
cudaError_t lm_get_trace(LMDevice* lm, cudaStream_t st, int stream, int pass, double* records, int* info, double* para);

}  // namespace vb
