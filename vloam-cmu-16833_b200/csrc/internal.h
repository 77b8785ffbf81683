// vloam_b200 — declarations shared between the C-ABI layer (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

#include <vector>

struct vloam_lidar_params;

namespace vb {

// Kernel ids for the launch counter / optional CUDA-event timing (vloam_ctx_enable_timing).
enum KernelId {
  K_SR_FIND_ENDS = 0, K_SR_CLASSIFY, K_SR_SCAN, K_SR_SCATTER, K_SR_CURVATURE, K_SR_PICK, K_SR_VOXEL, K_SR_PACK,
  K_LO_SET_MOTION, K_LO_ASSOCIATE, K_LO_SOLVE, K_LO_EXPORT, K_LO_INIT, K_LO_BUILD_GRID, K_LO_ASSOCIATE_BRUTE,
  K_LM_PREPARE, K_LM_VOXEL, K_LM_GRID, K_LM_ASSOCIATE, K_LM_FIT, K_LM_SOLVE, K_LM_INSERT, K_LM_REFILTER, K_LM_PLACE, K_LM_MISC,
  K_LO_ACCUMULATE, K_LO_STEP, K_LM_ACCUMULATE, K_LM_STEP,
  K_VO_PROJECT, K_VO_BUCKET, K_VO_QUERY, K_VO_SOLVE, K_VO_MISC, K_VO_MATCH, K_VO_DETECT, K_VO_DESCRIBE,
  K_COUNT
};
const char* kernel_name(int id);

// Counts every launch; when `enabled`, brackets each launch with CUDA events on the launching stream.
struct Profiler {
  bool enabled = false;
  long long launches = 0;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[K_COUNT] = {0};
  long long cnt[K_COUNT] = {0};
  cudaEvent_t cur_a = nullptr;
  int cur_id = -1;
  cudaEvent_t get();
  void begin(int id, cudaStream_t st);
  void end(cudaStream_t st);
  void collect();  // after a stream synchronize: fold the recorded events into ms[] / cnt[]
  void clear();
  ~Profiler();
};
#define VB_LAUNCH(prof, id, st, ...)      \
  do {                                    \
    (prof)->begin((id), (st));            \
    __VA_ARGS__;                          \
    (prof)->end((st));                    \
  } while (0)

// sr_kernels.cu
void launch_scan_registration(Profiler* prof, cudaStream_t st, int B, int cap, const float* xyz, int stride, size_t slab_floats,
                              const int* n_points_dev, float min_range, int n_scans, SRHeader* hdr, uint8_t* ring8,
                              int* blockHist, float4* cloud, float* curv, uint8_t* gapflag, int8_t* label, int* featIdx,
                              float4* lessFlatStage, float4* sharp, int* sharpIdx, float4* lessSharp, int* lessSharpIdx,
                              float4* flat, int* flatIdx, float4* lessFlat);

cudaError_t sr_prepare_device(int device);   // per-device function attributes, once per context

// lo_kernels.cu
void launch_lo_init(Profiler* prof, cudaStream_t st, LOState* lo, int B);
// Grid index over the (corner, surf) target clouds of every stream: [B][2] headers, cell tables and sorted copies.
struct GNState;
struct LOGrid {
  // wide solve (gn_split.cuh): residual records [B][kMaxSharp + kMaxFlat], per-stream state, tile partials, rank-summed counts
  double* gnRecV = nullptr;    // [B][7][kMaxSharp + kMaxFlat]
  float4* gnRecP = nullptr;    // [B][kMaxSharp + kMaxFlat]
  GNState* gnState = nullptr;
  double* gnPartial = nullptr;
  double* gnCounts = nullptr;
  void* ncclComm = nullptr;    // point-sharded streams: partial normal equations all-reduced across ranks
  GridHeader* hdr = nullptr;   // [B][2]
  int* cellStart = nullptr;    // [B][2][kGridCap + 1]
  int* cursor = nullptr;       // [B][2][kGridCap + 1] scatter cursors (scratch)
  float4* sorted[2] = {nullptr, nullptr};  // corner: [B][kMaxLessSharp], surf: [B][cap]
};
void launch_lo_pass(Profiler* prof, cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const SRHeader* hdrLast, LOState* lo,
                    const float4* sharp, const float4* flat, const float4* cornerLast, const float4* surfLast,
                    const LOGrid* grid, int4* corr, int pass, int max_iterations, int integrate, const double* prior,
                    const ShardView* shard = nullptr, int solverMode = 0, bool distortion = false);
// kd-tree rebuild of laser_odometry.cpp:525-526: index the current scan's less-sharp / less-flat clouds.
void launch_lo_build_grid(Profiler* prof, cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const float4* lessSharp,
                          const float4* lessFlat, const LOGrid* grid);
void launch_lo_export(Profiler* prof, cudaStream_t st, const LOState* lo, double* pose, int B);
void launch_lo_set_motion(Profiler* prof, cudaStream_t st, LOState* lo, const double* motion, int B);
void launch_lo_set_pose(Profiler* prof, cudaStream_t st, LOState* lo, const double* pose, int B);
cudaError_t lo_prepare_device(int device);   // per-device function attributes (opt-in shared memory), once per context

// lm_kernels.cu — laser mapping state of a batch of streams
struct LMDevice;
cudaError_t lm_create(Profiler* prof, cudaStream_t st, int B, int cap, const vloam_lidar_params* p, LMDevice** out);
void lm_destroy(LMDevice* lm);
void lm_reset(LMDevice* lm);
cudaError_t lm_ensure_alloc(LMDevice* lm, cudaStream_t st);
void lm_note_run(LMDevice* lm, bool skip_frame);
bool lm_graph_safe(const LMDevice* lm);
cudaError_t lm_run(LMDevice* lm, cudaStream_t st, const SRHeader* hdrCur, const float4* cornerLast, const float4* surfLast,
                   const LOState* lo, bool skip_frame);
cudaError_t lm_get_pose(LMDevice* lm, cudaStream_t st, double* pose_out);
cudaError_t lm_get_cloud(LMDevice* lm, cudaStream_t st, int stream, int which, float* out, int capacity, int* n_out);
cudaError_t lm_set_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, const float* xyzi, int n);
// vo_detect.cu: Shi-Tomasi key-point detection (image_util.cpp:11-37)
struct VODetect;
cudaError_t vo_detect_run(VODetect** d, Profiler* prof, cudaStream_t st, int B, const uint8_t* images, int H, int W, int maxCorners,
                          double quality, double minDistance, int* status_out);
cudaError_t vo_detect_status(VODetect* d, cudaStream_t st, int* status_out);
cudaError_t vo_detect_read(VODetect* d, cudaStream_t st, float* corners, int* n);
cudaError_t vo_detect_response(VODetect* d, cudaStream_t st, int stream, float* out, size_t pixels);
const float* vo_detect_corners_device(const VODetect* d);
const int* vo_detect_counts_device(const VODetect* d);
int vo_detect_height(const VODetect* d);
int vo_detect_width(const VODetect* d);
const uint8_t* vo_detect_image_device(const VODetect* d);
int vo_detect_max_corners(const VODetect* d);
// vo_orb.cu: ORB description of given key points (image_util.cpp:162-212)
void launch_vo_orb_describe(Profiler* prof, cudaStream_t st, int B, const uint8_t* img, int H, int W, const float* kp, const int* nKp,
                            int kpStride, int maxK, float* keptXY, int* keptIdx, uint8_t* desc, int* nKept);
void vo_detect_destroy(VODetect* d);
cudaError_t lm_get_registered(LMDevice* lm, cudaStream_t st, int stream, const float4* cloud, int n, float* out, int capacity, int* n_out);
cudaError_t lm_get_map_cloud(LMDevice* lm, cudaStream_t st, int stream, float* out, int capacity, int* n_out);
cudaError_t lm_get_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, float* out, int capacity, int* n_out);
cudaError_t lm_get_info(LMDevice* lm, cudaStream_t st, int* info);
cudaError_t lm_get_map_stats(LMDevice* lm, cudaStream_t st, int* stats);
cudaError_t lm_get_trace(LMDevice* lm, cudaStream_t st, int stream, int pass, double* records, int* info, double* para);
cudaError_t lm_get_status(LMDevice* lm, cudaStream_t st, int* status);
cudaError_t lm_get_counters(LMDevice* lm, cudaStream_t st, long long* counters);
void lm_set_debug_stats(LMDevice* lm, bool on);
void lm_set_nccl(LMDevice* lm, void* comm, int rank, int world);
cudaError_t lm_get_queries(LMDevice* lm, cudaStream_t st, int stream, int pass, int kind, int* out, int capacity, int* n_out);

}  // namespace vb

#include <string>
// The context behind the opaque vloam_ctx handle (shared by capi.cu and vo_kernels.cu).
struct vloam_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host->device uploads overlap the previous scan's kernels
  bool batch_copy_unsupported = false;  // cudaMemcpyBatchAsync refused once: keep to per-stream copies
  cudaStream_t copy_stream2 = nullptr; // second upload queue: per-stream copies alternate between the two, so the DMA set-up of one
                                       // copy hides behind the transfer of another (a batch is up to a few hundred 1.5 MB copies)
  cudaEvent_t ev_copy2 = nullptr;
  std::string last_error;
  vb::Profiler prof;
};
