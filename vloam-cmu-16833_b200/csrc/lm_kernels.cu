// placeholder, replaced below
#include "../../include/vloam_b200.h"
#include "internal.h"
namespace vb {
struct LMDevice { int B; };
cudaError_t lm_create(Profiler*, cudaStream_t, int B, int, const vloam_lidar_params*, LMDevice** out) { *out = new LMDevice{B}; return cudaSuccess; }
void lm_destroy(LMDevice* lm) { delete lm; }
void lm_reset(LMDevice*) {}
cudaError_t lm_run(LMDevice*, cudaStream_t, const SRHeader*, const float4*, const float4*, const LOState*, bool) { return cudaErrorNotSupported; }
cudaError_t lm_get_pose(LMDevice*, cudaStream_t, double*) { return cudaErrorNotSupported; }
cudaError_t lm_get_cloud(LMDevice*, cudaStream_t, int, int, float*, int, int*) { return cudaErrorNotSupported; }
cudaError_t lm_set_cube(LMDevice*, cudaStream_t, int, int, int, const float*, int) { return cudaErrorNotSupported; }
cudaError_t lm_get_cube(LMDevice*, cudaStream_t, int, int, int, float*, int, int*) { return cudaErrorNotSupported; }
cudaError_t lm_get_info(LMDevice*, cudaStream_t, int*) { return cudaErrorNotSupported; }
cudaError_t lm_get_trace(LMDevice*, cudaStream_t, int, int, double*, int*, double*) { return cudaErrorNotSupported; }
}
