// placeholder, replaced below
#include "../../include/vloam_b200.h"
extern "C" {
int vloam_vo_create(vloam_ctx*, int, int, int, vloam_vo**) { return VLOAM_E_STATE; }
int vloam_vo_destroy(vloam_vo*) { return VLOAM_E_STATE; }
int vloam_vo_set_calibration(vloam_vo*, const float*, const float*, const float*) { return VLOAM_E_STATE; }
int vloam_vo_reset(vloam_vo*) { return VLOAM_E_STATE; }
int vloam_vo_process_cloud(vloam_vo*, const float*, const int*, int, size_t) { return VLOAM_E_STATE; }
int vloam_vo_process_cloud_device(vloam_vo*, const float*, const int*, int, size_t) { return VLOAM_E_STATE; }
int vloam_vo_query_depth(vloam_vo*, int, int, const float*, int, float*) { return VLOAM_E_STATE; }
int vloam_vo_get_buckets(vloam_vo*, int, int, float*, float*, float*, int*) { return VLOAM_E_STATE; }
int vloam_vo_solve(vloam_vo*, const float*, const float*, const int*, const double*, int, int, double*) { return VLOAM_E_STATE; }
}
