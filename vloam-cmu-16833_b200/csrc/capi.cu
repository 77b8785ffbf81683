// vloam_b200 — C-ABI implementation (include/vloam_b200.h): handles, device buffers, launch sequencing.
// No torch types, no CPU fallback: every entry point either runs the CUDA path or returns an error code.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/vloam_b200.h"
#include "common.cuh"
#include "gn_split.cuh"
#include "internal.h"

using namespace vb;



namespace vb {
const char* kernel_name(int id) {
  static const char* names[K_COUNT] = {
      "sr_find_ends", "sr_classify", "sr_scan", "sr_scatter", "sr_curvature", "sr_pick_features", "sr_less_flat_voxel", "sr_pack",
      "lo_set_motion", "lo_associate", "lo_solve", "lo_export_pose", "lo_init_state", "lo_build_grid", "lo_associate_brute",
      "lm_prepare", "lm_voxel", "lm_index", "lm_associate", "lm_fit", "lm_solve", "lm_insert", "lm_refilter", "lm_place", "lm_misc",
      "lo_accumulate", "lo_step", "lm_accumulate", "lm_step",
      "vo_project", "vo_bucket", "vo_query", "vo_solve", "vo_misc", "vo_bf_match", "vo_detect", "vo_orb_describe"};
  return (id >= 0 && id < K_COUNT) ? names[id] : "?";
}
cudaEvent_t Profiler::get() {
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void Profiler::begin(int id, cudaStream_t st) {
  ++launches;
  if (!enabled) return;
  cur_a = get(); cur_id = id;
  cudaEventRecord(cur_a, st);
}
void Profiler::end(cudaStream_t st) {
  if (!enabled || cur_id < 0) return;
  cudaEvent_t b = get();
  cudaEventRecord(b, st);
  recs.push_back({cur_id, cur_a, b});
  cur_id = -1;
}
void Profiler::collect() {
  for (const Rec& r : recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.id] += t; cnt[r.id]++; }
    pool.push_back(r.a); pool.push_back(r.b);
  }
  recs.clear();
}
void Profiler::clear() { for (int i = 0; i < K_COUNT; ++i) { ms[i] = 0; cnt[i] = 0; } }
Profiler::~Profiler() {
  collect();
  for (cudaEvent_t e : pool) cudaEventDestroy(e);
}
}  // namespace vb

// ---------------------------------------------------------------------------------------------------------------
// NCCL binding (BASELINE configs[4]: "NCCL allreduce of 6x6 J'J per GN iteration").  The library does not link NCCL: the
// four entry points it needs are resolved at run time from the libnccl the process already carries (torch's) or from
// VLOAM_NCCL_LIB.  Only the point-sharded mode uses it; everything else runs without NCCL present.
struct NcclId { char internal[128]; };      // ncclUniqueId
namespace {
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // 1. the libnccl the process already carries (a framework loaded earlier), 2. VLOAM_NCCL_LIB, 3. the system's.
    // A process that loads a framework with its own NCCL *later* must point VLOAM_NCCL_LIB at that copy: the dynamic loader
    // resolves the framework's dependency by soname to whichever libnccl.so.2 came first (the Python mirror does this).
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char* names[] = {getenv("VLOAM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (api.lib) break;
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    }
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclId, int)>(dlsym(api.lib, "ncclCommInitRank"));
      api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(api.lib, "ncclAllReduce"));
      api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.lib = nullptr;
    }
  }
  return api.lib ? &api : nullptr;
}
}  // namespace

namespace vb {
// The exchange step of the wide solve for point-sharded streams: sum the partial normal equations (or the correspondence
// counts) over the ranks, in place, on the solve's stream.  ncclFloat64 = 8, ncclSum = 0.
void gn_allreduce_partials(void* ncclComm, double* partial, size_t count, cudaStream_t st) {
  NcclApi* api = nccl_api();
  if (api && ncclComm) api->AllReduce(partial, partial, count, 8, 0, ncclComm, st);
}
}  // namespace vb

namespace {

int fail(vloam_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->last_error = what;
    if (e != cudaSuccess) { c->last_error += ": "; c->last_error += cudaGetErrorString(e); }
  }
  return code;
}

#define CU(ctx, call)                                                        \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return fail((ctx), VLOAM_E_CUDA, #call, e__);    \
  } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e == cudaSuccess && n) e = cudaMemset(*p, 0, n * sizeof(T));
  return e;
}

}  // namespace

struct vloam_lidar {
  vloam_ctx* ctx = nullptr;
  vloam_lidar_params p{};
  int B = 0, cap = 0, nblk = 0;
  long long frame = -1;      // index of the last registered scan; -1 = none
  bool lo_done_for_frame = false;
  long long lo_frames = 0;   // LaserOdometry::frameCount
  // input staging
  int last_stride = 3;
  // point-sharded solve (vloam_shard_*): this rank's exchange slots, the peers' (IPC-mapped or raw), counters, error flag
  ShardView shard;
  ShardSlot* d_xbuf = nullptr;
  void* ipc_open[kMaxShard] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void* nccl = nullptr;      // ncclComm_t of the point-sharded group (vloam_shard_nccl_init)
  float* d_in[2] = {nullptr, nullptr};  // [B][cap][in_stride], double-buffered so the next upload overlaps this scan's kernels
  int in_stride = 4;                    // floats per point the input slabs are sized for (grown on demand, <= kMaxInputStride)
  int* d_n[2] = {nullptr, nullptr};     // [B]
  cudaEvent_t ev_in_ready[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr};
  bool in_used[2] = {false, false};
  long long host_scans = 0;
  // scan registration
  SRHeader* d_hdr[2] = {nullptr, nullptr};
  uint8_t* d_ring8 = nullptr;
  int* d_blockHist = nullptr;
  float4* d_cloud[2] = {nullptr, nullptr};  // laserCloud of the current / previous scan (laserCloudFullRes)
  float* d_curv = nullptr;
  uint8_t* d_gapflag = nullptr;
  int8_t* d_label = nullptr;
  int* d_featIdx = nullptr;
  float4* d_lessFlatStage = nullptr;
  float4* d_sharp = nullptr; int* d_sharpIdx = nullptr;
  float4* d_lessSharp[2] = {nullptr, nullptr}; int* d_lessSharpIdx = nullptr;
  float4* d_flat = nullptr; int* d_flatIdx = nullptr;
  float4* d_lessFlat[2] = {nullptr, nullptr};
  // laser odometry
  LOState* d_lo = nullptr;
  int4* d_corr[2] = {nullptr, nullptr};  // per outer pass (kept for parity read-out)
  double* d_prior = nullptr;             // [B][7]
  double* d_pose = nullptr;              // [B][16]
  double* h_pose[2] = {nullptr, nullptr};  // pinned, one per frame parity
  cudaEvent_t ev_pose[2] = {nullptr, nullptr};
  bool pose_valid[2] = {false, false};
  LOGrid grid;
  // laser mapping
  LMDevice* lm = nullptr;
  // CUDA graphs of one frame's launch sequence (vloam_lidar_process): [buffer parity][odometry initialised][mapping skipped]
  cudaGraphExec_t graph[2][2][2] = {{{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}};
  bool capturing = false;
  long long graph_launches[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};   // kernel launches one replay stands for (bench.py's gpu_launches)
  const double* graph_prior[2][2][2] = {{{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}};
  int cur() const { return (int)(frame & 1); }
};

extern "C" {

// ------------------------------------------------------------------------------------------------ context
int vloam_ctx_create(int device, vloam_ctx** out) {
  if (!out) return VLOAM_E_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return VLOAM_E_CUDA;
  vloam_ctx* c = new (std::nothrow) vloam_ctx();
  if (!c) return VLOAM_E_NOMEM;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_copy2, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return VLOAM_E_CUDA;
  }
  c->stream = c->own_stream;
  // opt-in shared-memory sizes of the kernels, once per context on its device; a failure here would otherwise surface
  // later as an opaque launch error
  if (sr_prepare_device(device) != cudaSuccess || lo_prepare_device(device) != cudaSuccess) {
    cudaStreamDestroy(c->own_stream); cudaStreamDestroy(c->copy_stream); cudaStreamDestroy(c->copy_stream2);
    delete c;
    return VLOAM_E_CUDA;
  }
  *out = c;
  return VLOAM_OK;
}
int vloam_ctx_destroy(vloam_ctx* c) {
  if (!c) return VLOAM_E_INVALID;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->copy_stream2) cudaStreamDestroy(c->copy_stream2);
  if (c->ev_copy2) cudaEventDestroy(c->ev_copy2);
  delete c;
  return VLOAM_OK;
}
int vloam_ctx_set_stream(vloam_ctx* c, void* s) {
  if (!c) return VLOAM_E_INVALID;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return VLOAM_OK;
}
int vloam_ctx_synchronize(vloam_ctx* c) {
  if (!c) return VLOAM_E_INVALID;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}
const char* vloam_last_error(vloam_ctx* c) { return c ? c->last_error.c_str() : "null context"; }
long long vloam_ctx_launch_count(vloam_ctx* c) { return c ? c->prof.launches : 0; }
int vloam_ctx_enable_timing(vloam_ctx* c, int on) {
  if (!c) return VLOAM_E_INVALID;
  c->prof.enabled = on != 0;
  return VLOAM_OK;
}
int vloam_ctx_kernel_count(void) { return K_COUNT; }
const char* vloam_ctx_kernel_name(int id) { return kernel_name(id); }
int vloam_ctx_get_kernel_timings(vloam_ctx* c, double* ms, long long* counts, int reset) {
  if (!c) return VLOAM_E_INVALID;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  c->prof.collect();
  for (int i = 0; i < K_COUNT; ++i) { if (ms) ms[i] = c->prof.ms[i]; if (counts) counts[i] = c->prof.cnt[i]; }
  if (reset) c->prof.clear();
  return VLOAM_OK;
}

// ------------------------------------------------------------------------------------------------ lidar handle
int vloam_lidar_params_default(vloam_lidar_params* p) {
  if (!p) return VLOAM_E_INVALID;
  p->batch = 1;
  p->max_points = 131072;
  p->scan_line = 64;                   // loam_velodyne_HDL_64_kitti.launch:3
  p->minimum_range = 5.0;              // :13
  p->mapping_line_resolution = 0.4;    // :15
  p->mapping_plane_resolution = 0.8;   // :16
  p->mapping_skip_frame = 1;           // :6
  p->detach_VO_LO = 1;                 // vloam_main.launch:4
  p->lo_outer_passes = 2;
  p->lo_max_iterations = 4;
  p->lm_outer_passes = 2;
  p->lm_max_iterations = 4;
  p->map_capacity_points = 1 << 21;
  p->debug_keep_submap = 0;
  p->solver_mode = 0;
  p->distortion = 0;                   // laser_odometry.h:90
  return VLOAM_OK;
}

int vloam_lidar_destroy(vloam_lidar* h) {
  if (!h) return VLOAM_E_INVALID;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->d_in[i]); cudaFree(h->d_n[i]);
    if (h->ev_in_ready[i]) cudaEventDestroy(h->ev_in_ready[i]);
    if (h->ev_in_free[i]) cudaEventDestroy(h->ev_in_free[i]);
    if (h->ev_pose[i]) cudaEventDestroy(h->ev_pose[i]);
    if (h->h_pose[i]) cudaFreeHost(h->h_pose[i]);
  }
  for (int i = 0; i < 2; ++i) { cudaFree(h->d_hdr[i]); cudaFree(h->d_cloud[i]); cudaFree(h->d_lessSharp[i]); cudaFree(h->d_lessFlat[i]); cudaFree(h->d_corr[i]); }
  cudaFree(h->d_ring8); cudaFree(h->d_blockHist); cudaFree(h->d_curv); cudaFree(h->d_gapflag); cudaFree(h->d_label); cudaFree(h->d_featIdx);
  cudaFree(h->d_lessFlatStage); cudaFree(h->d_sharp); cudaFree(h->d_sharpIdx); cudaFree(h->d_lessSharpIdx);
  cudaFree(h->d_flat); cudaFree(h->d_flatIdx); cudaFree(h->d_lo); cudaFree(h->d_prior); cudaFree(h->d_pose);
  cudaFree(h->grid.hdr); cudaFree(h->grid.cellStart); cudaFree(h->grid.cursor);
  cudaFree(h->grid.gnRecV); cudaFree(h->grid.gnRecP); cudaFree(h->grid.gnState); cudaFree(h->grid.gnPartial); cudaFree(h->grid.gnCounts);
  for (int i = 0; i < 2; ++i) { cudaFree(h->grid.sorted[i]); }
  for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int k = 0; k < 2; ++k) if (h->graph[a][b][k]) cudaGraphExecDestroy(h->graph[a][b][k]);
  if (h->nccl) { if (NcclApi* api = nccl_api()) api->CommDestroy(h->nccl); h->nccl = nullptr; }
  if (h->lm) lm_destroy(h->lm);
  for (int r = 0; r < kMaxShard; ++r) if (h->ipc_open[r]) cudaIpcCloseMemHandle(h->ipc_open[r]);
  cudaFree(h->d_xbuf); cudaFree(h->shard.seq); cudaFree(h->shard.error);
  delete h;
  return VLOAM_OK;
}

// ------------------------------------------------------------------------------------------------ point-sharded solve
static int shard_alloc(vloam_lidar* h) {
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  if (h->d_xbuf) return VLOAM_OK;
  CU(c, cudaMalloc(&h->d_xbuf, (size_t)h->B * sizeof(ShardSlot)));
  CU(c, cudaMalloc(&h->shard.seq, (size_t)h->B * sizeof(unsigned long long)));
  CU(c, cudaMalloc(&h->shard.error, sizeof(int)));
  CU(c, cudaMemset(h->d_xbuf, 0, (size_t)h->B * sizeof(ShardSlot)));
  CU(c, cudaMemset(h->shard.seq, 0, (size_t)h->B * sizeof(unsigned long long)));
  CU(c, cudaMemset(h->shard.error, 0, sizeof(int)));
  return VLOAM_OK;
}

int vloam_shard_buffer(vloam_lidar* h, void** dev_ptr, size_t* bytes) {
  if (!h || !dev_ptr) return VLOAM_E_INVALID;
  const int r = shard_alloc(h);
  if (r) return r;
  *dev_ptr = h->d_xbuf;
  if (bytes) *bytes = (size_t)h->B * sizeof(ShardSlot);
  return VLOAM_OK;
}

int vloam_shard_ipc_handle(vloam_lidar* h, unsigned char* handle64) {
  if (!h || !handle64) return VLOAM_E_INVALID;
  const int r = shard_alloc(h);
  if (r) return r;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t hd;
  CU(h->ctx, cudaIpcGetMemHandle(&hd, h->d_xbuf));
  std::memcpy(handle64, &hd, 64);
  return VLOAM_OK;
}

int vloam_shard_enable(vloam_lidar* h, int rank, int world, void* const* peer_ptrs) {
  if (!h || !peer_ptrs || world < 1 || world > kMaxShard || rank < 0 || rank >= world) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  if (h->B > 128) return fail(c, VLOAM_E_CAPACITY, "point-sharded mode keeps one resident CTA per stream: batch <= 128");
  const int r = shard_alloc(h);
  if (r) return r;
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < world; ++i) {
    if (!peer_ptrs[i]) return VLOAM_E_INVALID;
    h->shard.peer[i] = static_cast<ShardSlot*>(peer_ptrs[i]);
  }
  h->shard.peer[rank] = h->d_xbuf;
  h->shard.rank = rank; h->shard.world = world;
  CU(c, cudaMemset(h->d_xbuf, 0, (size_t)h->B * sizeof(ShardSlot)));
  CU(c, cudaMemset(h->shard.seq, 0, (size_t)h->B * sizeof(unsigned long long)));
  CU(c, cudaMemset(h->shard.error, 0, sizeof(int)));
  CU(c, cudaDeviceSynchronize());
  return VLOAM_OK;
}

int vloam_shard_open_ipc(vloam_lidar* h, int rank, int world, const unsigned char* handles) {
  if (!h || !handles || world < 1 || world > kMaxShard || rank < 0 || rank >= world) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  const int r0 = shard_alloc(h);
  if (r0) return r0;
  void* ptrs[kMaxShard] = {nullptr};
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ptrs[r] = h->d_xbuf; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handles + (size_t)r * 64, 64);
    if (h->ipc_open[r]) { cudaIpcCloseMemHandle(h->ipc_open[r]); h->ipc_open[r] = nullptr; }
    CU(c, cudaIpcOpenMemHandle(&h->ipc_open[r], hd, cudaIpcMemLazyEnablePeerAccess));
    ptrs[r] = h->ipc_open[r];
  }
  return vloam_shard_enable(h, rank, world, ptrs);
}

int vloam_shard_disable(vloam_lidar* h) {
  if (!h) return VLOAM_E_INVALID;
  CU(h->ctx, cudaStreamSynchronize(h->ctx->stream));
  h->shard.rank = 0; h->shard.world = 1;
  return VLOAM_OK;
}

int vloam_shard_status(vloam_lidar* h, int* error_bits) {
  if (!h || !error_bits) return VLOAM_E_INVALID;
  *error_bits = 0;
  if (!h->shard.error) return VLOAM_OK;
  vloam_ctx* c = h->ctx;
  CU(c, cudaMemcpyAsync(error_bits, h->shard.error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

int vloam_shard_nccl_unique_id(unsigned char* id128) {
  if (!id128) return VLOAM_E_INVALID;
  NcclApi* api = nccl_api();
  if (!api) return VLOAM_E_STATE;
  NcclId id;
  if (api->GetUniqueId(&id) != 0) return VLOAM_E_CUDA;
  std::memcpy(id128, id.internal, 128);
  return VLOAM_OK;
}

int vloam_shard_nccl_init(vloam_lidar* h, int rank, int world, const unsigned char* id128) {
  if (!h || !id128 || world < 1 || world > kMaxShard || rank < 0 || rank >= world) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  NcclApi* api = nccl_api();
  if (!api) return fail(c, VLOAM_E_STATE, "vloam_shard_nccl_init: no libnccl in the process (set VLOAM_NCCL_LIB)");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  if (h->nccl) { api->CommDestroy(h->nccl); h->nccl = nullptr; }
  NcclId id;
  std::memcpy(id.internal, id128, 128);
  const int r = api->CommInitRank(&h->nccl, world, id, rank);
  if (r != 0) { h->nccl = nullptr; return fail(c, VLOAM_E_CUDA, api->GetErrorString ? api->GetErrorString(r) : "ncclCommInitRank failed"); }
  h->shard.rank = rank; h->shard.world = world;       // lo_associate / lm_knn deal the queries to the ranks
  h->grid.ncclComm = h->nccl;
  lm_set_nccl(h->lm, h->nccl, rank, world);
  return VLOAM_OK;
}

int vloam_shard_nccl_destroy(vloam_lidar* h) {
  if (!h) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  NcclApi* api = nccl_api();
  if (h->nccl && api) api->CommDestroy(h->nccl);
  h->nccl = nullptr; h->grid.ncclComm = nullptr;
  lm_set_nccl(h->lm, nullptr, 0, 1);
  h->shard.rank = 0; h->shard.world = 1;
  return VLOAM_OK;
}

int vloam_lidar_create(vloam_ctx* c, const vloam_lidar_params* p, vloam_lidar** out) {
  if (!c || !p || !out) return VLOAM_E_INVALID;
  *out = nullptr;
  if (p->batch < 1 || p->max_points < 64 || (p->scan_line != 16 && p->scan_line != 32 && p->scan_line != 64))
    return fail(c, VLOAM_E_INVALID, "vloam_lidar_create: batch >= 1, max_points >= 64, scan_line in {16,32,64}");
  if (p->mapping_skip_frame < 1 || p->lo_outer_passes < 0 || p->lo_max_iterations < 0 || p->lm_outer_passes < 0 || p->lm_max_iterations < 0 ||
      p->solver_mode < 0 || p->solver_mode > 2 || !(p->mapping_line_resolution > 0.0) || !(p->mapping_plane_resolution > 0.0) || !(p->minimum_range >= 0.0))
    return fail(c, VLOAM_E_INVALID, "vloam_lidar_create: mapping_skip_frame >= 1, pass / iteration counts >= 0, resolutions > 0, minimum_range >= 0");
  CU(c, cudaSetDevice(c->device));
  vloam_lidar* h = new (std::nothrow) vloam_lidar();
  if (!h) return VLOAM_E_NOMEM;
  h->ctx = c; h->p = *p; h->B = p->batch;
  h->cap = (p->max_points + kClassifyBlock - 1) / kClassifyBlock * kClassifyBlock;
  h->nblk = h->cap / kClassifyBlock;
  const size_t B = h->B, cap = h->cap;
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  for (int i = 0; i < 2; ++i) {
    A(dalloc(&h->d_in[i], B * cap * 4)); A(dalloc(&h->d_n[i], B));
    A(cudaEventCreateWithFlags(&h->ev_in_ready[i], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&h->ev_in_free[i], cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&h->ev_pose[i], cudaEventDisableTiming));
    A(cudaMallocHost((void**)&h->h_pose[i], B * 16 * sizeof(double)));
  }
  for (int i = 0; i < 2; ++i) {
    A(dalloc(&h->d_hdr[i], B)); A(dalloc(&h->d_cloud[i], B * cap));
    A(dalloc(&h->d_lessSharp[i], B * kMaxLessSharp)); A(dalloc(&h->d_lessFlat[i], B * cap));
    A(dalloc(&h->d_corr[i], B * (kMaxSharp + kMaxFlat)));
  }
  A(dalloc(&h->d_ring8, B * cap)); A(dalloc(&h->d_blockHist, B * h->nblk * kMaxRings));
  A(dalloc(&h->d_curv, B * cap)); A(dalloc(&h->d_gapflag, B * cap)); A(dalloc(&h->d_label, B * cap));
  A(dalloc(&h->d_featIdx, B * kMaxRings * kSectors * 26)); A(dalloc(&h->d_lessFlatStage, B * cap));
  A(dalloc(&h->d_sharp, B * kMaxSharp)); A(dalloc(&h->d_sharpIdx, B * kMaxSharp));
  A(dalloc(&h->d_lessSharpIdx, B * kMaxLessSharp));
  A(dalloc(&h->d_flat, B * kMaxFlat)); A(dalloc(&h->d_flatIdx, B * kMaxFlat));
  A(dalloc(&h->d_lo, B)); A(dalloc(&h->d_prior, B * 7)); A(dalloc(&h->d_pose, B * 16));
  A(dalloc(&h->grid.hdr, B * 2)); A(dalloc(&h->grid.cellStart, B * 2 * (kGridCap + 1))); A(dalloc(&h->grid.cursor, B * 2 * (kGridCap + 1)));
  A(dalloc(&h->grid.sorted[0], B * kMaxLessSharp));
  A(dalloc(&h->grid.sorted[1], B * cap));
  A(dalloc(&h->grid.gnRecV, B * 7 * (kMaxSharp + kMaxFlat))); A(dalloc(&h->grid.gnRecP, B * (kMaxSharp + kMaxFlat))); A(dalloc(&h->grid.gnState, B)); A(dalloc(&h->grid.gnPartial, B * kGnTiles * 28));
  A(dalloc(&h->grid.gnCounts, B * 2));
  if (e != cudaSuccess) { vloam_lidar_destroy(h); return fail(c, e == cudaErrorMemoryAllocation ? VLOAM_E_NOMEM : VLOAM_E_CUDA, "vloam_lidar_create: allocation", e); }
  launch_lo_init(&c->prof, c->stream, h->d_lo, h->B);
  e = lm_create(&c->prof, c->stream, h->B, h->cap, p, &h->lm);
  if (e != cudaSuccess) { vloam_lidar_destroy(h); return fail(c, VLOAM_E_CUDA, "vloam_lidar_create: map allocation", e); }
  e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) { vloam_lidar_destroy(h); return fail(c, VLOAM_E_CUDA, "vloam_lidar_create: initialisation", e); }
  *out = h;
  return VLOAM_OK;
}

int vloam_lidar_reset(vloam_lidar* h) {
  if (!h) return VLOAM_E_INVALID;
  // ScanRegistration::reset clears the per-scan clouds (scan_registration.cpp:89-98) — the device buffers are
  // overwritten by the next scan; LaserMapping::reset zeroes the valid / surround cube counts (laser_mapping.cpp:127-131).
  lm_reset(h->lm);
  return VLOAM_OK;
}

// ------------------------------------------------------------------------------------------------ scan registration
static int run_scan_registration(vloam_lidar* h, const float* xyz_dev, const int* n_dev, int stride, size_t slab_points) {
  vloam_ctx* c = h->ctx;
  (void)cudaGetLastError();   // a stale, unrelated error must not be blamed on the launches below
  h->frame++;
  h->lo_done_for_frame = false;
  const int cur = h->cur();
  launch_scan_registration(&c->prof, c->stream, h->B, h->cap, xyz_dev, stride, slab_points * (size_t)stride, n_dev,
                           (float)h->p.minimum_range, h->p.scan_line, h->d_hdr[cur], h->d_ring8, h->d_blockHist,
                           h->d_cloud[cur], h->d_curv, h->d_gapflag, h->d_label, h->d_featIdx, h->d_lessFlatStage, h->d_sharp,
                           h->d_sharpIdx, h->d_lessSharp[cur], h->d_lessSharpIdx, h->d_flat, h->d_flatIdx,
                           h->d_lessFlat[cur]);
  CU(c, cudaGetLastError());
  return VLOAM_OK;
}

// Records of up to 16 floats (a sensor_msgs/PointCloud2 point_step of 64 bytes) are taken as they are: x, y, z first.
constexpr int kMaxInputStride = 16;
}  // extern "C"

// Upload one scan per stream on the copy stream into the next input slot; the main stream waits for it.  src(b) = host
// address of stream b's points; contiguous != nullptr: the slabs are `slab_points` apart in one buffer (one DMA for the whole
// batch when they line up).
template <typename Src>
static int upload_only(vloam_lidar* h, Src src, const float* contiguous, const int* n_points, int stride, size_t slab_points) {
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  if (stride > h->in_stride) {   // first scan with wider records: re-size both input slabs (rare; synchronises)
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2; ++i) {
      cudaFree(h->d_in[i]); h->d_in[i] = nullptr;
      CU(c, dalloc(&h->d_in[i], (size_t)h->B * h->cap * stride));
      h->in_used[i] = false;
    }
    h->in_stride = stride;
  }
  for (int b = 0; b < h->B; ++b) {
    if (n_points[b] < 0 || (size_t)n_points[b] > slab_points) return fail(c, VLOAM_E_INVALID, "n_points[b] exceeds slab_points");
    if (n_points[b] > h->cap) return fail(c, VLOAM_E_CAPACITY, "scan larger than max_points");
  }
  // Upload on the copy stream into the input slot the previous-but-one scan used; the kernels of the previous scan
  // (other slot) keep running meanwhile.  ev_in_free[slot] = the kernels that last read this slot have finished.
  const int slot = (int)(h->host_scans & 1);
  if (h->in_used[slot]) CU(c, cudaStreamWaitEvent(c->copy_stream, h->ev_in_free[slot], 0));
  // one DMA for the whole batch when the slabs line up (same count everywhere): a copy per stream costs a few
  // microseconds of set-up each, which shows at PCIe rate (1.5 MB slabs take ~30 us)
  bool same = contiguous != nullptr;
  for (int b = 1; b < h->B; ++b) same = same && n_points[b] == n_points[0];
  if (same && n_points[0] > 0 && h->B > 1) {
    const size_t row = (size_t)n_points[0] * stride * sizeof(float);
    if (slab_points == (size_t)h->cap && (size_t)n_points[0] == slab_points)
      CU(c, cudaMemcpyAsync(h->d_in[slot], contiguous, row * h->B, cudaMemcpyHostToDevice, c->copy_stream));
    else
      CU(c, cudaMemcpy2DAsync(h->d_in[slot], (size_t)h->cap * stride * sizeof(float), contiguous, slab_points * stride * sizeof(float), row,
                              (size_t)h->B, cudaMemcpyHostToDevice, c->copy_stream));
  } else {
    // One buffer per stream.  All of them go to the driver as ONE batch (cudaMemcpyBatchAsync, CUDA 12.8+): measured on B200
    // with 192 pinned 1.5 MB buffers, 55.4 GB/s — the rate of a single contiguous copy — against 48.4 GB/s for one
    // cudaMemcpyAsync per buffer on one or two queues (scripts/probes/h2d_batch_probe.cu).  Fallback: a copy per stream,
    // alternating between two upload queues (the second one joins the first before the ready event).
    static const bool noBatch = [] { const char* e = getenv("VLOAM_UPLOAD_BATCH"); return e && atoi(e) == 0; }();
    bool done = false;
    if (!noBatch && !c->batch_copy_unsupported && h->B > 1) {
      std::vector<void*> dsts, srcs;
      std::vector<size_t> sizes;
      for (int b = 0; b < h->B; ++b)
        if (n_points[b]) {
          dsts.push_back(h->d_in[slot] + (size_t)b * h->cap * stride);
          srcs.push_back(const_cast<float*>(src(b)));
          sizes.push_back((size_t)n_points[b] * stride * sizeof(float));
        }
      // only for pinned (page-locked / registered) sources: a pageable buffer keeps cudaMemcpyAsync's semantics (staged before
      // the call returns, so the caller may reuse it at once)
      bool pinned = true;
      for (void* sp : srcs) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, sp) != cudaSuccess || pa.type != cudaMemoryTypeHost) { pinned = false; break; }
      }
      (void)cudaGetLastError();
      if (dsts.empty()) done = true;
      else if (pinned) {
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;     // the caller's buffers are read in stream order, like cudaMemcpyAsync
        size_t attrIdx = 0, failIdx = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &at, &attrIdx, 1, &failIdx, c->copy_stream);
        if (e == cudaSuccess) done = true;
        else { (void)cudaGetLastError(); c->batch_copy_unsupported = true; }   // older driver / pageable source: per-stream copies
      }
    }
    if (!done) {
      if (h->in_used[slot]) CU(c, cudaStreamWaitEvent(c->copy_stream2, h->ev_in_free[slot], 0));
      for (int b = 0; b < h->B; ++b)
        if (n_points[b])
          CU(c, cudaMemcpyAsync(h->d_in[slot] + (size_t)b * h->cap * stride, src(b),
                                (size_t)n_points[b] * stride * sizeof(float), cudaMemcpyHostToDevice, (b & 1) ? c->copy_stream2 : c->copy_stream));
      CU(c, cudaEventRecord(c->ev_copy2, c->copy_stream2));
      CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy2, 0));
    }
  }
  CU(c, cudaMemcpyAsync(h->d_n[slot], n_points, h->B * sizeof(int), cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaEventRecord(h->ev_in_ready[slot], c->copy_stream));
  CU(c, cudaStreamWaitEvent(c->stream, h->ev_in_ready[slot], 0));
  return VLOAM_OK;
}
// ... and run scan registration on it
template <typename Src>
static int upload_and_register(vloam_lidar* h, Src src, const float* contiguous, const int* n_points, int stride, size_t slab_points) {
  vloam_ctx* c = h->ctx;
  const int slot = (int)(h->host_scans & 1);
  int r = upload_only(h, src, contiguous, n_points, stride, slab_points);
  if (r) return r;
  r = run_scan_registration(h, h->d_in[slot], h->d_n[slot], stride, (size_t)h->cap);
  if (r) return r;
  CU(c, cudaEventRecord(h->ev_in_free[slot], c->stream));
  h->in_used[slot] = true;
  h->last_stride = stride;
  h->host_scans++;
  return VLOAM_OK;
}

extern "C" {

int vloam_scan_registration(vloam_lidar* h, const float* xyz, const int* n_points, int stride, size_t slab_points) {
  if (!h || !xyz || !n_points || stride < 3 || stride > kMaxInputStride) return VLOAM_E_INVALID;
  return upload_and_register(h, [&](int b) { return xyz + (size_t)b * slab_points * stride; }, xyz, n_points, stride, slab_points);
}

int vloam_scan_registration_ptrs(vloam_lidar* h, const float* const* xyz_ptrs, const int* n_points, int stride) {
  if (!h || !xyz_ptrs || !n_points || stride < 3 || stride > kMaxInputStride) return VLOAM_E_INVALID;
  for (int b = 0; b < h->B; ++b) if (n_points[b] > 0 && !xyz_ptrs[b]) return VLOAM_E_INVALID;
  return upload_and_register(h, [&](int b) { return xyz_ptrs[b]; }, nullptr, n_points, stride, (size_t)h->cap);
}

int vloam_get_input_device(vloam_lidar* h, const float** xyz_dev, const int** n_dev, int* stride_floats, size_t* slab_points) {
  if (!h || !xyz_dev || !n_dev) return VLOAM_E_INVALID;
  if (h->host_scans == 0) return fail(h->ctx, VLOAM_E_STATE, "vloam_get_input_device before vloam_scan_registration");
  const int slot = (int)((h->host_scans - 1) & 1);
  *xyz_dev = h->d_in[slot]; *n_dev = h->d_n[slot];
  if (stride_floats) *stride_floats = h->last_stride;
  if (slab_points) *slab_points = (size_t)h->cap;
  return VLOAM_OK;
}

int vloam_input_consumed(vloam_lidar* h) {
  if (!h) return VLOAM_E_INVALID;
  if (h->host_scans == 0) return fail(h->ctx, VLOAM_E_STATE, "vloam_input_consumed before vloam_scan_registration");
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventRecord(h->ev_in_free[(h->host_scans - 1) & 1], c->stream));
  return VLOAM_OK;
}

int vloam_scan_registration_device(vloam_lidar* h, const float* xyz_dev, const int* n_dev, int stride, size_t slab_points) {
  if (!h || !xyz_dev || !n_dev || stride < 3 || stride > kMaxInputStride || slab_points == 0) return VLOAM_E_INVALID;
  CU(h->ctx, cudaSetDevice(h->ctx->device));
  // the counts live in device memory: the kernels clamp them to min(max_points, slab_points) and report VLOAM_STREAM_CAPACITY
  return run_scan_registration(h, xyz_dev, n_dev, stride, slab_points);
}

static int fetch_headers(vloam_lidar* h, std::vector<SRHeader>* out, int slot) {
  vloam_ctx* c = h->ctx;
  out->resize(h->B);
  CU(c, cudaMemcpyAsync(out->data(), h->d_hdr[slot], h->B * sizeof(SRHeader), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

int vloam_get_stream_status(vloam_lidar* h, int* status) {
  if (!h || !status) return VLOAM_E_INVALID;
  if (h->frame < 0) return fail(h->ctx, VLOAM_E_STATE, "no scan registered yet");
  std::vector<SRHeader> hd;
  int r = fetch_headers(h, &hd, h->cur());
  if (r) return r;
  for (int b = 0; b < h->B; ++b) status[b] = hd[b].status;
  return VLOAM_OK;
}

int vloam_get_feature_counts(vloam_lidar* h, int* counts) {
  if (!h || !counts) return VLOAM_E_INVALID;
  if (h->frame < 0) return fail(h->ctx, VLOAM_E_STATE, "no scan registered yet");
  std::vector<SRHeader> hd;
  int r = fetch_headers(h, &hd, h->cur());
  if (r) return r;
  for (int b = 0; b < h->B; ++b) {
    counts[b * 5 + 0] = hd[b].cloudSize; counts[b * 5 + 1] = hd[b].nSharp; counts[b * 5 + 2] = hd[b].nLessSharp;
    counts[b * 5 + 3] = hd[b].nFlat; counts[b * 5 + 4] = hd[b].nLessFlat;
  }
  return VLOAM_OK;
}

static int copy_out(vloam_ctx* c, void* dst, const void* src_dev, size_t elem, int n, int capacity, int* n_out) {
  if (n_out) *n_out = n;
  const int m = n < capacity ? n : capacity;
  if (m > 0 && dst) {
    CU(c, cudaMemcpyAsync(dst, src_dev, (size_t)m * elem, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return VLOAM_OK;
}

int vloam_get_cloud(vloam_lidar* h, int stream, int which, float* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  if (h->frame < 0) return fail(c, VLOAM_E_STATE, "no scan registered yet");
  CU(c, cudaSetDevice(c->device));
  if (which == VLOAM_CLOUD_MAP)
    return lm_get_map_cloud(h->lm, c->stream, stream, out, capacity, n_out) == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "lm_get_map_cloud");
  if (which == VLOAM_CLOUD_FULL_REGISTERED) {
    std::vector<SRHeader> hd0;
    int r0 = fetch_headers(h, &hd0, h->cur());
    if (r0) return r0;
    return lm_get_registered(h->lm, c->stream, stream, h->d_cloud[h->cur()] + (size_t)stream * h->cap, hd0[stream].cloudSize, out, capacity, n_out) == cudaSuccess
               ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "lm_get_registered");
  }
  if (which >= VLOAM_CLOUD_CORNER_STACK) return lm_get_cloud(h->lm, c->stream, stream, which, out, capacity, n_out) == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "lm_get_cloud");
  std::vector<SRHeader> hd;
  // CORNER_LAST / SURF_LAST are the current scan's less-sharp / less-flat clouds once laser odometry has run
  // (the swap at laser_odometry.cpp:511-517), the previous scan's before.
  int slot = h->cur();
  if ((which == VLOAM_CLOUD_CORNER_LAST || which == VLOAM_CLOUD_SURF_LAST) && !h->lo_done_for_frame) {
    if (h->frame < 1) { if (n_out) *n_out = 0; return VLOAM_OK; }
    slot ^= 1;
  }
  int r = fetch_headers(h, &hd, slot);
  if (r) return r;
  const SRHeader& H = hd[stream];
  const size_t b = stream;
  switch (which) {
    case VLOAM_CLOUD_FULL: return copy_out(c, out, h->d_cloud[slot] + b * h->cap, 16, H.cloudSize, capacity, n_out);
    case VLOAM_CLOUD_SHARP: return copy_out(c, out, h->d_sharp + b * kMaxSharp, 16, H.nSharp, capacity, n_out);
    case VLOAM_CLOUD_LESS_SHARP:
    case VLOAM_CLOUD_CORNER_LAST: return copy_out(c, out, h->d_lessSharp[slot] + b * kMaxLessSharp, 16, H.nLessSharp, capacity, n_out);
    case VLOAM_CLOUD_FLAT: return copy_out(c, out, h->d_flat + b * kMaxFlat, 16, H.nFlat, capacity, n_out);
    case VLOAM_CLOUD_LESS_FLAT:
    case VLOAM_CLOUD_SURF_LAST: return copy_out(c, out, h->d_lessFlat[slot] + b * h->cap, 16, H.nLessFlat, capacity, n_out);
  }
  return fail(c, VLOAM_E_INVALID, "vloam_get_cloud: unknown selector");
}

int vloam_get_curvature(vloam_lidar* h, int stream, float* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  if (h->frame < 0) return fail(h->ctx, VLOAM_E_STATE, "no scan registered yet");
  std::vector<SRHeader> hd;
  int r = fetch_headers(h, &hd, h->cur());
  if (r) return r;
  return copy_out(h->ctx, out, h->d_curv + (size_t)stream * h->cap, 4, hd[stream].cloudSize, capacity, n_out);
}
int vloam_get_labels(vloam_lidar* h, int stream, int8_t* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  if (h->frame < 0) return fail(h->ctx, VLOAM_E_STATE, "no scan registered yet");
  std::vector<SRHeader> hd;
  int r = fetch_headers(h, &hd, h->cur());
  if (r) return r;
  return copy_out(h->ctx, out, h->d_label + (size_t)stream * h->cap, 1, hd[stream].cloudSize, capacity, n_out);
}
int vloam_get_feature_indices(vloam_lidar* h, int stream, int which, int* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  if (h->frame < 0) return fail(h->ctx, VLOAM_E_STATE, "no scan registered yet");
  std::vector<SRHeader> hd;
  int r = fetch_headers(h, &hd, h->cur());
  if (r) return r;
  const size_t b = stream;
  switch (which) {
    case VLOAM_CLOUD_SHARP: return copy_out(h->ctx, out, h->d_sharpIdx + b * kMaxSharp, 4, hd[stream].nSharp, capacity, n_out);
    case VLOAM_CLOUD_LESS_SHARP: return copy_out(h->ctx, out, h->d_lessSharpIdx + b * kMaxLessSharp, 4, hd[stream].nLessSharp, capacity, n_out);
    case VLOAM_CLOUD_FLAT: return copy_out(h->ctx, out, h->d_flatIdx + b * kMaxFlat, 4, hd[stream].nFlat, capacity, n_out);
  }
  return fail(h->ctx, VLOAM_E_INVALID, "vloam_get_feature_indices: which must be SHARP, LESS_SHARP or FLAT");
}

// ------------------------------------------------------------------------------------------------ laser odometry
static int run_laser_odometry(vloam_lidar* h, const double* prior_dev) {
  vloam_ctx* c = h->ctx;
  (void)cudaGetLastError();
  if (h->frame < 0) return fail(c, VLOAM_E_STATE, "laser odometry before scan registration");
  if (h->lo_done_for_frame) return fail(c, VLOAM_E_STATE, "laser odometry already run for this scan");
  const int cur = h->cur(), last = cur ^ 1;
  if (h->frame >= 1) {  // systemInited (laser_odometry.cpp:196-205)
    const double* prior = h->p.detach_VO_LO ? nullptr : prior_dev;
    const int passes = h->p.lo_outer_passes;
    for (int pass = 0; pass < passes; ++pass) {
      launch_lo_pass(&c->prof, c->stream, h->B, h->cap, h->d_hdr[cur], h->d_hdr[last], h->d_lo, h->d_sharp, h->d_flat,
                     h->d_lessSharp[last], h->d_lessFlat[last], &h->grid, h->d_corr[pass < 2 ? pass : 1], pass < 2 ? pass : 1,
                     h->p.lo_max_iterations, pass == passes - 1, prior, h->shard.world > 1 ? &h->shard : nullptr, h->p.solver_mode, h->p.distortion != 0);   // (with an NCCL communicator
                                                                                       // the pass takes the wide solve and sums there)
    }
  }
  // laser_odometry.cpp:511-526: the current less-sharp / less-flat clouds become "last" and are indexed for the next scan
  launch_lo_build_grid(&c->prof, c->stream, h->B, h->cap, h->d_hdr[cur], h->d_lessSharp[cur], h->d_lessFlat[cur], &h->grid);
  launch_lo_export(&c->prof, c->stream, h->d_lo, h->d_pose, h->B);
  CU(c, cudaGetLastError());
  {  // asynchronous read-back of the poses into the pinned buffer of this frame's parity
    const int par = (int)(h->frame & 1);
    CU(c, cudaMemcpyAsync(h->h_pose[par], h->d_pose, (size_t)h->B * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (!h->capturing) CU(c, cudaEventRecord(h->ev_pose[par], c->stream));   // (a captured frame records it after the graph launch)
    h->pose_valid[par] = true;
  }
  h->lo_done_for_frame = true;
  h->lo_frames++;
  return VLOAM_OK;
}

int vloam_laser_odometry_async(vloam_lidar* h, const double* prior_dev) {
  if (!h) return VLOAM_E_INVALID;
  CU(h->ctx, cudaSetDevice(h->ctx->device));
  return run_laser_odometry(h, prior_dev);
}

static int read_pose(vloam_lidar* h, int par, double* pose_out, int* corr_out) {
  vloam_ctx* c = h->ctx;
  if (!h->pose_valid[par]) return fail(c, VLOAM_E_STATE, "no laser odometry result for that frame yet");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(h->ev_pose[par]));  // waits for that frame's read-back only, not for the whole stream
  if (h->shard.world > 1 && h->shard.error != nullptr && h->nccl == nullptr) {   // peer-memory exchange: a peer that did not answer inside the solve kernel makes the result invalid
    int bits = 0;
    CU(c, cudaMemcpy(&bits, h->shard.error, sizeof(int), cudaMemcpyDeviceToHost));
    if (bits) return fail(c, VLOAM_E_STATE, "point-sharded solve: a peer rank did not publish its normal equations in time (vloam_shard_status)");
  }
  for (int b = 0; b < h->B; ++b) {
    if (pose_out) std::memcpy(pose_out + (size_t)b * 14, h->h_pose[par] + (size_t)b * 16, 14 * sizeof(double));
    if (corr_out) { corr_out[b * 2] = (int)h->h_pose[par][b * 16 + 14]; corr_out[b * 2 + 1] = (int)h->h_pose[par][b * 16 + 15]; }
  }
  return VLOAM_OK;
}
int vloam_get_lo_pose(vloam_lidar* h, double* pose_out, int* corr_out) {
  if (!h) return VLOAM_E_INVALID;
  if (!h->lo_done_for_frame) return fail(h->ctx, VLOAM_E_STATE, "laser odometry has not run for the current scan");
  return read_pose(h, (int)(h->frame & 1), pose_out, corr_out);
}
int vloam_get_lo_pose_prev(vloam_lidar* h, double* pose_out, int* corr_out) {
  if (!h) return VLOAM_E_INVALID;
  if (h->frame < 1) return fail(h->ctx, VLOAM_E_STATE, "no previous scan");
  return read_pose(h, (int)((h->frame - 1) & 1), pose_out, corr_out);
}

int vloam_laser_odometry(vloam_lidar* h, const double* prior, double* pose_out, int* corr_out) {
  if (!h) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  const double* prior_dev = nullptr;
  if (prior && !h->p.detach_VO_LO) {
    CU(c, cudaMemcpyAsync(h->d_prior, prior, (size_t)h->B * 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    prior_dev = h->d_prior;
  }
  int r = run_laser_odometry(h, prior_dev);
  if (r) return r;
  if (pose_out || corr_out) return vloam_get_lo_pose(h, pose_out, corr_out);
  return VLOAM_OK;
}

int vloam_set_lo_motion(vloam_lidar* h, const double* motion) {
  if (!h || !motion) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(h->d_prior, motion, (size_t)h->B * 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_lo_set_motion(&c->prof, c->stream, h->d_lo, h->d_prior, h->B);
  CU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

int vloam_set_lo_pose(vloam_lidar* h, const double* pose) {
  if (!h || !pose) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(h->d_prior, pose, (size_t)h->B * 7 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_lo_set_pose(&c->prof, c->stream, h->d_lo, h->d_prior, h->B);
  CU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

int vloam_get_lo_trace(vloam_lidar* h, int stream, int pass, int* corr, double* records, int* info, double* para) {
  if (!h || stream < 0 || stream >= h->B || pass < 0 || pass > 1) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  LOState st;
  CU(c, cudaMemcpyAsync(&st, h->d_lo + stream, sizeof(LOState), cudaMemcpyDeviceToHost, c->stream));
  if (corr)
    CU(c, cudaMemcpyAsync(corr, h->d_corr[pass] + (size_t)stream * (kMaxSharp + kMaxFlat),
                          (size_t)(kMaxSharp + kMaxFlat) * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  const SolveTrace& t = st.trace[pass];
  if (info) { info[0] = t.n_records; info[1] = t.termination; info[2] = t.n_corner; info[3] = t.n_plane; }
  if (para) std::memcpy(para, t.para, sizeof(t.para));
  if (records) std::memcpy(records, t.rec, sizeof(t.rec));
  return VLOAM_OK;
}

// ------------------------------------------------------------------------------------------------ laser mapping
int vloam_laser_mapping(vloam_lidar* h, double* pose_out) {
  if (!h) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  if (!h->lo_done_for_frame) return fail(c, VLOAM_E_STATE, "laser mapping before laser odometry");
  // LaserOdometry::output (laser_odometry.cpp:618-628): frames with frameCount % mapping_skip_frame != 0 are skipped
  const bool skip = (h->lo_frames % h->p.mapping_skip_frame) != 0;
  const int cur = h->cur();
  cudaError_t e = lm_run(h->lm, c->stream, h->d_hdr[cur], h->d_lessSharp[cur], h->d_lessFlat[cur], h->d_lo, skip);
  if (e != cudaSuccess) return fail(c, VLOAM_E_CUDA, "vloam_laser_mapping", e);
  if (pose_out) return vloam_get_lm_pose(h, pose_out);
  return VLOAM_OK;
}
int vloam_get_lm_pose(vloam_lidar* h, double* pose_out) {
  if (!h || !pose_out) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_pose(h->lm, c->stream, pose_out);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_pose", e);
}

// ------------------------------------------------------------------------------------------------ one frame in one call
// The LiDAR part of vloam_main_node.cpp:125-180 — reset, scanRegistrationIO, laserOdometryIO, laserMappingIO — for the scan
// already sitting in input slot `slot`.  With use_graph the ~40 launches are captured into a CUDA graph the first time a
// (buffer parity, odometry initialised, mapping skipped) combination occurs and replayed afterwards: one launch per frame.
static int process_frame(vloam_lidar* h, int slot, int stride, const double* prior_dev, int use_graph) {
  vloam_ctx* c = h->ctx;
  const bool inited = h->frame + 1 >= 1;                       // LaserOdometry::systemInited for the frame about to run
  const bool skip = ((h->lo_frames + 1) % h->p.mapping_skip_frame) != 0;
  const int par = (int)((h->frame + 1) & 1);
  const bool graphable = use_graph && !c->prof.enabled && h->shard.world <= 1 && lm_graph_safe(h->lm) && par == slot;
  auto run_direct = [&]() -> int {
    vloam_lidar_reset(h);
    int r = run_scan_registration(h, h->d_in[slot], h->d_n[slot], stride, (size_t)h->cap);
    if (r) return r;
    r = run_laser_odometry(h, prior_dev);
    if (r) return r;
    return vloam_laser_mapping(h, nullptr);
  };
  if (!graphable) return run_direct();
  cudaGraphExec_t& g = h->graph[par][inited ? 1 : 0][skip ? 1 : 0];
  if (g && h->graph_prior[par][inited ? 1 : 0][skip ? 1 : 0] != prior_dev) { cudaGraphExecDestroy(g); g = nullptr; }   // another prior buffer: capture again
  if (!g) {
    cudaError_t e = lm_ensure_alloc(h->lm, c->stream);         // no allocation inside a capture
    if (e != cudaSuccess) return fail(c, VLOAM_E_CUDA, "vloam_lidar_process: map allocation", e);
    CU(c, cudaStreamSynchronize(c->stream));
    const long long launches0 = c->prof.launches;
    CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
    const int r = run_direct();                                // records the launches, advances the host-side frame state
    h->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess) return fail(c, VLOAM_E_CUDA, "vloam_lidar_process: stream capture", ce);
    const cudaError_t ie = cudaGraphInstantiate(&g, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { g = nullptr; return fail(c, VLOAM_E_CUDA, "vloam_lidar_process: graph instantiation", ie); }
    h->graph_prior[par][inited ? 1 : 0][skip ? 1 : 0] = prior_dev;
    h->graph_launches[par][inited ? 1 : 0][skip ? 1 : 0] = c->prof.launches - launches0;
  } else {
    // replay: the host-side state the three stages leave behind
    vloam_lidar_reset(h);
    h->frame++;
    h->pose_valid[par] = true;
    h->lo_done_for_frame = true;
    h->lo_frames++;
    lm_note_run(h->lm, skip);
    c->prof.launches += h->graph_launches[par][inited ? 1 : 0][skip ? 1 : 0];
  }
  CU(c, cudaGraphLaunch(g, c->stream));
  CU(c, cudaEventRecord(h->ev_pose[par], c->stream));          // the frame's poses are in the pinned buffer once the graph has run
  return VLOAM_OK;
}

int vloam_lidar_process(vloam_lidar* h, const float* xyz, const int* n_points, int stride, size_t slab_points, const double* prior_dev,
                        int use_graph) {
  if (!h || !xyz || !n_points || stride < 3 || stride > kMaxInputStride) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  // the upload is the front half of vloam_scan_registration; the kernels then run from the slot (directly or as a graph)
  h->host_scans = h->frame + 1;                                 // input slot parity == frame parity on this path
  const int slot = (int)(h->host_scans & 1);
  const int r = upload_only(h, [&](int b) { return xyz + (size_t)b * slab_points * stride; }, xyz, n_points, stride, slab_points);
  if (r) return r;
  const int r2 = process_frame(h, slot, stride, h->p.detach_VO_LO ? nullptr : prior_dev, use_graph);
  if (r2) return r2;
  CU(c, cudaEventRecord(h->ev_in_free[slot], c->stream));
  h->in_used[slot] = true; h->last_stride = stride; h->host_scans++;
  return VLOAM_OK;
}

int vloam_lidar_process_ptrs(vloam_lidar* h, const float* const* xyz_ptrs, const int* n_points, int stride, const double* prior_dev, int use_graph) {
  if (!h || !xyz_ptrs || !n_points || stride < 3 || stride > kMaxInputStride) return VLOAM_E_INVALID;
  for (int b = 0; b < h->B; ++b) if (n_points[b] > 0 && !xyz_ptrs[b]) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  h->host_scans = h->frame + 1;
  const int slot = (int)(h->host_scans & 1);
  const int r = upload_only(h, [&](int b) { return xyz_ptrs[b]; }, nullptr, n_points, stride, (size_t)h->cap);
  if (r) return r;
  const int r2 = process_frame(h, slot, stride, h->p.detach_VO_LO ? nullptr : prior_dev, use_graph);
  if (r2) return r2;
  CU(c, cudaEventRecord(h->ev_in_free[slot], c->stream));
  h->in_used[slot] = true; h->last_stride = stride; h->host_scans++;
  return VLOAM_OK;
}

int vloam_lidar_process_device(vloam_lidar* h, const float* xyz_dev, const int* n_dev, int stride, size_t slab_points, const double* prior_dev,
                               int use_graph) {
  if (!h || !xyz_dev || !n_dev || stride < 3 || stride > kMaxInputStride || slab_points == 0) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  if (!use_graph) {
    vloam_lidar_reset(h);
    int r = run_scan_registration(h, xyz_dev, n_dev, stride, slab_points);
    if (r) return r;
    r = run_laser_odometry(h, h->p.detach_VO_LO ? nullptr : prior_dev);
    if (r) return r;
    return vloam_laser_mapping(h, nullptr);
  }
  // a graph replays fixed addresses: the scan is first copied (device to device) into the handle's input slot
  if (stride > h->in_stride) return fail(c, VLOAM_E_INVALID, "vloam_lidar_process_device with graphs: stride above the handle's input stride (4)");
  const int slot = (int)((h->frame + 1) & 1);
  const size_t rowBytes = (slab_points < (size_t)h->cap ? slab_points : (size_t)h->cap) * stride * sizeof(float);
  CU(c, cudaMemcpy2DAsync(h->d_in[slot], (size_t)h->cap * stride * sizeof(float), xyz_dev, slab_points * stride * sizeof(float), rowBytes, (size_t)h->B,
                          cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaMemcpyAsync(h->d_n[slot], n_dev, (size_t)h->B * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  h->host_scans = h->frame + 2; h->last_stride = stride;
  return process_frame(h, slot, stride, h->p.detach_VO_LO ? nullptr : prior_dev, use_graph);
}

int vloam_map_set_cube(vloam_lidar* h, int stream, int kind, int cube, const float* xyzi, int n) {
  if (!h || stream < 0 || stream >= h->B || kind < 0 || kind > 1 || cube < 0 || cube >= 4851 || n < 0 || (n && !xyzi)) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_set_cube(h->lm, c->stream, stream, kind, cube, xyzi, n);
  if (e == cudaErrorInvalidValue) return fail(c, VLOAM_E_INVALID, "vloam_map_set_cube: a point lies outside the named cube");
  if (e == cudaErrorMemoryAllocation) return fail(c, VLOAM_E_CAPACITY, "vloam_map_set_cube: map_capacity_points exceeded");
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_map_set_cube", e);
}
int vloam_map_get_cube(vloam_lidar* h, int stream, int kind, int cube, float* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B || kind < 0 || kind > 1 || cube < 0 || cube >= 4851) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_cube(h->lm, c->stream, stream, kind, cube, out, capacity, n_out);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_map_get_cube", e);
}
int vloam_get_lm_info(vloam_lidar* h, int* info) {
  if (!h || !info) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_info(h->lm, c->stream, info);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_info", e);
}
int vloam_get_map_stats(vloam_lidar* h, int* stats) {
  if (!h || !stats) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_map_stats(h->lm, c->stream, stats);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_map_stats", e);
}
int vloam_get_lm_status(vloam_lidar* h, int* status) {
  if (!h || !status) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_status(h->lm, c->stream, status);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_status", e);
}
int vloam_get_lm_counters(vloam_lidar* h, long long* counters) {
  if (!h || !counters) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_counters(h->lm, c->stream, counters);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_counters", e);
}
int vloam_lidar_set_debug_stats(vloam_lidar* h, int on) {
  if (!h) return VLOAM_E_INVALID;
  lm_set_debug_stats(h->lm, on != 0);
  return VLOAM_OK;
}
int vloam_get_lm_queries(vloam_lidar* h, int stream, int pass, int kind, int* out, int capacity, int* n_out) {
  if (!h || stream < 0 || stream >= h->B || pass < 0 || pass > 1 || kind < 0 || kind > 1 || capacity < 0) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_queries(h->lm, c->stream, stream, pass, kind, out, capacity, n_out);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_queries", e);
}
int vloam_get_lm_trace(vloam_lidar* h, int stream, int pass, double* records, int* info, double* para) {
  if (!h || stream < 0 || stream >= h->B || pass < 0 || pass > 1) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  CU(c, cudaSetDevice(c->device));
  cudaError_t e = lm_get_trace(h->lm, c->stream, stream, pass, records, info, para);
  return e == cudaSuccess ? VLOAM_OK : fail(c, VLOAM_E_CUDA, "vloam_get_lm_trace", e);
}

}  // extern "C"
