// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// capi.cpp: extern "C" surface of the CPU oracle, loaded with ctypes by
// oracle/pyoracle.py from tests/, bench.py (cpu_baseline / --impl reference)
// and __graft_entry__.smoke().  Never linked into the product library.
#include <chrono>
#include <cstring>

#include "laser_mapping.hpp"
#include "laser_odometry.hpp"
#include "scan_registration.hpp"
#include "visual_odometry.hpp"

using namespace oracle;

namespace {
const Cloud* sr_cloud(const ScanRegistrationOutput* o, int which) {
  switch (which) {
    case 0: return &o->laserCloud;
    case 1: return &o->cornerPointsSharp;
    case 2: return &o->cornerPointsLessSharp;
    case 3: return &o->surfPointsFlat;
    case 4: return &o->surfPointsLessFlat;
  }
  return nullptr;
}
Cloud make_cloud(const float* xyzi, int n) {
  Cloud c(n);
  if (n) std::memcpy(c.data(), xyzi, sizeof(PointXYZI) * (size_t)n);
  return c;
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------- scan registration
void* orc_sr_run(const float* xyz, int n, int stride, int n_scans, double min_range, int literal_unstable) {
  auto* o = new ScanRegistrationOutput();
  scan_registration(xyz, n, stride, n_scans, min_range, o, literal_unstable != 0);
  return o;
}
void orc_sr_free(void* h) { delete static_cast<ScanRegistrationOutput*>(h); }
int orc_sr_status(void* h) { return static_cast<ScanRegistrationOutput*>(h)->status; }
int orc_sr_count(void* h, int which) { return (int)sr_cloud(static_cast<ScanRegistrationOutput*>(h), which)->size(); }
void orc_sr_copy_cloud(void* h, int which, float* out) {
  const Cloud* c = sr_cloud(static_cast<ScanRegistrationOutput*>(h), which);
  if (!c->empty()) std::memcpy(out, c->data(), sizeof(PointXYZI) * c->size());
}
// which: 0 curvature(float) 1 label 2 picked 3 scanStartInd 4 scanEndInd 5 sharpInd 6 lessSharpInd 7 flatInd 8 ringLessFlatCount
int orc_sr_array_len(void* h, int which) {
  auto* o = static_cast<ScanRegistrationOutput*>(h);
  switch (which) {
    case 0: return (int)o->curvature.size();
    case 1: return (int)o->label.size();
    case 2: return (int)o->picked.size();
    case 3: return (int)o->scanStartInd.size();
    case 4: return (int)o->scanEndInd.size();
    case 5: return (int)o->sharpInd.size();
    case 6: return (int)o->lessSharpInd.size();
    case 7: return (int)o->flatInd.size();
    case 8: return (int)o->ringLessFlatCount.size();
  }
  return 0;
}
void orc_sr_copy_array(void* h, int which, void* out) {
  auto* o = static_cast<ScanRegistrationOutput*>(h);
  auto cp = [&](const void* p, size_t bytes) { if (bytes) std::memcpy(out, p, bytes); };
  switch (which) {
    case 0: cp(o->curvature.data(), o->curvature.size() * 4); break;
    case 1: cp(o->label.data(), o->label.size() * 4); break;
    case 2: cp(o->picked.data(), o->picked.size() * 4); break;
    case 3: cp(o->scanStartInd.data(), o->scanStartInd.size() * 4); break;
    case 4: cp(o->scanEndInd.data(), o->scanEndInd.size() * 4); break;
    case 5: cp(o->sharpInd.data(), o->sharpInd.size() * 4); break;
    case 6: cp(o->lessSharpInd.data(), o->lessSharpInd.size() * 4); break;
    case 7: cp(o->flatInd.data(), o->flatInd.size() * 4); break;
    case 8: cp(o->ringLessFlatCount.data(), o->ringLessFlatCount.size() * 4); break;
  }
}

// ---------------------------------------------------------------- voxel grid / kNN / small linalg (unit-test hooks)
int orc_voxel_grid(const float* xyzi, int n, float leaf, int literal_unstable, float* out) {
  Cloud in = make_cloud(xyzi, n), o;
  voxel_grid_filter(in, leaf, &o, literal_unstable != 0);
  if (!o.empty()) std::memcpy(out, o.data(), sizeof(PointXYZI) * o.size());
  return (int)o.size();
}
void orc_knn(const float* target_xyzi, int nt, const float* query_xyzi, int nq, int k, int brute, int* out_idx,
             float* out_d) {
  Cloud t = make_cloud(target_xyzi, nt), q = make_cloud(query_xyzi, nq);
  KdTree tree;
  if (!brute) tree.set_input(&t);
  for (int i = 0; i < nq; ++i) {
    int got = brute ? brute_nearest_k(t, q[i], k, out_idx + (size_t)i * k, out_d + (size_t)i * k)
                    : tree.nearest_k(q[i], k, out_idx + (size_t)i * k, out_d + (size_t)i * k);
    for (int j = got; j < k; ++j) { out_idx[(size_t)i * k + j] = -1; out_d[(size_t)i * k + j] = -1.f; }
  }
}
void orc_sym_eig3(const double* A, double* evals, double* evecs) {
  double ev[3][3];
  sym_eig3(A, evals, ev);
  std::memcpy(evecs, ev, sizeof(ev));
}
void orc_colpiv_qr_solve_5x3(const double* A, const double* b, double* x) { colpiv_qr_solve3<5>(A, b, x); }

// Evaluate one lidar factor at x = [q(xyzw), t]: kind 0 edge (pts = cp,a,b), 1 plane (cp,j,l,m), 2 plane-norm (cp,n,[d]).
// Writes r[<=3] and the 7-column auto-diff Jacobian J[nres*7].  Returns nres.
int orc_factor_eval(int kind, const double* pts, const double* x, double* r, double* J) {
  auto v = [&](int i) { return Vec3{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}; };
  if (kind == 0) {
    LidarEdgeFunctor f; f.curr_point = v(0); f.last_point_a = v(1); f.last_point_b = v(2); f.s = 1.0;
    AutoDiffBlock43<LidarEdgeFunctor, 3>(f).evaluate(x, r, J);
    return 3;
  } else if (kind == 1) {
    AutoDiffBlock43<LidarPlaneFunctor, 1>(LidarPlaneFunctor(v(0), v(1), v(2), v(3), 1.0)).evaluate(x, r, J);
    return 1;
  }
  LidarPlaneNormFunctor f; f.curr_point = v(0); f.plane_unit_norm = v(1); f.negative_OA_dot_norm = pts[6];
  AutoDiffBlock43<LidarPlaneNormFunctor, 1>(f).evaluate(x, r, J);
  return 1;
}

// ---------------------------------------------------------------- laser odometry
void* orc_lo_create(int detach_VO_LO, int mapping_skip_frame) {
  auto* lo = new LaserOdometry();
  lo->detach_VO_LO = detach_VO_LO != 0;
  lo->mapping_skip_frame = mapping_skip_frame;
  return lo;
}
void orc_lo_free(void* h) { delete static_cast<LaserOdometry*>(h); }
void orc_lo_set_iterations(void* h, int passes, int lm_iters) {
  auto* lo = static_cast<LaserOdometry*>(h);
  lo->num_outer_passes = passes;
  lo->lm_max_iterations = lm_iters;
}
void orc_lo_solve_sr(void* h, void* sr, const double* prior_q, const double* prior_t) {
  auto* lo = static_cast<LaserOdometry*>(h);
  auto* o = static_cast<ScanRegistrationOutput*>(sr);
  lo->solveLO(o->laserCloud, o->cornerPointsSharp, o->cornerPointsLessSharp, o->surfPointsFlat, o->surfPointsLessFlat,
              prior_q, prior_t);
}
void orc_lo_solve_clouds(void* h, const float* full, int nfull, const float* sharp, int nsharp, const float* less_sharp,
                         int nless_sharp, const float* flat, int nflat, const float* less_flat, int nless_flat,
                         const double* prior_q, const double* prior_t) {
  auto* lo = static_cast<LaserOdometry*>(h);
  lo->solveLO(make_cloud(full, nfull), make_cloud(sharp, nsharp), make_cloud(less_sharp, nless_sharp),
              make_cloud(flat, nflat), make_cloud(less_flat, nless_flat), prior_q, prior_t);
}
// pose[18]: q_last_curr(4) t_last_curr(3) q_w_curr(4) t_w_curr(3) corner_corr plane_corr frameCount systemInited
void orc_lo_get_state(void* h, double* pose) {
  auto* lo = static_cast<LaserOdometry*>(h);
  for (int i = 0; i < 4; ++i) pose[i] = lo->para_q[i];
  for (int i = 0; i < 3; ++i) pose[4 + i] = lo->para_t[i];
  pose[7] = lo->q_w_curr.x; pose[8] = lo->q_w_curr.y; pose[9] = lo->q_w_curr.z; pose[10] = lo->q_w_curr.w;
  pose[11] = lo->t_w_curr.x; pose[12] = lo->t_w_curr.y; pose[13] = lo->t_w_curr.z;
  pose[14] = lo->corner_correspondence; pose[15] = lo->plane_correspondence;
  pose[16] = lo->frameCount; pose[17] = lo->systemInited;
}
void orc_lo_set_distortion(void* h, int on) { static_cast<LaserOdometry*>(h)->DISTORTION = on != 0; }
void orc_lo_set_skip(void* h, int mapping_skip_frame) { static_cast<LaserOdometry*>(h)->mapping_skip_frame = mapping_skip_frame; }
// checkpoint / resume hook mirrored by vloam_set_lo_pose: overwrite the accumulated odometry pose
void orc_lo_set_pose(void* h, const double* q, const double* t) {
  auto* lo = static_cast<LaserOdometry*>(h);
  lo->q_w_curr = Quat{q[0], q[1], q[2], q[3]};
  lo->t_w_curr = Vec3{t[0], t[1], t[2]};
}
void orc_lo_set_motion(void* h, const double* q, const double* t) {
  auto* lo = static_cast<LaserOdometry*>(h);
  for (int i = 0; i < 4; ++i) lo->para_q[i] = q[i];
  for (int i = 0; i < 3; ++i) lo->para_t[i] = t[i];
}
int orc_lo_trace_passes(void* h) { return (int)static_cast<LaserOdometry*>(h)->trace.size(); }
// sizes[3] = #corner correspondences, #plane correspondences, #LM iteration records
void orc_lo_trace_sizes(void* h, int pass, int* sizes) {
  const LOPassTrace& t = static_cast<LaserOdometry*>(h)->trace[pass];
  sizes[0] = (int)t.corner.size() / 3; sizes[1] = (int)t.plane.size() / 4; sizes[2] = (int)t.summary.iterations.size();
}
// iters: per record [cost, candidate_cost, model_cost_change, relative_decrease, radius, valid, successful]; para[7] after pass
void orc_lo_trace_copy(void* h, int pass, int* corner, int* plane, double* iters, double* para, int* termination) {
  const LOPassTrace& t = static_cast<LaserOdometry*>(h)->trace[pass];
  if (!t.corner.empty()) std::memcpy(corner, t.corner.data(), t.corner.size() * 4);
  if (!t.plane.empty()) std::memcpy(plane, t.plane.data(), t.plane.size() * 4);
  for (size_t i = 0; i < t.summary.iterations.size(); ++i) {
    const LMIteration& it = t.summary.iterations[i];
    double* o = iters + i * 7;
    o[0] = it.cost; o[1] = it.candidate_cost; o[2] = it.model_cost_change; o[3] = it.relative_decrease;
    o[4] = it.radius; o[5] = it.step_is_valid; o[6] = it.step_is_successful;
  }
  for (int i = 0; i < 4; ++i) para[i] = t.para_q[i];
  for (int i = 0; i < 3; ++i) para[4 + i] = t.para_t[i];
  *termination = t.summary.termination;
}
int orc_lo_last_count(void* h, int which) {
  auto* lo = static_cast<LaserOdometry*>(h);
  return (int)(which == 0 ? lo->laserCloudCornerLast.size() : which == 1 ? lo->laserCloudSurfLast.size() : lo->laserCloudFullRes.size());
}
void orc_lo_last_copy(void* h, int which, float* out) {
  auto* lo = static_cast<LaserOdometry*>(h);
  const Cloud& c = which == 0 ? lo->laserCloudCornerLast : which == 1 ? lo->laserCloudSurfLast : lo->laserCloudFullRes;
  if (!c.empty()) std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}

// ---------------------------------------------------------------- laser mapping
void* orc_lm_create(double line_res, double plane_res) {
  auto* lm = new LaserMapping();
  lm->lineRes = line_res;
  lm->planeRes = plane_res;
  return lm;
}
void orc_lm_free(void* h) { delete static_cast<LaserMapping*>(h); }
void orc_lm_reset(void* h) { static_cast<LaserMapping*>(h)->reset(); }
void orc_lm_set_iterations(void* h, int passes, int lm_iters) {
  auto* lm = static_cast<LaserMapping*>(h);
  lm->num_outer_passes = passes;
  lm->lm_max_iterations = lm_iters;
}
void orc_lm_input_from_lo(void* h, void* lo_) {  // lidar_odometry_mapping.cpp:125-136 with LaserOdometry::output's skip flag
  auto* lm = static_cast<LaserMapping*>(h);
  auto* lo = static_cast<LaserOdometry*>(lo_);
  lm->input(lo->laserCloudCornerLast, lo->laserCloudSurfLast, lo->laserCloudFullRes, lo->q_w_curr, lo->t_w_curr, lo->skip_frame());
}
// pose[8]: the /aft_mapped_to_init pose of the last input() (q xyzw, t), skip_frame
void orc_lm_published_pose(void* h, double* pose) {
  auto* lm = static_cast<LaserMapping*>(h);
  lm->published_pose(pose);
  pose[7] = lm->skip_frame ? 1.0 : 0.0;
}
void orc_lm_input_clouds(void* h, const float* corner, int ncorner, const float* surf, int nsurf, const double* q_odom,
                         const double* t_odom) {
  auto* lm = static_cast<LaserMapping*>(h);
  lm->input(make_cloud(corner, ncorner), make_cloud(surf, nsurf), Cloud(), Quat{q_odom[0], q_odom[1], q_odom[2], q_odom[3]},
            Vec3{t_odom[0], t_odom[1], t_odom[2]});
}
void orc_lm_solve(void* h) { static_cast<LaserMapping*>(h)->solveMapping(); }
// state[14+]: parameters(7) q_wmap_wodom(4) t_wmap_wodom(3) cen(3) validNum
void orc_lm_get_state(void* h, double* s) {
  auto* lm = static_cast<LaserMapping*>(h);
  for (int i = 0; i < 7; ++i) s[i] = lm->parameters[i];
  s[7] = lm->q_wmap_wodom.x; s[8] = lm->q_wmap_wodom.y; s[9] = lm->q_wmap_wodom.z; s[10] = lm->q_wmap_wodom.w;
  s[11] = lm->t_wmap_wodom.x; s[12] = lm->t_wmap_wodom.y; s[13] = lm->t_wmap_wodom.z;
  s[14] = lm->laserCloudCenWidth; s[15] = lm->laserCloudCenHeight; s[16] = lm->laserCloudCenDepth;
  s[17] = lm->laserCloudValidNum;
}
// Seed one cube of the map directly (used to pre-build the 1 M-point map of BASELINE config 3).
void orc_lm_set_cube(void* h, int which, int cube, const float* xyzi, int n) {
  auto* lm = static_cast<LaserMapping*>(h);
  (which == 0 ? lm->laserCloudCornerArray : lm->laserCloudSurfArray)[cube] = make_cloud(xyzi, n);
}
int orc_lm_cube_count(void* h, int which, int cube) {
  auto* lm = static_cast<LaserMapping*>(h);
  return (int)(which == 0 ? lm->laserCloudCornerArray : lm->laserCloudSurfArray)[cube].size();
}
void orc_lm_cube_copy(void* h, int which, int cube, float* out) {
  auto* lm = static_cast<LaserMapping*>(h);
  const Cloud& c = (which == 0 ? lm->laserCloudCornerArray : lm->laserCloudSurfArray)[cube];
  if (!c.empty()) std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}
long long orc_lm_map_points(void* h, int which) {
  auto* lm = static_cast<LaserMapping*>(h);
  long long n = 0;
  for (const Cloud& c : (which == 0 ? lm->laserCloudCornerArray : lm->laserCloudSurfArray)) n += (long long)c.size();
  return n;
}
// which: 0 corner stack, 1 surf stack, 2 corner-from-map, 3 surf-from-map, 4 laserCloudFullRes (in the map frame once
// orc_lm_publish_registered has run for the frame), 5 the /laser_cloud_map concatenation
static Cloud lm_cloud_of(LaserMapping* lm, int which) {
  switch (which) {
    case 0: return lm->laserCloudCornerStack;
    case 1: return lm->laserCloudSurfStack;
    case 2: return lm->laserCloudCornerFromMap;
    case 3: return lm->laserCloudSurfFromMap;
    case 4: return lm->laserCloudFullRes;
    default: return lm->map_cloud();
  }
}
int orc_lm_cloud_count(void* h, int which) { return (int)lm_cloud_of(static_cast<LaserMapping*>(h), which).size(); }
void orc_lm_cloud_copy(void* h, int which, float* out) {
  const Cloud c = lm_cloud_of(static_cast<LaserMapping*>(h), which);
  if (!c.empty()) std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}
void orc_lm_publish_registered(void* h) { static_cast<LaserMapping*>(h)->publish_registered(); }
int orc_lm_trace_passes(void* h) { return (int)static_cast<LaserMapping*>(h)->trace.size(); }
void orc_lm_trace_sizes(void* h, int pass, int* sizes) {
  const LMPassTrace& t = static_cast<LaserMapping*>(h)->trace[pass];
  sizes[0] = t.corner_num; sizes[1] = t.surf_num; sizes[2] = (int)t.summary.iterations.size();
}
void orc_lm_trace_copy(void* h, int pass, int* corner_q, int* surf_q, double* iters, double* para, int* termination) {
  const LMPassTrace& t = static_cast<LaserMapping*>(h)->trace[pass];
  if (!t.corner_query.empty()) std::memcpy(corner_q, t.corner_query.data(), t.corner_query.size() * 4);
  if (!t.surf_query.empty()) std::memcpy(surf_q, t.surf_query.data(), t.surf_query.size() * 4);
  for (size_t i = 0; i < t.summary.iterations.size(); ++i) {
    const LMIteration& it = t.summary.iterations[i];
    double* o = iters + i * 7;
    o[0] = it.cost; o[1] = it.candidate_cost; o[2] = it.model_cost_change; o[3] = it.relative_decrease;
    o[4] = it.radius; o[5] = it.step_is_valid; o[6] = it.step_is_successful;
  }
  for (int i = 0; i < 7; ++i) para[i] = t.parameters[i];
  *termination = t.summary.termination;
}

// ---------------------------------------------------------------- whole LiDAR pipeline (CPU baseline timing)
struct OrcPipeline {
  int n_scans; double min_range;
  LaserOdometry lo;
  LaserMapping lm;
  double ms[3] = {0, 0, 0};  // accumulated SR / LO / LM wall time
  long long scans = 0;
};
void* orc_pipe_create(int n_scans, double min_range, double line_res, double plane_res) {
  auto* p = new OrcPipeline();
  p->n_scans = n_scans; p->min_range = min_range;
  p->lm.lineRes = line_res; p->lm.planeRes = plane_res;
  return p;
}
void orc_pipe_free(void* h) { delete static_cast<OrcPipeline*>(h); }
void* orc_pipe_lo(void* h) { return &static_cast<OrcPipeline*>(h)->lo; }
void* orc_pipe_lm(void* h) { return &static_cast<OrcPipeline*>(h)->lm; }
// One frame of vloam_main's LiDAR part (vloam_main_node.cpp:134,165-167): reset, scanRegistrationIO, laserOdometryIO,
// [laserMappingIO].  Returns 0, or 1 if the scan had no valid point.
int orc_pipe_process(void* h, const float* xyz, int n, int stride, int do_mapping) {
  auto* p = static_cast<OrcPipeline*>(h);
  p->lm.reset();
  double t0 = now_ms();
  ScanRegistrationOutput sr;
  scan_registration(xyz, n, stride, p->n_scans, p->min_range, &sr);
  double t1 = now_ms();
  if (sr.status) return 1;
  p->lo.solveLO(sr.laserCloud, sr.cornerPointsSharp, sr.cornerPointsLessSharp, sr.surfPointsFlat, sr.surfPointsLessFlat,
                nullptr, nullptr);
  double t2 = now_ms();
  if (do_mapping) {  // lidar_odometry_mapping.cpp:125-140: input() always, solveMapping() unless the frame is skipped
    const bool skip = p->lo.skip_frame();
    p->lm.input(p->lo.laserCloudCornerLast, p->lo.laserCloudSurfLast, p->lo.laserCloudFullRes, p->lo.q_w_curr, p->lo.t_w_curr, skip);
    if (!skip) p->lm.solveMapping();
  }
  double t3 = now_ms();
  p->ms[0] += t1 - t0; p->ms[1] += t2 - t1; p->ms[2] += t3 - t2; p->scans++;
  return 0;
}
void orc_pipe_timings(void* h, double* ms3, long long* scans) {
  auto* p = static_cast<OrcPipeline*>(h);
  ms3[0] = p->ms[0]; ms3[1] = p->ms[1]; ms3[2] = p->ms[2]; *scans = p->scans;
}

}  // extern "C"

#include "capi_vo.inc"
