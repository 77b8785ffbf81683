/* vloam_b200.h — C ABI of the B200-native VLOAM per-scan hot path.
 *
 * Drop-in boundary for the reference's LiDAR / visual odometry classes
 * (YukunXia/VLOAM-CMU-16833).  Every entry point replaces one method the
 * reference's caller (src/vloam_main/src/vloam_main_node.cpp:125-180) reaches
 * through vloam::LidarOdometryMapping / vloam::VisualOdometry; the reference
 * interface each one replaces is cited next to it.  INTEGRATION.md shows the
 * adapter a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: opaque handles, caller-owned host buffers, int status codes
 *     (0 = OK, negative = error; never aborts, never throws across the ABI —
 *     the adapter maps missing-parameter errors to ROS_BREAK()).
 *   - a handle drives `batch` independent streams in lock-step; batch = 1 is
 *     the reference's single-sensor case.  Per-stream arrays are `batch`
 *     consecutive slabs.
 *   - device state is owned by the library; one CUDA stream per context;
 *     a handle is not thread-safe, distinct contexts are.
 *   - quaternions are (x, y, z, w) like Eigen::Quaterniond::coeffs(); clouds are
 *     pcl::PointXYZI records (x, y, z, intensity), 16 bytes.
 */
#ifndef VLOAM_B200_H
#define VLOAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vloam_ctx vloam_ctx;
typedef struct vloam_lidar vloam_lidar;
typedef struct vloam_vo vloam_vo;

enum {
  VLOAM_OK = 0,
  VLOAM_E_INVALID = -1,     /* bad argument */
  VLOAM_E_CUDA = -2,        /* CUDA runtime error, see vloam_last_error() */
  VLOAM_E_NOMEM = -3,
  VLOAM_E_STATE = -4,       /* call order violated (e.g. odometry before scan registration) */
  VLOAM_E_CAPACITY = -5     /* a scan exceeded the capacity the handle was created with */
};

/* per-stream status bits (vloam_get_stream_status) */
enum {
  VLOAM_STREAM_EMPTY = 1,          /* no point survived the NaN / minimum-range filters (the reference would crash) */
  VLOAM_STREAM_RING_OVERFLOW = 2,  /* a ring had more than 4096 points or a sector more than 1024 */
  VLOAM_STREAM_VOXEL_OVERFLOW = 4, /* pcl::VoxelGrid's "leaf size too small" path was taken (input returned unfiltered) */
  VLOAM_STREAM_CAPACITY = 8        /* a device-resident scan claimed more points than max_points / slab_points: the excess was ignored */
};

/* per-stream laser-mapping status bits (vloam_get_lm_status); 0 = the scan was mapped and inserted.  The reference's map
 * grows without bound (std::vector per cube); here it lives in a pool of map_capacity_points per stream and kind. */
enum {
  VLOAM_LM_WORKLIST_OVERFLOW = 1,  /* one scan would rewrite more than 200 cubes: its points were not inserted */
  VLOAM_LM_SCRATCH_OVERFLOW = 2,   /* the re-filter scratch would overflow: the scan's points were not inserted */
  VLOAM_LM_CORNER_MAP_FULL = 4,    /* map_capacity_points exceeded: the corner map kept its previous content */
  VLOAM_LM_SURF_MAP_FULL = 8       /* the same for the surf map */
};

/* cloud selectors for vloam_get_cloud */
enum {
  VLOAM_CLOUD_FULL = 0,        /* laserCloud            scan_registration.cpp:507 */
  VLOAM_CLOUD_SHARP = 1,       /* cornerPointsSharp     :508 */
  VLOAM_CLOUD_LESS_SHARP = 2,  /* cornerPointsLessSharp :509 */
  VLOAM_CLOUD_FLAT = 3,        /* surfPointsFlat        :510 */
  VLOAM_CLOUD_LESS_FLAT = 4,   /* surfPointsLessFlat    :511 */
  VLOAM_CLOUD_CORNER_LAST = 5, /* laserCloudCornerLast  laser_odometry.cpp:620 */
  VLOAM_CLOUD_SURF_LAST = 6,   /* laserCloudSurfLast    laser_odometry.cpp:621 */
  VLOAM_CLOUD_CORNER_STACK = 7,/* laserCloudCornerStack laser_mapping.cpp:432-435 */
  VLOAM_CLOUD_SURF_STACK = 8,  /* laserCloudSurfStack   laser_mapping.cpp:437-440 */
  VLOAM_CLOUD_CORNER_MAP = 9,  /* laserCloudCornerFromMap laser_mapping.cpp:422-428 */
  VLOAM_CLOUD_SURF_MAP = 10,   /* laserCloudSurfFromMap */
  VLOAM_CLOUD_MAP = 11,        /* /laser_cloud_map: every cube's corner then surf points, laser_mapping.cpp:778-790 */
  VLOAM_CLOUD_FULL_REGISTERED = 12 /* /velodyne_cloud_registered: laserCloudFullRes through pointAssociateToMap with the
                                      current mapping pose, laser_mapping.cpp:797-805 */
};

/* ------------------------------------------------------------------ context */
int vloam_ctx_create(int device, vloam_ctx** ctx);
int vloam_ctx_destroy(vloam_ctx* ctx);
/* Run on a caller-provided cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
int vloam_ctx_set_stream(vloam_ctx* ctx, void* cuda_stream);
int vloam_ctx_synchronize(vloam_ctx* ctx);
const char* vloam_last_error(vloam_ctx* ctx);
/* Number of kernel launches issued by this context since creation (bench.py's gpu_launches). */
long long vloam_ctx_launch_count(vloam_ctx* ctx);
/* Optional per-kernel timing: when enabled every kernel launch is bracketed by CUDA events on the launching
 * stream; vloam_ctx_get_kernel_timings synchronises and returns accumulated milliseconds / launch counts per
 * kernel id (0 .. vloam_ctx_kernel_count()-1).  Replaces the reference's TicToc prints. */
int vloam_ctx_enable_timing(vloam_ctx* ctx, int on);
int vloam_ctx_kernel_count(void);
const char* vloam_ctx_kernel_name(int id);
int vloam_ctx_get_kernel_timings(vloam_ctx* ctx, double* ms, long long* counts, int reset);

/* ------------------------------------------------------------------ LiDAR odometry + mapping
 * Parameters = the ROS parameters the reference reads
 * (src/lidar_odometry_mapping/launch/loam_velodyne_HDL_64_kitti.launch:3-16,
 *  src/vloam_main/launch/vloam_main.launch:4). */
typedef struct vloam_lidar_params {
  int batch;                       /* independent streams driven in lock-step (>= 1) */
  int max_points;                  /* capacity per scan (points) */
  int scan_line;                   /* 16 / 32 / 64         scan_registration.cpp:48-59 */
  double minimum_range;            /*                       scan_registration.cpp:51 */
  double mapping_line_resolution;  /*                       laser_mapping.cpp:95 */
  double mapping_plane_resolution; /*                       laser_mapping.cpp:97 */
  int mapping_skip_frame;          /* >= 1                  laser_odometry.cpp:53, 618-628 */
  int detach_VO_LO;                /* 1: ignore the VO prior   laser_odometry.cpp:47,223 */
  int lo_outer_passes;             /* 2                     laser_odometry.cpp:211 */
  int lo_max_iterations;           /* 4                     laser_odometry.cpp:460 */
  int lm_outer_passes;             /* 2                     laser_mapping.cpp:458 */
  int lm_max_iterations;           /* 4                     laser_mapping.cpp:612 */
  int map_capacity_points;         /* capacity of the rolling map per stream and per feature kind */
  int debug_keep_submap;           /* 1: keep a copy of laserCloudCornerFromMap / SurfFromMap (laser_mapping.cpp:422-428) of the
                                      last scan for vloam_get_cloud(VLOAM_CLOUD_*_MAP); costs one sub-map copy per scan */
  int solver_mode;                 /* how ceres::Solve is laid out on the GPU: 0 = by batch size, 1 = one CTA (or cluster) per stream,
                                      whole solve in one launch (small batches: fewest launches), 2 = wide: one launch over all
                                      residual blocks of all streams per Levenberg-Marquardt evaluation + a warp-per-stream step */
  int distortion;                  /* laser_odometry.h:90 DISTORTION (a compile-time constant of the reference, false as shipped): 1 = every
                                      feature is interpolated inside the sweep, s = frac(intensity) / 0.1, pose applied as
                                      Identity.slerp(s, q_last_curr), s * t_last_curr (laser_odometry.cpp:149-167, lidarFactor.hpp:28-35) */
} vloam_lidar_params;

/* Fills the reference's KITTI HDL-64 launch-file values. */
int vloam_lidar_params_default(vloam_lidar_params* p);

/* LidarOdometryMapping::LidarOdometryMapping() + init()   lidar_odometry_mapping.cpp:40-63 */
int vloam_lidar_create(vloam_ctx* ctx, const vloam_lidar_params* p, vloam_lidar** h);
int vloam_lidar_destroy(vloam_lidar* h);
/* LidarOdometryMapping::reset()   lidar_odometry_mapping.cpp:65-71 (call once per frame, before scan registration) */
int vloam_lidar_reset(vloam_lidar* h);

/* LidarOdometryMapping::scanRegistrationIO(cloud)   lidar_odometry_mapping.cpp:73-94
 *   -> ScanRegistration::input   scan_registration.cpp:131-449
 * xyz: host buffer, `batch` slabs of `slab_points` points, each point `stride_floats` floats starting with x, y, z
 * (3 = packed xyz, 4 = pcl::PointXYZ or a KITTI .bin record, up to 16 = a sensor_msgs/PointCloud2 payload with
 * point_step = 4 * stride_floats, taken without the pcl::fromROSMsg copy of vloam_main_node.cpp:148);
 * n_points[b] = valid points in slab b.  Pinned host memory makes the upload asynchronous. */
int vloam_scan_registration(vloam_lidar* h, const float* xyz, const int* n_points, int stride_floats, size_t slab_points);
/* Same with one host buffer per stream (xyz_ptrs[batch]: every sensor's driver owns its own message buffer; nothing is
 * gathered on the host).  n_points[b] <= max_points. */
int vloam_scan_registration_ptrs(vloam_lidar* h, const float* const* xyz_ptrs, const int* n_points, int stride_floats);
/* Same with the scans already resident in device memory (xyz_dev and n_points_dev are device pointers).  The counts
 * cannot be checked on the host: a count above min(max_points, slab_points) is clamped and reported per stream as
 * VLOAM_STREAM_CAPACITY. */
int vloam_scan_registration_device(vloam_lidar* h, const float* xyz_dev, const int* n_points_dev, int stride_floats,
                                   size_t slab_points);
/* The device copy of the scan most recently uploaded by vloam_scan_registration ([batch][slab_points][stride] floats and
 * [batch] counts), so that a second consumer on the same context — VisualOdometry::processPointCloud gets the same cloud
 * in vloam_main_node.cpp:151,164 — does not upload it again.  Valid until the second next vloam_scan_registration; call
 * vloam_input_consumed after enqueueing the extra reader so the buffer is not recycled under it. */
int vloam_get_input_device(vloam_lidar* h, const float** xyz_dev, const int** n_points_dev, int* stride_floats, size_t* slab_points);
int vloam_input_consumed(vloam_lidar* h);

/* One frame of the caller's loop in one call: LidarOdometryMapping::reset, scanRegistrationIO, laserOdometryIO and laserMappingIO
 * (vloam_main_node.cpp:134,165-167) for the given scans; results are read with vloam_get_lo_pose / vloam_get_lm_pose as usual.
 * use_graph != 0: the ~40 kernel launches of a frame are captured into a CUDA graph the first time a (buffer parity, odometry
 * initialised, mapping skipped) combination occurs and replayed afterwards — one launch per frame, which is what the latency of
 * a single stream is made of.  prior_dev: NULL or the device array vloam_laser_odometry_async takes.  The device variant copies the
 * scans into the handle's own input slot first (a graph replays fixed addresses). */
int vloam_lidar_process(vloam_lidar* h, const float* xyz, const int* n_points, int stride_floats, size_t slab_points,
                        const double* prior_dev, int use_graph);
int vloam_lidar_process_ptrs(vloam_lidar* h, const float* const* xyz_ptrs, const int* n_points, int stride_floats, const double* prior_dev,
                             int use_graph);
int vloam_lidar_process_device(vloam_lidar* h, const float* xyz_dev, const int* n_points_dev, int stride_floats, size_t slab_points,
                               const double* prior_dev, int use_graph);

/* status[batch]: VLOAM_STREAM_* bits of the last scan registration. */
int vloam_get_stream_status(vloam_lidar* h, int* status);
/* counts[batch][5]: sizes of laserCloud, sharp, lessSharp, flat, lessFlat   (ScanRegistration::output :501-512) */
int vloam_get_feature_counts(vloam_lidar* h, int* counts);
/* Copy one cloud of one stream to the host (ScanRegistration::output / LaserOdometry::output hand these out as
 * pcl::PointCloud<PointXYZI>::Ptr).  n_out receives the size; at most capacity_points records are written. */
int vloam_get_cloud(vloam_lidar* h, int stream, int which, float* xyzi_out, int capacity_points, int* n_out);
/* Parity / debug views of the reference's scratch arrays (scan_registration.h:90-93). */
int vloam_get_curvature(vloam_lidar* h, int stream, float* out, int capacity, int* n_out);
int vloam_get_labels(vloam_lidar* h, int stream, int8_t* out, int capacity, int* n_out);
/* which = VLOAM_CLOUD_SHARP / LESS_SHARP / FLAT: indices into laserCloud, in the reference's push order. */
int vloam_get_feature_indices(vloam_lidar* h, int stream, int which, int* out, int capacity, int* n_out);

/* LidarOdometryMapping::laserOdometryIO()   lidar_odometry_mapping.cpp:96-123
 *   -> LaserOdometry::input / solveLO / output   laser_odometry.cpp:135-146,187-536,610-629
 * prior: NULL, or [batch][7] = (q xyzw, t) of vloam_tf->velo_last_VOT_velo_curr, used when detach_VO_LO == 0.
 * pose_out[batch][14] = q_last_curr(4) t_last_curr(3) q_w_curr(4) t_w_curr(3); corr_out[batch][2] = corner / plane
 * correspondences of the last outer pass.  Either output pointer may be NULL. */
int vloam_laser_odometry(vloam_lidar* h, const double* prior, double* pose_out, int* corr_out);
/* Same without the blocking device->host read of the poses (they stay on the device until vloam_get_lo_pose). */
int vloam_laser_odometry_async(vloam_lidar* h, const double* prior_dev);
int vloam_get_lo_pose(vloam_lidar* h, double* pose_out, int* corr_out);
/* Pose of the PREVIOUS scan's laser odometry.  Lets a caller keep one scan in flight: enqueue scan k (upload +
 * kernels, asynchronous), then read scan k-1's result while k runs, so uploads overlap compute. */
int vloam_get_lo_pose_prev(vloam_lidar* h, double* pose_out, int* corr_out);
/* Overwrite q_last_curr / t_last_curr (the motion prior the next solve starts from), motion[batch][7]. */
int vloam_set_lo_motion(vloam_lidar* h, const double* motion);
/* Overwrite the accumulated odometry pose q_w_curr / t_w_curr (laser_odometry.cpp:80-81), pose[batch][7] = q(xyzw) t:
 * checkpoint / resume of a stream in the middle of a trajectory (the reference can only start at the origin). */
int vloam_set_lo_pose(vloam_lidar* h, const double* pose);

/* Parity read-out of one outer pass of the last laser odometry solve for one stream:
 * corr[(768 + 1536)][4] = (closestPointInd, minPointInd2, minPointInd3, valid) per query slot (corner queries
 * first), records[8][7] = cost, candidate_cost, model_cost_change, relative_decrease, radius, valid, successful;
 * info[4] = n_records, termination, n_corner, n_plane; para[7] = parameters after the pass. */
int vloam_get_lo_trace(vloam_lidar* h, int stream, int pass, int* corr, double* records, int* info, double* para);

/* LidarOdometryMapping::laserMappingIO()   lidar_odometry_mapping.cpp:125-154
 *   -> LaserMapping::input / solveMapping   laser_mapping.cpp:167-196,198-708
 * pose_out[batch][14] = q_w_curr(4) t_w_curr(3) q_wmap_wodom(4) t_wmap_wodom(3) (the /aft_mapped_to_init pose,
 * laser_mapping.cpp:720-729); NULL to skip the device->host read. */
int vloam_laser_mapping(vloam_lidar* h, double* pose_out);
int vloam_get_lm_pose(vloam_lidar* h, double* pose_out);
/* status[batch][2] = VLOAM_LM_* bits of the last mapped scan, OR of the bits of every scan since the handle was created.
 * A set bit never invalidates the poses; it says that stream's map did not take (part of) a scan. */
int vloam_get_lm_status(vloam_lidar* h, int* status);
/* Seed / read one 50 m map cube (index i + 21 j + 441 k, laser_mapping.cpp:412) of one stream; kind 0 = corner,
 * 1 = surf.  Used to pre-build the 1 M-point map of the benchmark and by the parity tests.  Seeded points must lie inside
 * the cube (laser_mapping.cpp:643-652 with the stream's current centre offsets), like every point the mapping inserts:
 * VLOAM_E_INVALID otherwise; VLOAM_E_CAPACITY when map_capacity_points is exceeded. */
int vloam_map_set_cube(vloam_lidar* h, int stream, int kind, int cube, const float* xyzi, int n);
int vloam_map_get_cube(vloam_lidar* h, int stream, int kind, int cube, float* xyzi_out, int capacity_points, int* n_out);
/* info[batch][8] = cenWidth, cenHeight, cenDepth, validNum, cornerFromMapNum, surfFromMapNum, cornerStackNum, surfStackNum */
int vloam_get_lm_info(vloam_lidar* h, int* info);
int vloam_get_lm_trace(vloam_lidar* h, int stream, int pass, double* records, int* info, double* para);
/* Work counters of laser mapping, cumulative since the handle was created (the benchmark reports their per-scan means; they
 * replace the reference's ROS_INFO statistics).  counters[batch][18]:
 *   0 scans mapped, 1 scans whose map passed the gate of laser_mapping.cpp:448, 2 / 3 Levenberg-Marquardt iterations run in
 *   outer pass 0 / 1, 4 / 5 corner / surf cubes that had to be indexed on entering the 5 x 5 x 3 window, 6 / 7 corner / surf cubes
 *   rewritten, 8 cubes re-filtered by merging (9: of those, patched inside their slab), 10 cubes re-filtered by a full voxel
 *   sort, 11 cubes outside the window that only grew, 12 voxels inserted by merges, 13 re-packs of a map pool, 14 down-sampled
 *   scan points (k-NN queries per pass), 15 residual blocks of the last pass; 16 / 17 k-NN queries / candidate points tested
 *   while vloam_lidar_set_debug_stats is on. */
int vloam_get_lm_counters(vloam_lidar* h, long long* counters);
int vloam_lidar_set_debug_stats(vloam_lidar* h, int on);
/* Parity read-out: the queries (indices into laserCloudCornerStack, kind 0, or laserCloudSurfStack, kind 1) that produced a
 * residual block in outer pass `pass` of the last scan (laser_mapping.cpp:472-581), ascending; n_out receives their number,
 * at most `capacity` are written. */
int vloam_get_lm_queries(vloam_lidar* h, int stream, int pass, int kind, int* out, int capacity, int* n_out);
/* Map storage read-out, stats[batch][2][10] per stream and feature kind (0 corner, 1 surf): points in the map, high-water
 * mark of the slab pool, pool index, non-empty cubes, cubes known to be fixed points of their voxel filter (skipped by
 * the per-scan re-filter of laser_mapping.cpp:689-702 until they receive a point), cubes rewritten by the last scan,
 * re-packs so far, slab capacity in use, points in the rewritten cubes, column-index table slots in use. */
int vloam_get_map_stats(vloam_lidar* h, int* stats);

/* ------------------------------------------------------------------ point-sharded solve (multi-GPU)
 * BASELINE configs[4] names an all-reduce of the 6x6 normal equations per Gauss-Newton iteration.  Here every rank runs
 * scan registration on the same scans, owns 1/world of each stream's correspondences (association + residuals) and the
 * laser-odometry solve kernel sums the 28 accumulators (J'J upper triangle, J'r, cost) across ranks itself, through peer
 * memory, once per LM evaluation: no host round trip and no separate collective launch.  Replaces what the reference
 * does in one thread inside ceres::Solve (laser_odometry.cpp:457-463).  Every rank gets bit-identical poses.
 *   vloam_shard_buffer / _ipc_handle : this handle's exchange buffer (device pointer / 64-byte cudaIpcMemHandle_t)
 *   vloam_shard_open_ipc             : handles[world][64] gathered from all ranks (one process per GPU)
 *   vloam_shard_enable               : the same with raw device pointers (several handles inside one process)
 * All ranks must finish enabling before any of them runs vloam_laser_odometry (barrier on the caller's side); batch <= 128.
 * vloam_shard_status: non-zero if a peer did not answer within the kernel's polling budget (results are then invalid). */
int vloam_shard_buffer(vloam_lidar* h, void** dev_ptr, size_t* bytes);
int vloam_shard_ipc_handle(vloam_lidar* h, unsigned char* handle64);
int vloam_shard_open_ipc(vloam_lidar* h, int rank, int world, const unsigned char* handles);
int vloam_shard_enable(vloam_lidar* h, int rank, int world, void* const* peer_ptrs);
int vloam_shard_disable(vloam_lidar* h);
int vloam_shard_status(vloam_lidar* h, int* error_bits);
/* The same split with the exchange done by NCCL, as BASELINE configs[4] words it: every Levenberg-Marquardt evaluation of laser
 * odometry AND laser mapping is one wide accumulate launch (tiles x streams, 28-double partial normal equations per tile),
 * an ncclAllReduce of the partials over the ranks, and a warp-per-stream step kernel; each rank associates and accumulates
 * its slice of the queries.  The library resolves NCCL at run time from the process (torch's libnccl) or VLOAM_NCCL_LIB.
 *   vloam_shard_nccl_unique_id : rank 0 creates the 128-byte ncclUniqueId, the caller broadcasts it (any transport)
 *   vloam_shard_nccl_init      : collective over the group; afterwards every laser odometry / mapping call is collective too */
int vloam_shard_nccl_unique_id(unsigned char* id128);
int vloam_shard_nccl_init(vloam_lidar* h, int rank, int world, const unsigned char* id128);
int vloam_shard_nccl_destroy(vloam_lidar* h);

/* ------------------------------------------------------------------ visual odometry (depth association + residuals)
 * VisualOdometry::setUpPointCloud   visual_odometry.cpp:132-155: cam_T_velo[16], rect0_T_cam[16], P_rect0[12], row-major float */
int vloam_vo_create(vloam_ctx* ctx, int batch, int max_points, int max_matches, vloam_vo** h);
int vloam_vo_destroy(vloam_vo* h);
int vloam_vo_set_calibration(vloam_vo* h, const float* cam_T_velo, const float* rect0_T_cam, const float* P_rect0);
/* VisualOdometry::reset()   visual_odometry.cpp:86-90 (advances the ping-pong slot) */
int vloam_vo_reset(vloam_vo* h);
/* VisualOdometry::processPointCloud   visual_odometry.cpp:157-186 -> PointCloudUtil::projectPointCloud / downsamplePointCloud
 * (point_cloud_util.cpp:148-174,205-260).  Host buffers as in vloam_scan_registration. */
int vloam_vo_process_cloud(vloam_vo* h, const float* xyz, const int* n_points, int stride_floats, size_t slab_points);
int vloam_vo_process_cloud_device(vloam_vo* h, const float* xyz_dev, const int* n_points_dev, int stride_floats, size_t slab_points);
/* PointCloudUtil::queryDepth   point_cloud_util.cpp:302-407; slot 0 = current frame, 1 = previous frame. */
int vloam_vo_query_depth(vloam_vo* h, int stream, int slot, const float* xy, int n, float* depth_out);
/* bucket grids (249 x 75, index ix * 75 + iy) of one stream: x, y, depth (float) and count (int). */
int vloam_vo_get_buckets(vloam_vo* h, int stream, int slot, float* bx, float* by, float* bd, int* bc);
/* VisualOdometry::solveNlsAll   visual_odometry.cpp:254-450.  prev_uv / curr_uv: [batch][max_matches][2] matched
 * keypoint pixels (cv::KeyPoint::pt), n_matches[batch]; init: NULL (reset_VO_to_identity) or [batch][6] = angle-axis,
 * t of cam0_curr_LOT_cam0_prev.  out[batch][8] = angles_0to1(3) t_0to1(3) counter32 counter22. */
int vloam_vo_solve(vloam_vo* h, const float* prev_uv, const float* curr_uv, const int* n_matches, const double* init,
                   int remove_VO_outlier, int max_iterations, double* out);
/* The same solve without host involvement: matches, counts and the optional initial guess already live in device memory
 * (same layouts), nothing is copied back and the call does not synchronise; read the result with vloam_vo_get_result. */
int vloam_vo_solve_device_async(vloam_vo* h, const float* prev_uv_dev, const float* curr_uv_dev, const int* n_matches_dev,
                                const double* init_dev, int remove_VO_outlier, int max_iterations);
int vloam_vo_get_result(vloam_vo* h, double* out /* [batch][8] as vloam_vo_solve */);
/* ImageUtil::detKeypoints with DetectorType::ShiTomasi   image_util.cpp:11-37 (the detector visual_odometry.cpp:34 selects):
 * cv::goodFeaturesToTrack(img, max_corners = 1024, quality_level = 0.03, min_distance = 7.5, Mat(), blockSize = 5, false, 0.04)
 * = the minimum-eigenvalue response of cv::cornerMinEigenVal(img, 5, 3), thresholded at quality_level * max, its 3 x 3 local
 * maxima sorted by response, and the greedy pass that keeps a corner when no kept corner is closer than min_distance.
 * images: host or device memory, [batch][height][width] bytes (8-bit grey, the cv::Mat rows packed).  corners_xy[batch][max_corners][2] =
 * cv::Point2f (x, y) of the corners in OpenCV's order, n_corners[batch]; both may be NULL (results stay on the device:
 * vloam_vo_get_corner_buffers).  The block size is the reference's 5.  VLOAM_E_CAPACITY: an image whose response has more
 * local maxima than a quarter of its pixels (large plateaus of exactly equal response). */
int vloam_vo_detect_corners(vloam_vo* h, const uint8_t* images, int height, int width, int max_corners, double quality_level,
                            double min_distance, float* corners_xy, int* n_corners);
/* The response map of the last detection (cv::cornerMinEigenVal), stream `stream`, row-major height x width floats. */
int vloam_vo_get_corner_response(vloam_vo* h, int stream, float* out, size_t capacity_pixels);
int vloam_vo_get_corner_buffers(vloam_vo* h, const float** corners_xy_dev, const int** n_corners_dev);
/* ImageUtil::descKeypoints with DescriptorType::ORB   image_util.cpp:162-212 (the descriptor visual_odometry.cpp:35 selects):
 * cv::ORB::create()->compute(img, keypoints, descriptors) on key points of octave 0 and the default angle -1, which is what
 * detKeypoints builds from its corners (image_util.cpp:29-35).  ORB (a) drops the key points whose rounded position is not in
 * [31, cols - 31) x [31, rows - 31) and rewrites the caller's vector to the survivors, in order; (b) blurs the image (7 x 7,
 * sigma 2; float taps, OpenCV's rounding sequence) and (c) evaluates its 256 learned intensity comparisons around each survivor.
 * images: host [batch][height][width] bytes, or NULL = the images of the last vloam_vo_detect_corners (still on the device).
 * keypoints_xy / n_keypoints: host [batch][max_matches][2] = cv::KeyPoint::pt and [batch], or both NULL = the corners of the last
 * detection (on the device; its max_corners must not exceed max_matches).  Outputs (each may be NULL; the results also stay on
 * the device as keypoints[i] / descriptors[i] of the processImage chain): kept_xy[batch][max_matches][2] the surviving key
 * points, kept_index[batch][max_matches] their positions in the input list, descriptors[batch][max_matches][32] the rows of
 * the cv::Mat, n_kept[batch]. */
int vloam_vo_describe_orb(vloam_vo* h, const uint8_t* images, int height, int width, const float* keypoints_xy, const int* n_keypoints,
                          float* kept_xy, int* kept_index, uint8_t* descriptors, int* n_kept);
/* VisualOdometry::processImage   visual_odometry.cpp:92-130 with the reference's selections (visual_odometry.cpp:34-37: ShiTomasi,
 * ORB, BF matcher, kNN selector): keypoints[i] = detKeypoints(img), descriptors[i] = descKeypoints(keypoints[i], img) and, from
 * the second frame on (count > 0), matches = matchDescriptors(descriptors[1 - i], descriptors[i]) — one upload of the images,
 * everything else on the device; the matched pixel pairs land in the buffers vloam_vo_get_match_buffers names, in solveNlsAll's
 * layout.  Call vloam_vo_reset first (visual_odometry.cpp:86-90), as the reference's frame loop does.  images:
 * [batch][height][width] bytes in host memory (pinned for an asynchronous upload) or already in device memory; n_keypoints /
 * n_matches [batch]: optional host read-outs (NULL, NULL = nothing is read back and the stream is not synchronised: the call
 * only enqueues; vloam_vo_get_detect_status then tells whether a frame overflowed the candidate list; the first frame reports
 * 0 matches).  max_matches must be >= 1024 (the detector's maxCorners). */
int vloam_vo_process_image(vloam_vo* h, const uint8_t* images, int height, int width, int* n_keypoints, int* n_matches);
int vloam_vo_get_detect_status(vloam_vo* h, int* overflowed);
/* keypoints[slot] / descriptors[slot] of that chain (slot 0 = current frame, 1 = previous frame): keypoints_xy
 * [batch][max_matches][2], descriptors [batch][max_matches][32], n_keypoints[batch]; each may be NULL. */
int vloam_vo_get_frame_features(vloam_vo* h, int slot, float* keypoints_xy, uint8_t* descriptors, int* n_keypoints);
/* The match list of the last vloam_vo_process_image / vloam_vo_match_descriptors: matches[batch][max_matches][3] =
 * (queryIdx, trainIdx, distance), n_matches[batch]. */
int vloam_vo_get_matches(vloam_vo* h, int* matches, int* n_matches);
/* ImageUtil::matchDescriptors   image_util.cpp:214-296, in the configuration VisualOdometry selects (visual_odometry.cpp:34-37):
 * cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) over 32-byte binary descriptors (cv::ORB) and the ratio test
 * `m[0].distance < ratio * m[1].distance` (ratio = 0.8, :277).  desc_query / desc_train: [batch][max_matches][32] bytes
 * (rows of the cv::Mat the extractor returns), n_query / n_train [batch] rows in use; kp_query / kp_train: NULL, or
 * [batch][max_matches][2] = cv::KeyPoint::pt of the same rows.  matches_out[batch][max_matches][3] = (queryIdx, trainIdx,
 * distance) of the accepted matches in query order, n_matches_out[batch].  Host buffers. */
int vloam_vo_match_descriptors(vloam_vo* h, const uint8_t* desc_query, const int* n_query, const uint8_t* desc_train, const int* n_train,
                               const float* kp_query, const float* kp_train, double ratio, int* matches_out, int* n_matches_out);
/* Parity read-out: knn[batch][max_matches][4] = (trainIdx of the nearest, of the second nearest, their distances) per query row:
 * the knn_matches of image_util.cpp:263 before the ratio test (-1: fewer than two train descriptors). */
int vloam_vo_get_knn(vloam_vo* h, int* knn);
/* The matched pixel pairs of that call (when keypoints were given) stay on the device in solveNlsAll's layout
 * ([batch][max_matches][2] each, counts [batch]): pass them to vloam_vo_solve_device_async. */
int vloam_vo_get_match_buffers(vloam_vo* h, const float** query_uv_dev, const float** train_uv_dev, const int** n_matches_dev);
/* ... and their host copy (parity read-out), query_uv / train_uv [batch][max_matches][2]. */
int vloam_vo_get_match_uv(vloam_vo* h, float* query_uv, float* train_uv);
/* VO result -> LO prior, on the device: cam0_curr_T_cam0_last (visual_odometry.cpp:426-430) ->
 * velo_last_VOT_velo_curr = velo_T_cam0 * cam0_curr_T_cam0_last^-1 * velo_T_cam0^-1 (VloamTF::VO2VeloAndBase, vloam_tf.cpp:59-63),
 * written as [batch][7] = q(x y z w) t to prior_dev, the layout vloam_laser_odometry_async reads (laser_odometry.cpp:225-232).
 * velo_T_cam0: row-major 4x4 (host). */
int vloam_vo_export_lo_prior(vloam_vo* h, const double* velo_T_cam0, double* prior_dev);
/* Parity read-out of the last solve of one stream: records[8][7] (as vloam_get_lo_trace; later records are dropped),
 * info[4] = n_records, termination, counter32, counter22; para[7] = final parameters (6 used). */
int vloam_vo_get_trace(vloam_vo* h, int stream, double* records, int* info, double* para);
/* Residual blocks of the last solve of one stream, per match slot: type[max_matches] = 0 none / 1 CostFunctor32 /
 * 2 CostFunctor22; obs[max_matches][5] = the functor's constructor arguments (visual_odometry.cpp:361-365, 408-412). */
int vloam_vo_get_residuals(vloam_vo* h, int stream, int* type, double* obs);

#ifdef __cplusplus
}
#endif
#endif /* VLOAM_B200_H */
