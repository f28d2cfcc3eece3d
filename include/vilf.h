/*
 * vilf.h — C ABI of libvilf_cuda.so: the B200-native (sm_100a) lidar-odometry front end that replaces
 * the CPU hot path of RichExplor/VIL_Fusion (F-LOAM-derived).
 *
 * The reference has no FFI layer: the path is two header-only C++ classes instantiated in the ROS node
 * (src/visual_inertial_lidar/feature_tracker/feature_tracker_node.cpp:13-19).  Every entry point below
 * cites the reference member it stands in for.  Abbreviations (all under
 * src/visual_inertial_lidar/feature_tracker/include/):
 *   FE = featureExtraction.hpp   EM = EstimationMapping.hpp   LF = lidarFactor.hpp   CM = common.h
 *   NODE = ../feature_tracker_node.cpp   FX = featureExtract.hpp (the ring-field / range-image extractor)
 *   SC = src/global_fusion/include/Scancontext/Scancontext.h (SCManager)
 *
 * Conventions: plain pointers and sizes only; every function returns a vilf_status (0 = OK); point
 * clouds are packed float[n][4] = x, y, z, intensity (the payload of pcl::PointXYZI, CM:25); poses are
 * double[7] = qx, qy, qz, qw, tx, ty, tz (parameter_opti, EM:383).  All buffers passed in are HOST
 * memory unless a name ends in _dev.  One handle == one lidar sequence (one EstimationMapping +
 * featureExtraction pair); handles created together by vilf_create_batch share a device context and can
 * be stepped in lock-step by the *_batch calls (independent sequences, no data exchanged between them).
 * There is no CPU fallback: every call needs a CUDA device and fails with VILF_ERR_CUDA otherwise.
 */
#ifndef VILF_H_
#define VILF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vilf_handle vilf_handle;

typedef enum vilf_status {
  VILF_OK = 0,
  VILF_ERR_INVALID = 1,     /* bad argument */
  VILF_ERR_CUDA = 2,        /* CUDA runtime error (see vilf_last_error) */
  VILF_ERR_CAPACITY = 3,    /* an input or the local map exceeds the configured capacity */
  VILF_ERR_UNSUPPORTED = 4, /* e.g. a ring with more than 6*2048+10 returns */
  VILF_ERR_STATE = 5        /* call sequence error (e.g. update before map init) */
} vilf_status;

/* Parameters.  Defaults (vilf_default_config) are the reference's KITTI YAML values
 * (config/kitti/velodyne_param_64.yaml:9-23) and its hard-coded constants. */
typedef struct vilf_config {
  int32_t n_scan;          /* FE:45  /N_SCAN: 16, 32 or 64 use the reference's vertical-angle formulas
                              (FE:75-102); any other non-zero value bins everything into ring 0 like
                              FE:103-106; 0 = explicit ring ids supplied with every scan */
  int32_t n_rings;         /* number of rings when n_scan == 0 (<= 128) */
  double lidar_min;        /* FE:46  /lidarMinRange, XY range */
  double lidar_max;        /* FE:47  /lidarMaxRange */
  double edge_threshold;   /* FE:48  /edgeThreshold */
  double edge_leaf;        /* EM:82  /EdgeLeafSize */
  double surf_leaf;        /* EM:83  /SurfLeafSize */
  double crop_half;        /* EM:327-332 (100 m) */
  double knn_gate;         /* EM:129, :189 squared-distance gate on the 5th neighbour (1.0) */
  double huber;            /* EM:263 HuberLoss(0.1) */
  int32_t outer_iters;     /* EM:260 (2) */
  int32_t lm_max_iters;    /* EM:277 max_num_iterations (4) */
  int32_t max_scan_points; /* capacity: points per scan */
  int32_t max_map_points;  /* capacity: points per local map (edge and surf each) */
  int32_t max_ring_points; /* capacity: returns per ring (sizes the selection kernel's shared memory); 0 = 12298, the kernel's limit */
  int32_t flags;           /* VILF_FLAG_* bits; 0 by default */
  /* The ring-field / range-image extractor (class featureExtract, FX = include/featureExtract.hpp) replaces the ring-angle
   * extractor as stage 1 when VILF_FLAG_RANGE_IMAGE is set.  It needs n_scan == 0 (ring ids come with every scan, FX:341),
   * n_rings = its N_SCAN (FX:86); lidar_min / lidar_max are its lidarMinRange / lidarMaxRange (FX:88-89, defaults 3 / 200). */
  int32_t horizon_scan;      /* FX:85  Horizon_SCAN (1800), at most 6138 */
  int32_t downsample_rate;   /* FX:87  downsampleRate (1): only rings with ring % rate == 0 are used */
  double ri_edge_threshold;  /* FX:90  edgeThreshold (1.0), on the squared range curvature */
  double ri_surf_threshold;  /* FX:91  surfThreshold (0.1) */
} vilf_config;

/* Implementation-selection flags (results are bit-identical either way; used by the parity tests and for profiling). */
#define VILF_FLAG_NO_CLUSTER 1 /* never take the one-cluster-per-cloud kernels; always the grid-wide multi-launch path */
#define VILF_FLAG_NO_GRAPH 2   /* launch every kernel of a frame individually instead of replaying the captured CUDA graph */
/* Map storage.  Two implementations of the local maps give identical results:
 *  - cell-ordered maps (k_cellmap.cu): the map stays sorted by (search cell, voxel); createSubMap is a merge of the sorted
 *    map with the frame's few thousand new points, the search table comes out of the same pass, the 5-NN search runs one
 *    warp per query.  Cost grows with the new points, not with the map: the path for large maps.
 *  - radix-sorted maps (k_cluster.cu / k_voxel.cu + k_knn.cu): the whole map is voxel-filtered by a radix sort every frame
 *    and a hashed grid is rebuilt; one 8-CTA cluster per cloud.  Faster while a map fits one cluster (<= 2^19 points).
 * Default: cell-ordered when max_map_points + max_scan_points > 2^19, radix-sorted otherwise; the flags force one. */
#define VILF_FLAG_LEGACY_MAP 4 /* force the radix-sorted maps */
#define VILF_FLAG_CELL_MAP 8   /* force the cell-ordered maps */
#define VILF_FLAG_RANGE_IMAGE 16 /* stage 1 = featureExtract::extractFeature (FX:96-115) instead of featureExtraction::extractFeature */

int vilf_default_config(vilf_config* cfg);

/* ---- lifetime (EM:75-103 constructor/initParameter/allocateMemory, FE:43-52 initParam) ---- */
int vilf_create(const vilf_config* cfg, int device, vilf_handle** out);
/* `count` sequences sharing one device context; out[i] is sequence i. */
int vilf_create_batch(const vilf_config* cfg, int device, int count, vilf_handle** out);
int vilf_destroy(vilf_handle* h); /* destroying any handle of a batch destroys the whole batch */
const char* vilf_last_error(const vilf_handle* h);
/* Pinned host memory for scan buffers (optional; makes the H2D copy asynchronous). */
int vilf_host_alloc(void** p, uint64_t bytes);
int vilf_host_free(void* p);
/* sensor_msgs/PointCloud2 payload -> the packed float[n][4] every scan entry point takes (pcl::fromROSMsg at NODE:339-340):
 * n_points = width * height, point_step and the byte offsets of the float32 fields x, y, z, intensity from msg.fields
 * (off_intensity < 0: no such field, 0 is stored).  Host-side; write straight into a vilf_host_alloc buffer to keep the
 * upload asynchronous. */
int vilf_pack_pointcloud2(const uint8_t* data, int n_points, int point_step, int off_x, int off_y, int off_z, int off_intensity, float* xyzi_out);
/* The inverse: packed float[n][4] -> PointCloud2 data bytes, the job of pcl::toROSMsg on the /GlobalMap cloud (NODE:439-440).
 * For pcl::PointXYZI the wire layout is point_step 32, x / y / z at 0 / 4 / 8, intensity at 16; bytes no field covers are
 * zeroed.  off_intensity < 0 drops the intensity. */
int vilf_unpack_pointcloud2(const float* xyzi, int n_points, int point_step, int off_x, int off_y, int off_z, int off_intensity, uint8_t* data_out);
/* What the node publishes after optimation_processing (NODE:388-446), host-side arithmetic in Eigen's operation order:
 * q_estimator = Quaterniond(globalOdom.rotation()), t_estimator = globalOdom.translation() (NODE:388-389) from vilf_get_pose's
 * rt12 -> path_pose_out {qx,qy,qz,qw,tx,ty,tz} (the /path and tf pose, NODE:391-436); the /Odometry message carries the
 * RELATIVE pose q_last^-1 * q_estimator, q_last^-1 * (t_estimator - t_last) (NODE:400-401) -> relative_out; then
 * last <- current (NODE:445-446).  `last` is the caller's q_last / t_last, {0,0,0,1,0,0,0} after the first frame (NODE:376-377). */
int vilf_node_outputs(const double rt12[12], double last[7], double relative_out[7], double path_pose_out[7]);
/* cudaMemcpyAsync(HostToDevice) on a caller-supplied stream (bench.py measures the host link with it). */
int vilf_memcpy_h2d_async(void* dst_dev, const void* src_host, uint64_t bytes, void* cuda_stream);

/* ---- the per-frame path, as the node drives it (NODE:339-389) ---- */
/* extractFeature (FE:223-232) + first frame ? localMapInited (EM:105-115) : optimation_processing
 * (EM:235-296), all on the device; only the pose (7 doubles) comes back.  ring may be NULL unless
 * cfg.n_scan == 0.  Blocking. */
int vilf_process_scan(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, double pose_out[7]);
/* Asynchronous pair: submit enqueues H2D + the frame's kernels + D2H of the pose and returns a ticket;
 * wait blocks until that frame is done.  Up to 8 frames may be in flight; the scan buffer must stay
 * valid until its ticket has been waited for. */
int vilf_submit_scan(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, int64_t* ticket);
int vilf_wait(vilf_handle* h, int64_t ticket, double pose_out[7]);
/* Lock-step over all sequences of a batch (hs = the array vilf_create_batch filled).  When the scans of the batch lie in one host
 * array from vilf_host_alloc with a row pitch of max_scan_points points (xyzi[i] == xyzi[0] + i * max_scan_points * 4 floats; ring ids
 * likewise) they are moved with one strided copy, rows being read up to the longest scan of the batch: 10 % more of the PCIe link than a copy per scan. */
int vilf_submit_scan_batch(vilf_handle* const* hs, int count, const float* const* xyzi, const int* n,
                           const uint16_t* const* ring, int64_t* ticket);
int vilf_wait_batch(vilf_handle* const* hs, int count, int64_t ticket, double* poses_out /* [count][7] */);
/* Same step, but the scans are already resident in device memory (bench: inputs in HBM). */
int vilf_submit_scan_batch_dev(vilf_handle* const* hs, int count, const float* const* xyzi_dev, const int* n,
                               const uint16_t* const* ring_dev, int64_t* ticket);

/* ---- the reference's method surface, one call each ---- */
/* featureExtraction::extractFeature (FE:223-232).  Results stay on the device (and become the input of
 * vilf_map_init / vilf_update); counts are returned. */
int vilf_feature_extract(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, int* n_edge, int* n_surf);
/* which: 0 = edge, 1 = surf.  pts [cap][4] and src [cap] (index of each feature in the input scan) may be NULL. */
int vilf_get_features(vilf_handle* h, int which, float* pts, int32_t* src, int cap, int* n);
/* EstimationMapping::localMapInited (EM:105-115) on the last extracted features / on explicit clouds. */
int vilf_map_init(vilf_handle* h);
int vilf_map_init_points(vilf_handle* h, const float* edge, int n_edge, const float* surf, int n_surf);
/* EstimationMapping::optimation_processing (EM:235-296) on the last extracted features / on explicit clouds. */
int vilf_update(vilf_handle* h, double pose_out[7]);
int vilf_update_points(vilf_handle* h, const float* edge, int n_edge, const float* surf, int n_surf, double pose_out[7]);
/* globalOdom (EM:387): pose as quaternion + translation, and optionally the 3x3 rotation (row-major) + t. */
int vilf_get_pose(vilf_handle* h, double pose_out[7], double* rt12_or_null);
/* Overwrite parameter_opti (EM:383; q_w_c / t_w_c alias it) and, when update_odom != 0, globalOdom as EM:291-293 does. */
int vilf_set_pose(vilf_handle* h, const double pose[7], int update_odom);
/* The constant-velocity prediction that opens optimation_processing (EM:238-243): globalOdom <- globalOdom *
 * (globalOdom_last^-1 * globalOdom), globalOdom_last <- old globalOdom, parameter_opti <- predicted pose.
 * ONLY for the step-by-step path (vilf_predict, vilf_factors, vilf_solve, vilf_create_submap): vilf_update,
 * vilf_update_points and vilf_process_scan apply the same prediction themselves, so calling vilf_predict before one of
 * them predicts twice. */
int vilf_predict(vilf_handle* h, double pose_out[7]);
/* EstimationMapping::createSubMap (EM:298-352) on explicit voxel-filtered scan features at the current parameter_opti:
 * transform + append to both local maps, crop box +-crop_half about the pose, voxel filter, rebuild the search grids. */
int vilf_create_submap(vilf_handle* h, const float* edge_ds, int n_edge, const float* surf_ds, int n_surf);
/* Clouds owned by the object. which: 0 localMapEdge, 1 localMapSurf (EM:394-395), 2/3 the voxel-filtered
 * scan edge/surf features (EM:246-251), 4 cloudRegistered, 5 cloudNoRegistered (EM:391-392; getMapCloud
 * EM:365-375 returns 4+5 or 5). */
int vilf_get_cloud(vilf_handle* h, int which, float* out, int cap, int* n);

/* ---- next to the path: lidar depth for visual features ----
 * getFeatureDepth steps 4.1-4.4 (NODE:54-140): features (normalised image coordinates x, y, 1) and cloud on the unit
 * sphere, exact 3-NN, plane-ray intersection, the reference's range rules; depth_out[i] = -1 where no reliable depth
 * exists.  cloud_cam != NULL: an explicit camera-frame cloud [n][4].  cloud_cam == NULL: the scan already resident from
 * the last vilf_feature_extract / vilf_process_scan is used, after the reference's field-of-view filter and
 * pcl::transformPointCloud(LIDAR_CAMERA_EX) (NODE:348-361), T_lidar_cam = that 4x4 matrix, row-major — no second upload.
 * num_bins = 360 in the reference (NODE:51).  nn_out [m][3] (optional): indices of the three neighbours in the cloud
 * (explicit cloud) / in the scan (resident scan).  n_cloud (optional): points that entered the search. */
int vilf_feature_depth(vilf_handle* h, const float* cloud_cam, int n, const double T_lidar_cam[16], const float* feats, int m, int num_bins,
                       float* depth_out, int32_t* nn_out, int* n_cloud);

/* ---- next to the path: ScanContext place recognition ----
 * SCManager of src/global_fusion/include/Scancontext/Scancontext.h (SC below) on the GPU: the polar max-height descriptor
 * (makeScancontext SC:42-83), ring key / sector key (SC:86-115), the column-shift cosine distance (fastAlignUsingVkey,
 * distDirectSC, distanceBtnScanContext SC:119-193) and the loop-candidate search (makeAndSaveScancontextAndKeys,
 * detectLoopClosureID SC:196-299).  The ring-key kd-tree of the reference is an exact k-NN; here it is an exact exhaustive
 * search with the same metric and summation order, snapshot semantics included (the "tree" is rebuilt every
 * tree_making_period-th detection from all keys but the num_exclude_recent newest, SC:227-238).  Descriptors are
 * double[num_ring][num_sector], ROW-major (Eigen::MatrixXd is column-major: transpose when handing one to Eigen).
 * pcl::IterativeClosestPoint (poseGraphOptimization.cpp:376-444) is not part of this interface. */
typedef struct vilf_sc vilf_sc;
typedef struct vilf_sc_params {
  double lidar_height;        /* SC:313 LIDAR_HEIGHT (2.0) */
  int32_t num_ring;           /* SC:315 PC_NUM_RING (20) */
  int32_t num_sector;         /* SC:316 PC_NUM_SECTOR (60) */
  double max_radius;          /* SC:318 PC_MAX_RADIUS (80.0; setMaximumRadius SC:305) */
  int32_t num_exclude_recent; /* SC:323 (30) */
  int32_t num_candidates;     /* SC:324 NUM_CANDIDATES_FROM_TREE (3; at most 4 here) */
  double search_ratio;        /* SC:327 (0.1) */
  double dist_thres;          /* SC:329 SC_DIST_THRES (0.2; setSCdistThres SC:300) */
  int32_t tree_making_period; /* SC:332 (30) */
  int32_t max_keyframes;      /* capacity: key frames kept on the device */
  int32_t max_points;         /* capacity: points of one host cloud */
  int32_t pad_;
} vilf_sc_params;
int vilf_sc_default_params(vilf_sc_params* p);
int vilf_sc_create(const vilf_sc_params* p, int device, vilf_sc** out);
int vilf_sc_destroy(vilf_sc* sc);
const char* vilf_sc_last_error(vilf_sc* sc);
/* SCManager::makeAndSaveScancontextAndKeys (SC:196-208) of a host cloud [n][4]. */
int vilf_sc_make_and_save(vilf_sc* sc, const float* xyzi, int n);
/* The same for the cloud the odometry handle would hand out as getMapCloud(MapCloud) (EM:371-375; what the node publishes on
 * /GlobalMap and global_fusion turns into a key frame) — taken where it lies in device memory, no copy through the host. */
int vilf_sc_make_and_save_resident(vilf_sc* sc, vilf_handle* h);
/* SCManager::detectLoopClosureID (SC:210-299): loop_id = matched key frame or -1, yaw_diff_rad = deg2rad(shift * 360 / num_sector);
 * min_dist / nn_idx = the values the reference prints (10000000 / 0 when the early return of SC:220-224 is taken). */
int vilf_sc_detect_loop_closure(vilf_sc* sc, int* loop_id, float* yaw_diff_rad, double* min_dist, int* nn_idx);
/* polarcontexts_[index], polarcontext_invkeys_[index], polarcontext_vkeys_[index]; negative index counts from the back. */
int vilf_sc_get(vilf_sc* sc, int index, double* desc, double* ringkey, double* sectorkey);
int vilf_sc_size(vilf_sc* sc, int* n);
/* SCManager::distanceBtnScanContext (SC:163-193) of two explicit descriptors / of two stored key frames. */
int vilf_sc_distance(vilf_sc* sc, const double* sc1, const double* sc2, double* dist, int* shift);
int vilf_sc_distance_between(vilf_sc* sc, int i, int j, double* dist, int* shift);
int vilf_sc_launch_count(vilf_sc* sc, int64_t* launches);

/* ---- stage-level entry points (unit parity against the oracle; they clobber per-frame scratch only) ---- */
/* pcl::VoxelGrid<PointXYZI>::filter (EM:248-251, :347-350). Returns n_out; *guard = 1 when PCL's
 * "leaf size too small" int32 guard fired and the output is the input. */
int vilf_voxel_downsample(vilf_handle* h, const float* pts, int n, float leaf, float* out, int cap, int* n_out, int* guard);
/* pcl::CropBox (EM:335-344) followed by VoxelGrid, i.e. one map maintenance step on an explicit cloud. */
int vilf_crop_voxel_downsample(vilf_handle* h, const float* pts, int n, const double center[3], double half, float leaf,
                               float* out, int cap, int* n_out);
/* pcl::CropBox::filter alone (closed AABB, order preserving). */
int vilf_crop_box(vilf_handle* h, const float* pts, int n, const double mn[3], const double mx[3], float* out, int cap, int* n_out);
/* pcl::KdTreeFLANN::nearestKSearch(k=5) (EM:128, :185) of nq queries against an explicit map.  Exact for
 * every rank whose squared distance is < cfg.knn_gate (all the reference ever uses, EM:129/:189); ranks
 * beyond the gate radius may be inexact or missing (idx -1, d2 FLT_MAX). */
int vilf_knn5(vilf_handle* h, const float* map, int m, const float* q, int nq, int32_t* idx, float* d2);
/* EdgeCostFactor / SurfCostFactor (EM:117-232) at `pose` against the CURRENT local maps: per input point
 * validity, line end points a,b (6 doubles) / plane n,d (4 doubles), and the 5 neighbours. Output arrays may be NULL. */
int vilf_factors(vilf_handle* h, const double pose[7], const float* edge, int n_edge, const float* surf, int n_surf,
                 uint8_t* edge_valid, double* edge_ab, int32_t* edge_nn, float* edge_d2,
                 uint8_t* surf_valid, double* surf_nd, int32_t* surf_nn, float* surf_d2);
/* Robustified normal equations of explicit factors at `pose`: H = J^T J (upper triangle, row-major, 21),
 * g = J^T r (6), cost = 1/2 sum rho (LF:21-52, :79-102 + HuberLoss).  edge_pab: 9 doubles per factor
 * (p, a, b); surf_pnd: 7 doubles per factor (p, n, d). */
int vilf_normal_equations(vilf_handle* h, const double pose[7], const double* edge_pab, int n_edge, const double* surf_pnd,
                          int n_surf, double H21[21], double g6[6], double* cost);
/* ceres::Solve (EM:263-283) on explicit factors.  trace rows: 16 doubles = iteration, step_valid,
 * step_successful, cost, candidate_cost, model_cost_change, relative_decrease, radius, step_norm, x[7]. */
int vilf_solve(vilf_handle* h, double pose_inout[7], const double* edge_pab, int n_edge, const double* surf_pnd, int n_surf,
               int max_iters, double* trace, int max_rows, int* n_rows, int* termination);
/* Per-solve summary of the last update: rows of 8 doubles = n_edge_factors, n_surf_factors, termination,
 * n_iterations, initial cost, final cost, 0, 0. */
int vilf_get_solves(vilf_handle* h, double* out, int max_rows, int* n_rows);

/* ---- state export / import (teacher-forced parity tests, checkpoint/resume) ---- */
/* state = x[7], globalOdom (R row-major 9 + t 3), globalOdom_last (12): 31 doubles. */
int vilf_state_export(vilf_handle* h, double state31[31]);
int vilf_state_import(vilf_handle* h, const double state31[31], const float* map_edge, int n_edge, const float* map_surf, int n_surf);

/* ---- measurement ---- */
/* Per-stage device time (CUDA events on the handle's stream) accumulated since the last reset.
 * stage: 0 extract, 1 scan downsample, 2 grid build, 3 kNN + fit, 4 solve, 5 map update, 6 whole frame.
 * Profiling adds event records between stages; enable only for roofline runs. */
int vilf_profile_enable(vilf_handle* h, int on);
int vilf_profile_read(vilf_handle* h, double ms_out[7], int64_t* frames, int reset);
/* Per-kernel view of the same events: tag = phase * 64 + kernel, phase 0 extract, 1 scan downsample,
 * 2 association + solve, 3 map update, 4 grid build; n_tags must be 320.  An interval runs from the end of the
 * previous kernel to the end of this one, i.e. it includes the launch gap. */
int vilf_profile_read_kernels(vilf_handle* h, double* ms_out, int64_t* launches_out, int n_tags, int reset);
const char* vilf_profile_kernel_name(int kernel);
/* Device-resident timing of one stage on an explicit cloud (roofline sweeps; the cloud is uploaded once, then the stage
 * runs `iters` times between CUDA events on the handle's stream, after one warm-up run).
 * stage 0: spatial-hash build over `map` + 5-NN of the nq queries -> ms_out[0] = build, ms_out[1] = query (per run); the grid
 *          cell is the one a map voxel-filtered at `leaf` gets (leaf <= 0: the handle's default), ms_out[2] = shells, ms_out[3] = cell.
 * stage 1: crop box (+-100 m about the origin) + voxel filter at `leaf` of the UNSORTED cloud `map` (radix path) -> ms_out[0] per run,
 *          ms_out[2] = voxels out.
 * stage 2: the per-frame map maintenance (createSubMap, EM:298-352): `map` is voxel-filtered at `leaf` once (untimed; ms_out[1] = that
 *          time, ms_out[2] = points of the filtered map), then every run appends the nq points `q`, applies the crop box and the
 *          voxel filter and rebuilds the search structure -> ms_out[0] per run, ms_out[3] = points out.  Cell-ordered maps do this
 *          as one merge update; with VILF_FLAG_LEGACY_MAP it is the radix filter over the concatenation (search grid not included). */
int vilf_bench_stage(vilf_handle* h, int stage, const float* map, int m, const float* q, int nq, float leaf, int iters, double ms_out[4]);
/* Number of kernels this library has launched on the handle's context since creation. */
int vilf_launch_count(vilf_handle* h, int64_t* launches);
/* The CUDA stream all work of this handle is issued on (for external cudaEvent timing). */
int vilf_get_stream(vilf_handle* h, void** cuda_stream);
/* Cluster-path voxel filter of the last frame: %globaltimer (ns) at its phase boundaries (start, keys, heads, centroids,
 * grid build, end). job: 0 scan edge, 1 scan surf, 2 map edge, 3 map surf. */
int vilf_debug_voxel_phases(vilf_handle* h, int job, int64_t out8[8]);
/* Counts on the device after the last frame: n_edge, n_surf, n_ds_edge, n_ds_surf, n_map_edge, n_map_surf, status bits, frames. */
int vilf_get_counts(vilf_handle* h, int32_t out8[8]);

#ifdef __cplusplus
}
#endif
#endif /* VILF_H_ */
