/* scgpu.h -- C ABI of libscgpu.so: the B200-native (sm_100a) Scan Context loop-closure hot path.
 *
 * This is the drop-in boundary for SC-LeGO-LOAM's SCManager.  Every entry point names the reference
 * interface it replaces ("SC.h" = SC-LeGO-LOAM/LeGO-LOAM/include/Scancontext.h, "SC.cpp" =
 * .../src/Scancontext.cpp, "mapOpt.cpp" = .../src/mapOptmization.cpp).  The C++ class the reference's
 * caller compiles against (same name, same six public methods, same constants) is include/Scancontext.h;
 * its inline methods marshal into the functions below and do nothing else.
 *
 * Conventions
 *   - plain C types, caller-owned buffers, no C++/torch types across the boundary;
 *   - every function returns SCGPU_OK (0) or a negative SCGPU_E_* code; scgpu_last_error() gives the text;
 *   - there is NO CPU fallback: without a CUDA device (or with a failing launch) calls return an error;
 *   - a handle is not thread-safe (the reference is not either: SURVEY.md 8(b)); calls may come from
 *     different host threads as long as they are serialised by the caller (mapOpt.cpp:844,1683).  Every call
 *     sets its CUDA device explicitly and uses the handle's own stream -- no thread-local CUDA state;
 *   - descriptors ("SC") cross the boundary as R*S values in COLUMN-major order (element (ring r, sector c)
 *     at c*R + r), which is the memory layout of the reference's Eigen::MatrixXd;
 *   - points cross the boundary as n records `stride` bytes apart, each beginning with float x, y, z
 *     (stride 32 = pcl::PointXYZI, 16 = float4, 12 = packed xyz).
 *
 * Results are bit-identical to the reference for bins / SC / ring key / candidate indices / shifts /
 * loop id / yaw and within 1e-5 relative for SC distances (tests/ compare against the reference compiled
 * verbatim); retrieval is exact brute force in the canonical (squared ring-key distance, index) order.
 */
#ifndef SCGPU_H
#define SCGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCGPU_OK 0
#define SCGPU_E_INVALID (-1)  /* bad argument / configuration */
#define SCGPU_E_CUDA (-2)     /* CUDA runtime error (see scgpu_last_error) */
#define SCGPU_E_NODEVICE (-3) /* no usable CUDA device: there is no CPU path */
#define SCGPU_E_EMPTY (-4)    /* detect on an empty database (reference: undefined behaviour, SC.cpp:251-252) */
#define SCGPU_E_IO (-5)       /* save / load failure */

/* flags */
#define SCGPU_FLAG_EXACT_BINNING 2u /* disable the FP32 front end of the binning kernel: every point is binned by the
                                      bit-exact restatement (results are identical either way; for A/B timing) */
#define SCGPU_FLAG_NO_SCREENING 4u /* exhaustive search scores every entry with the FP64 pair kernel (no FP32 screening
                                      pass, no second copy of the database); results are identical, for A/B timing */
#define SCGPU_FLAG_NO_TMA_BUILD 8u /* bin with the register-staged k_build instead of the TMA-staged k_build_tma (A/B) */
#define SCGPU_FLAG_PEER 16u /* this handle (shard_rank of shard_count > 1) is one shard of a PEER-SHARDED database: fixed
                                capacity (capacity_hint), one device allocation per shard that the other shards map over
                                NVLink (scgpu_peer_export / scgpu_peer_attach), ring keys replicated on every shard, candidate
                                rows fetched from their owner inside the scoring kernels.  See "peer-sharded database" below. */
#define SCGPU_FLAG_INTENSITY 32u /* intensity descriptor (the variant the reference's own comment names, SC.h:41): a bin holds the
                                    maximum point INTENSITY (the float at byte 16 of a pcl::PointXYZI record; stride >= 20)
                                    instead of z + LIDAR_HEIGHT (SC.cpp:168); everything downstream is unchanged */
#define SCGPU_FLAG_FRESH_TREE 1u /* search keys [0, size - exclude_recent) on EVERY detect instead of emulating the
                                    reference's periodically rebuilt KD-tree snapshot (SC.cpp:264-276) */

/* The reference's compile-time constants (SC.h:77-96) as run-time configuration, plus placement. */
typedef struct scgpu_config {
  int32_t num_ring;        /* PC_NUM_RING              SC.h:79  (20) */
  int32_t num_sector;      /* PC_NUM_SECTOR            SC.h:80  (60) */
  double lidar_height;     /* LIDAR_HEIGHT             SC.h:77  (2.0) */
  double max_radius;       /* PC_MAX_RADIUS            SC.h:81  (80.0) */
  int32_t exclude_recent;  /* NUM_EXCLUDE_RECENT       SC.h:86  (50) */
  int32_t num_candidates;  /* NUM_CANDIDATES_FROM_TREE SC.h:87  (10) */
  double search_ratio;     /* SEARCH_RATIO             SC.h:90  (0.1) */
  double dist_thres;       /* SC_DIST_THRES            SC.h:92  (0.5) */
  int32_t tree_period;     /* TREE_MAKING_PERIOD_      SC.h:95  (10) */
  int32_t device;          /* CUDA device ordinal */
  int32_t shard_rank;      /* this handle stores database entries i with i % shard_count == shard_rank */
  int32_t shard_count;     /* 1 = whole database on this device */
  uint64_t capacity_hint;  /* database entries (global) to reserve up front; storage grows on demand */
  uint32_t flags;          /* SCGPU_FLAG_* */
  int32_t n_devices;       /* > 1: ONE handle (one host process) drives a database sharded over devices[0 .. n_devices):
                              entry i lives on devices[i % n_devices]; every call below works on such a handle.  0 / 1: the
                              single device `device`.  (SURVEY.md 8(b) "device list".)  At most SCGPU_MAX_DEVICES. */
  int32_t devices[8];      /* CUDA ordinals; the same ordinal may appear more than once (shards sharing a device) */
} scgpu_config;
#define SCGPU_MAX_DEVICES 8

typedef struct scgpu_handle scgpu_handle;

/* Fills *cfg with the reference's constants (SC.h:77-96), device 0, one shard. */
int scgpu_default_config(scgpu_config* cfg);
/* SCManager() (SC.h:61) + the storage behind SC.h:99-106, resident in HBM. */
int scgpu_create(const scgpu_config* cfg, scgpu_handle** out);
int scgpu_destroy(scgpu_handle* h);
const char* scgpu_last_error(void);
/* Library / build identification (sm target, kernels compiled). */
const char* scgpu_version(void);

/* ---- the reference's public methods, one call each --------------------------------------------------- */

/* SCManager::makeScancontext (SC.h:63, SC.cpp:151-195): out = R*S doubles, column-major. */
int scgpu_make_sc(scgpu_handle* h, const void* pts, size_t n, size_t stride_bytes, double* out_sc);
/* SCManager::makeRingkeyFromScancontext (SC.h:64, SC.cpp:198-211): row means, out = R doubles. */
int scgpu_ringkey(scgpu_handle* h, const double* sc, double* out_ring);
/* SCManager::makeSectorkeyFromScancontext (SC.h:65, SC.cpp:214-227): column means, out = S doubles. */
int scgpu_sectorkey(scgpu_handle* h, const double* sc, double* out_sector);
/* SCManager::fastAlignUsingVkey (SC.h:67, SC.cpp:93-113): vkey1, vkey2 = S doubles. */
int scgpu_fast_align(scgpu_handle* h, const double* vkey1, const double* vkey2, int* out_shift);
/* SCManager::distDirectSC (SC.h:68, SC.cpp:69-90). */
int scgpu_dist_direct(scgpu_handle* h, const double* sc1, const double* sc2, double* out_dist);
/* SCManager::distanceBtnScanContext (SC.h:69, SC.cpp:116-148). */
int scgpu_distance(scgpu_handle* h, const double* sc1, const double* sc2, double* out_dist, int* out_shift);
/* SCManager::makeAndSaveScancontextAndKeys (SC.h:72, SC.cpp:230-244; call site mapOpt.cpp:1630).
 * Asynchronous: returns once the scan has been staged; the scan buffer is not retained. */
int scgpu_append_scan(scgpu_handle* h, const void* pts, size_t n, size_t stride_bytes);
/* SCManager::detectLoopClosureID (SC.h:73, SC.cpp:247-338; call site mapOpt.cpp:916).
 * *loop_id / *yaw_rad are the pair the reference returns.  The optional outputs carry what the reference
 * prints (SC.cpp:322-329): the nearest distance, nearest index and its shift. */
int scgpu_detect(scgpu_handle* h, int* loop_id, float* yaw_rad, double* nearest_dist, int* nearest_idx,
                 int* nearest_shift);
/* xy2theta (SC.h:53, SC.cpp:23-36) evaluated on the current device: azimuth in degrees. */
int scgpu_xy2theta(float x, float y, float* out_deg);
/* polarcontexts_.size() (SC.h:100). */
int scgpu_size(scgpu_handle* h, uint64_t* out_n);

/* ---- batched / bench / parity extras (no reference counterpart: the reference is one scan at a time) ---- */

/* n_scans scans of pts_per_scan points each, contiguous, appended in order (= n_scans calls of
 * scgpu_append_scan).  location: 0 = host memory (pinned or pageable), 1 = device memory on cfg.device. */
int scgpu_append_scans_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan,
                               size_t stride_bytes, int location);
/* Appends n ready-made descriptors (float, column-major R*S each; host memory): the keys are derived on
 * the device exactly as SC.cpp:233-235 would.  Used to load / pre-fill a database. */
int scgpu_append_descs(scgpu_handle* h, const float* sc, size_t n);
/* The bench "step": append n_scans scans (as above) and run one detect after each append, exactly as the
 * sequence { makeAndSaveScancontextAndKeys; detectLoopClosureID } x n_scans would, including the periodic
 * tree-snapshot state.  Outputs are host arrays of n_scans entries (nearest_* optional). */
int scgpu_replay_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan,
                         size_t stride_bytes, int location, int* loop_id, float* yaw_rad, double* nearest_dist,
                         int* nearest_idx, int* nearest_shift);
/* Asynchronous form of scgpu_replay_batched: the whole step is enqueued and the call returns; scgpu_replay_results waits for
 * the LAST enqueued step and copies its results out.  Steps may be enqueued back to back: the binning of a step overlaps the
 * query stage of the one before (two streams; its append waits until those queries are done).  Host scans (location 0) are
 * consumed before the call returns only as far as staging goes -- keep them valid until scgpu_replay_results. */
int scgpu_replay_async(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride_bytes,
                       int location);
int scgpu_replay_results(scgpu_handle* h, size_t n, int* loop_id, float* yaw_rad, double* nearest_dist, int* nearest_idx,
                         int* nearest_shift);
/* Detect for n_queries already-stored entries [first, first+n_queries) as if each had just been appended
 * (database truncated to first+i+1 for query i) with a FRESH snapshot (n_search = first+i+1-exclude_recent). */
int scgpu_query_batched(scgpu_handle* h, uint64_t first, size_t n_queries, int* loop_id, float* yaw_rad,
                        double* nearest_dist, int* nearest_idx, int* nearest_shift);
/* What the last scgpu_detect saw (parity dumps): K candidates in retrieval order, their squared ring-key
 * distances, per-candidate SC distance and shift, and the number of keys that were searchable. */
int scgpu_get_candidates(scgpu_handle* h, uint64_t* cand_idx, float* cand_d2, double* cand_dist, int* cand_shift,
                         uint64_t* n_search);
/* Same for query q of the last scgpu_replay_batched / scgpu_query_batched call. */
int scgpu_get_batch_candidates(scgpu_handle* h, size_t q, uint64_t* cand_idx, float* cand_d2, double* cand_dist,
                               int* cand_shift, uint64_t* n_search);
/* Entry i as stored: SC (float R*S), ring key (float R = polarcontext_invkeys_mat_), sector key (double S). */
int scgpu_get_entry(scgpu_handle* h, uint64_t i, float* sc, float* ring, double* sector);
/* Forget entries >= n (bench resets; never called by the reference) and reset the tree-snapshot state, so the
 * next detect takes a fresh snapshot exactly like the first detect of a new SCManager. */
int scgpu_truncate(scgpu_handle* h, uint64_t n);
/* Exhaustive search (BASELINE config 4/5): score query entry q against EVERY entry [0, n_search) with
 * distanceBtnScanContext semantics; strict-min in index order.  flipped != 0 also scores each candidate with
 * its columns reversed (forward first).  Outputs are for the global winner. */
int scgpu_exhaustive(scgpu_handle* h, uint64_t q, uint64_t n_search, int flipped, double* best_dist,
                     int* best_shift, int64_t* best_idx, int* best_flip);
/* nq exhaustive searches (query entry q[i] against entries [0, n_search[i])) enqueued back to back on the device
 * with one host synchronisation at the end; results as scgpu_exhaustive (forward search). */
int scgpu_exhaustive_batched(scgpu_handle* h, const uint64_t* q, const uint64_t* n_search, size_t nq, double* best_dist,
                             int* best_shift, int64_t* best_idx);
/* How many entries the last scgpu_exhaustive had to rescore with the exact FP64 kernel (the rest was ruled out
 * by the FP32 screening pass with a proven margin). */
int scgpu_exhaustive_stats(scgpu_handle* h, uint64_t* rescored);
/* Parity probe of the exhaustive search's screening pass (FP32 SIMT kernels for the windowed search, the tcgen05 tensor-core
 * kernel for the full-shift search): the screened distance of stored entry q against every entry [0, n_search) --
 * d32[i] approximates distanceBtnScanContext(q, i) within the selection margin; -1 = "cannot tell, the exact kernel decides",
 * +inf = no valid column pair.  shift (optional, full-shift configuration only): the argmin shift per entry. */
int scgpu_probe_screen(scgpu_handle* h, uint64_t q, uint64_t n_search, float* d32, uint32_t* shift);
/* ---- voxel-grid downsample in front of the path ---------------------------------------
 * The reference filters every raw scan with pcl::VoxelGrid before it reaches SCManager:
 * downSizeFilterScancontext.setLeafSize(0.5, 0.5, 0.5) (mapOpt.cpp:264), .filter() (mapOpt.cpp:1235-1237), and the
 * filtered cloud is what makeAndSaveScancontextAndKeys receives (mapOpt.cpp:1628-1630). */

/* = downSizeFilterScancontext.setLeafSize(leaf, leaf, leaf).  leaf > 0: every scan given to scgpu_append_scan /
 * scgpu_append_scans_batched / scgpu_replay_batched / scgpu_make_sc / scgpu_stage_build is first reduced to its voxel
 * centroids on the device (k_build_voxel: one thread-block cluster per scan, voxel table in distributed shared
 * memory), i.e. the caller passes the RAW scan and drops its own VoxelGrid.  leaf = 0 (default): scans are binned as
 * given.  Leaf indices are bit-identical to PCL's; a centroid is the correctly rounded mean of its points (PCL: FP32
 * running sum in std::sort order), so coordinates agree to a few ulp. */
int scgpu_set_downsample_leaf(scgpu_handle* h, float leaf);
/* = pcl::VoxelGrid<PointXYZI>::filter on one host scan: out_xyzn[i] = {centroid x, y, z, number of points} and
 * out_idx[i] = PCL's leaf index of voxel i (order unspecified; PCL emits ascending leaf index), *out_n = number of
 * voxels (an error if > cap).  min_b / div_b (3 ints each, optional): PCL's grid origin and extent.  *status
 * (optional): bit 0 = PCL's "leaf size too small" refusal (the input is passed through, out_idx = point index),
 * bits 8.. = number of key partitions the scan needed.  Intensity is not carried (the path reads x, y, z only). */
int scgpu_voxel_downsample(scgpu_handle* h, const void* pts, size_t n, size_t stride_bytes, float leaf, float* out_xyzn,
                           uint32_t* out_idx, size_t cap, size_t* out_n, int32_t* min_b, int32_t* div_b, int32_t* status);
/* ---- loop verification after the path ---------------------------------------------------------------
 * mapOptmization.cpp:1053-1078: once Scan Context has named a loop candidate, the latest keyframe cloud is aligned to the
 * history submap (+-25 keyframes around the candidate, mapOpt.cpp:928-949) with pcl::IterativeClosestPoint and the loop is
 * accepted when the ICP converged with fitness <= historyKeyframeFitnessScore (utility.h:139).  scgpu_verify_loop runs that
 * point-to-point ICP on the device (exact brute-force nearest neighbours, FP64 transform estimation; csrc/scgpu_icp.cuh).
 * PARITY UNPINNED: PCL is absent from the reference tree and from this image; the kernels and oracle/icp_oracle.cpp restate
 * PCL 1.8's published algorithm and defaults. */
typedef struct scgpu_icp_params {
  int32_t max_iterations;             /* icp.setMaximumIterations           (100)  mapOpt.cpp:1055 */
  int32_t seed_axis;                  /* -1: identity initial guess (the reference, mapOpt.cpp:1066); 0/1/2: seed with a rotation of
                                         seed_angle about x / y / z -- the yaw detectLoopClosureID returned (mapOpt.cpp:918, unused there) */
  double max_correspondence_distance; /* icp.setMaxCorrespondenceDistance   (100)  mapOpt.cpp:1054 */
  double transformation_epsilon;      /* icp.setTransformationEpsilon       (1e-6) mapOpt.cpp:1056 */
  double euclidean_fitness_epsilon;   /* icp.setEuclideanFitnessEpsilon     (1e-6) mapOpt.cpp:1057 */
  double fitness_threshold;           /* historyKeyframeFitnessScore        (1.5)  utility.h:139   */
  float seed_angle;                   /* radians */
  float reserved;
} scgpu_icp_params;
int scgpu_default_icp_params(scgpu_icp_params* p);
/* src / tgt: host clouds (n points, `stride` bytes apart, x y z first) = icp.setInputSource / setInputTarget.  Outputs (each
 * optional): T = icp.getFinalTransformation() (row-major 4x4, source -> target), fitness = icp.getFitnessScore(), converged =
 * icp.hasConverged(), iterations run, accepted = converged && fitness <= fitness_threshold (the test of mapOpt.cpp:1068). */
int scgpu_verify_loop(scgpu_handle* h, const void* src, size_t n_src, const void* tgt, size_t n_tgt, size_t stride_bytes,
                      const scgpu_icp_params* prm, double* T_row_major_16, double* fitness, int* converged, int* iterations,
                      int* accepted);
/* ---- submap assembly in front of the ICP (mapOptmization.cpp:928-949) ---------------------------------------------------------
 * The reference builds both ICP inputs from STORED keyframe clouds: each cloud is moved by a 6-DoF key pose with LeGO-LOAM's
 * transformPointCloud (mapOptmization.cpp:598-627: yaw about z, roll about x, pitch about y, translation; FP32), the clouds are
 * concatenated, the query side drops points with (int)intensity < 0 (932-939) and the history side goes through a voxel grid
 * (downSizeFilterHistoryKeyFrames, leaf 0.3 m, 264-268 / 948-949).  scgpu_pose6 = the fields of PointTypePose that function reads.
 * The transformed coordinates are bit-identical to the reference's (same FP32 operations; cosf / sinf taken on the host);
 * the voxel grid is k_build_voxel (PARITY UNPINNED, as above).
 *
 * scgpu_assemble_submap: clouds[i] (n_points[i] points, `stride` bytes apart, x y z first, intensity at byte `intensity_off`, 0 =
 * none) moved by poses[i], cloud after cloud (the reference's operator+=); drop_negative_intensity applies 932-939; leaf > 0 applies
 * the voxel grid to the union.  out_xyzw[i] = {x, y, z, intensity} (leaf = 0) or {centroid x, y, z, number of points} in unspecified
 * order (leaf > 0); *n_out = points (an error if > cap).
 *
 * scgpu_verify_loop_keyframes = mapOptmization.cpp:924-949 + 1053-1078 in one call, nothing returns to the host in between: the
 * query clouds (corner + surface cloud of the latest keyframe) moved by ONE pose -- that of the matched keyframe -- with negative
 * intensities dropped when intensity_off != 0; the history clouds (corner + surface clouds of the keyframes around the match) each by
 * its own pose, then the voxel grid (leaf; 0 = none); then the ICP of scgpu_verify_loop.  n_src_used / n_tgt_used (optional): points
 * that entered the ICP. */
typedef struct scgpu_pose6 {
  float x, y, z, roll, pitch, yaw;
} scgpu_pose6;
int scgpu_assemble_submap(scgpu_handle* h, const void* const* clouds, const size_t* n_points, const scgpu_pose6* poses, size_t n_clouds,
                          size_t stride_bytes, size_t intensity_off, int drop_negative_intensity, float leaf, float* out_xyzw, size_t cap,
                          size_t* n_out);
int scgpu_verify_loop_keyframes(scgpu_handle* h, const void* const* src_clouds, const size_t* src_points, size_t n_src_clouds,
                                const scgpu_pose6* src_pose, const void* const* tgt_clouds, const size_t* tgt_points,
                                const scgpu_pose6* tgt_poses, size_t n_tgt_clouds, size_t stride_bytes, size_t intensity_off, float leaf,
                                const scgpu_icp_params* prm, double* T_row_major_16, double* fitness, int* converged, int* iterations,
                                int* accepted, size_t* n_src_used, size_t* n_tgt_used);
/* Flat binary save / load of the descriptor database. */
int scgpu_save(scgpu_handle* h, const char* path);
int scgpu_load(scgpu_handle* h, const char* path);

/* ---- staged device-side API (used by the sharded multi-GPU path; all pointers are DEVICE pointers on
 *      cfg.device; `stream` is the caller's cudaStream_t, used as given -- NULL is the CUDA default stream --
 *      so the launches are ordered with the caller's own work, e.g. torch's current stream and NCCL) --------- */

/* Bytes of one packed descriptor record: float sc[R*S] | float ring[R] | double sector[S] | double colnorm[S],
 * padded to 16 bytes. */
int scgpu_record_bytes(scgpu_handle* h, size_t* out);
/* Stage 1+2: scans -> records (descriptor + keys). */
int scgpu_stage_build(scgpu_handle* h, const void* d_pts, size_t n_scans, size_t pts_per_scan, size_t stride_bytes,
                      void* d_records, void* stream);
/* Store records [0, n) as global entries first_global + i*global_step; only those with index % shard_count ==
 * shard_rank are kept by this handle (global_step = shard_count, first_global % shard_count == shard_rank:
 * a rank storing the entries it built itself).  Global size becomes max(size, first_global + (n-1)*global_step + 1). */
int scgpu_stage_append(scgpu_handle* h, const void* d_records, uint64_t first_global, uint64_t global_step, size_t n,
                       void* stream);
/* Declares the global database size (all shards) after appends made through other handles / ranks. */
int scgpu_stage_set_size(scgpu_handle* h, uint64_t n_global);
/* Stage 3 (local): for each query record, the K best (dist2 bits << 32 | global idx) keys among this shard's
 * entries with global idx < d_n_search[q], ascending; unfilled slots = UINT64_MAX. */
int scgpu_stage_topk(scgpu_handle* h, const void* d_query_records, size_t n_queries, const uint64_t* d_n_search,
                     uint64_t* d_keys_out, void* stream);
/* Merge `parts` key lists per query (layout [parts][n_queries][K]) into one [n_queries][K]. */
int scgpu_stage_merge(scgpu_handle* h, const uint64_t* d_keys_parts, int parts, size_t n_queries,
                      uint64_t* d_keys_out, void* stream);
/* Stage 4 (local): score the candidates this shard owns; per query the shard's best in retrieval order as
 * {double dist; int32 rank_in_list; int32 shift; int64 global_idx} (24 bytes), dist = +inf when none. */
int scgpu_stage_score(scgpu_handle* h, const void* d_query_records, size_t n_queries, const uint64_t* d_keys,
                      const uint64_t* d_n_search, void* d_best_out, void* stream);
/* The packed record of stored entry global_idx (which must live on this shard) -> d_record. */
int scgpu_stage_gather(scgpu_handle* h, uint64_t global_idx, void* d_record, void* stream);
/* Exhaustive search of this shard for nq query records (host array n_search[nq]): per query
 * {double dist; int32 n_rescored; int32 shift; int64 global_idx} of the shard's strict-min winner (dist = 1e7 when
 * none) -> d_best_out[nq].  The global winner is the minimum over shards by (dist, global_idx).  Queries are screened
 * in batches of up to 64 per launch so the small per-query kernels are shared. */
int scgpu_stage_exhaustive(scgpu_handle* h, const void* d_query_records, size_t nq, const uint64_t* n_search, void* d_best_out,
                           void* stream);
/* Reduce `parts` per-shard bests (layout [parts][n_queries]) to the reference's result per query
 * (d_n_search[q] == 0 marks a query that took the early return of SC.cpp:257-261). */
int scgpu_stage_finalize(scgpu_handle* h, const void* d_best_parts, int parts, size_t n_queries,
                         const uint64_t* d_n_search, int32_t* d_loop_id, float* d_yaw, double* d_nearest_dist,
                         int32_t* d_nearest_idx, int32_t* d_nearest_shift, void* stream);
/* The same with the column-reversed ("flipped") pass of BASELINE config 5 (flipped != 0: every entry is also scored with its
 * columns reversed, forward first); a flipped winner is flagged in bit 30 of the shift field. */
int scgpu_stage_exhaustive2(scgpu_handle* h, const void* d_query_records, size_t nq, const uint64_t* n_search, int flipped,
                            void* d_best_out, void* stream);
/* Fallback of scgpu_stage_exhaustive: when a result's n_rescored exceeds SCGPU_EXH_LIST_CAP the rescoring list of its batch
 * overflowed (a database of near-duplicates: the list is shared by the <= 64 queries of a batch) and the reported winner is
 * NOT reliable; this call scores every local entry of the shard exactly for that one query record (host-synchronous) and
 * writes the shard's true winner to d_best_out (same 24-byte layout). */
#define SCGPU_EXH_LIST_CAP 65536
int scgpu_stage_exhaustive_exact(scgpu_handle* h, const void* d_query_record, uint64_t n_search, void* d_best_out);

/* ---- peer-sharded database: the shards read and write each other's memory over NVLink ---------------------------------
 * (replaces the three all_gathers of the staged API above.)  Entry i lives on shard i % G.  Every shard keeps
 * a replica of ALL ring keys (80 B per entry: k_append stores a new entry's key into every replica with peer stores), so the
 * retrieval for a query runs on ONE device with no merge; the queries are partitioned -- the shard that binned a scan also
 * searches for it -- and the candidates' rows are fetched from their owner shards inside the scoring kernels (the per-warp TMA
 * bulk copy of k_cand_screen and the loads of the FP64 pair kernel address peer memory).  Per-query work divides by G.
 *
 *   one process, several GPUs : scgpu_config.n_devices / devices -- scgpu_create returns ONE handle for the whole database and
 *                               every call of this header works on it (cudaDeviceEnablePeerAccess, events between the streams);
 *   one process per GPU       : every rank creates its shard with SCGPU_FLAG_PEER (same configuration, capacity_hint = the
 *                               fixed global capacity), exchanges scgpu_peer_export blobs (e.g. torch.distributed all_gather)
 *                               and calls scgpu_peer_attach (cudaIpcOpenMemHandle); steps are then collective calls of
 *                               scgpu_peer_replay_async, synchronised by in-kernel flag barriers over peer memory -- no NCCL
 *                               call on the data path.  One GPU per process (a cross-process barrier cannot share a GPU). */
/* The partition of a batch over the shards (host arithmetic, no device needed): shard `rank` of G handles the scans
 * first_index, first_index + G, ... (count of them) of a batch whose first scan becomes global entry `first`. */
int scgpu_peer_partition(uint64_t first, size_t n_total, int G, int rank, size_t* first_index, size_t* count);
#define SCGPU_PEER_BLOB_BYTES 128
int scgpu_peer_export(scgpu_handle* h, void* blob, size_t blob_bytes);
/* blobs: n * SCGPU_PEER_BLOB_BYTES bytes, blob s exported by shard s (this shard's own included). */
int scgpu_peer_attach(scgpu_handle* h, const void* blobs, int n);
/* COLLECTIVE replay step: a batch of n_total scans is appended and a detect is run after each, exactly as
 * scgpu_replay_batched would on one device; scan i becomes global entry size + i and belongs to shard (size + i) % G.  Every
 * rank passes ITS scans of the batch (those i with (size + i) % G == shard_rank, in order, contiguous; host or device memory)
 * and the same n_total.  Asynchronous; scgpu_replay_results then returns the results of ALL n_total scans on every rank. */
int scgpu_peer_replay_async(scgpu_handle* h, const void* pts, size_t n_total, size_t pts_per_scan, size_t stride_bytes,
                            int location);
/* The cross-process barrier by itself on `stream` (all ranks call it the same number of times). */
int scgpu_peer_barrier(scgpu_handle* h, void* stream);

/* Host helper: the n_search sequence a run of n consecutive detects produces (one detect after each append,
 * database size first_size, first_size+1, ...), advancing the handle's snapshot state (SC.cpp:257-276). */
int scgpu_plan_n_search(scgpu_handle* h, uint64_t first_size, size_t n, uint64_t* out_n_search);
/* Parity probes (host arrays in, host arrays out): the device's atanf restatement for n values, and for n points
 * (xyz packed, 12 bytes each) the 0-based bin sector*R+ring of SC.cpp:178-179 (-1 = not binned), the stored height
 * of SC.cpp:168 and the azimuth of SC.cpp:172. */
int scgpu_probe_atanf(const float* x, size_t n, float* out);
int scgpu_probe_bins(scgpu_handle* h, const float* xyz, size_t n, int32_t* bin, float* height, float* theta_deg);
/* Device time (CUDA events on the handle's stream) of the last scgpu_replay_batched / scgpu_append_scans_batched /
 * scgpu_query_batched call: whole call, its k_build launches only, and everything after them. */
int scgpu_get_timing(scgpu_handle* h, double* ms_total, double* ms_build, double* ms_query);
/* Host side of the H2D path (no device needed): threads of the packing pool (min(16, cores / LOCAL_WORLD_SIZE), or
 * $SCGPU_HOST_THREADS) and whether pinned batches are packed to 12 bytes per point before they cross PCIe (a lone process -- LOCAL_WORLD_SIZE 1 -- with a pool of >= 12
 * threads, or $SCGPU_PACK_PINNED).  Pageable sources are always packed (they are staged through pinned memory anyway). */
int scgpu_host_info(int* pool_threads, int* packs_pinned);
/* Growth of the (non-peer) shard beyond its capacity: the arrays live behind address ranges reserved once
 * (cuMemAddressReserve) and grow by mapping more physical memory (cuMemCreate / cuMemMap) -- the pointers never change, so
 * nothing is synchronised or copied and work already enqueued is unaffected.  Where the driver offers no virtual memory
 * management (or with SCGPU_NO_VMM=1) the shard is reallocated and copied behind a device synchronisation, as in round 1.
 * Counts of both kinds of growth so far and the current local capacity (any pointer may be null). */
int scgpu_growth_stats(scgpu_handle* h, unsigned* in_place, unsigned* by_copy, uint64_t* capacity);
/* Device-side stopwatch over a sequence of calls (asynchronous ones included): CUDA events on the handle's own streams --
 * start behind everything enqueued so far, stop behind everything enqueued since; a device-list handle reports its slowest
 * shard.  scgpu_timer_stop waits for the work to finish. */
int scgpu_timer_start(scgpu_handle* h);
int scgpu_timer_stop(scgpu_handle* h, double* ms);
/* On-device self check of the binning front end: n pseudo-random points (mode 0: uniform over the ROI square;
 * mode 1: on / next to ring and sector boundaries), fast-path+fallback bin vs the exact restatement.
 * first_bad (optional, 5 floats): x, y, z, fast bin, exact bin of the first mismatch. */
int scgpu_probe_selfcheck(scgpu_handle* h, uint64_t n, uint64_t seed, int mode, uint64_t* mismatches,
                          uint64_t* fallbacks, float* first_bad);
/* Number of kernels launched by this handle so far (bench.py's gpu_launches). */
int scgpu_launch_count(scgpu_handle* h, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SCGPU_H */
