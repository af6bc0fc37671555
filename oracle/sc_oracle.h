/* oracle/sc_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker; never linked into or called by the product).
 *
 * Plain-C CPU restatement of the reference's Scan Context loop-closure path
 * (SC-LeGO-LOAM/LeGO-LOAM/src/Scancontext.cpp = "SC.cpp", include/Scancontext.h = "SC.h",
 * include/nanoflann.hpp = "nf.hpp").  Pinned against the reference itself compiled verbatim
 * (oracle/_ref/libscref*.so, see oracle/Makefile) by tests/test_oracle_vs_reference.py and the
 * committed fixtures in tests/golden/.  Un-pinned boundary: Eigen 3.3's reduction order is restated
 * (oracle/stubs/Eigen/Dense), real Eigen is not installed in this image.
 */
#ifndef SC_ORACLE_H
#define SC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SC.h:77-96 */
typedef struct {
  int R;                 /* PC_NUM_RING */
  int S;                 /* PC_NUM_SECTOR */
  double lidar_height;   /* LIDAR_HEIGHT */
  double max_radius;     /* PC_MAX_RADIUS */
  int exclude_recent;    /* NUM_EXCLUDE_RECENT */
  int num_candidates;    /* NUM_CANDIDATES_FROM_TREE */
  double search_ratio;   /* SEARCH_RATIO */
  double dist_thres;     /* SC_DIST_THRES */
  int tree_period;       /* TREE_MAKING_PERIOD_ */
} sco_params;

void sco_default_params(sco_params* p);

/* glibc 2.39 atanf == fdlibm s_atanf.c in plain IEEE binary32 (no FMA); swept in tests. */
float sco_atanf(float x);
/* SC.cpp:23-36 */
float sco_xy2theta(float x, float y);
/* SC.cpp:164-179 for one point: returns 0 when the point is skipped (range > max_radius), else 1 and the
 * 1-based ring / sector of SC.cpp:178-179 plus the stored height of SC.cpp:168. */
int sco_bin_point(const sco_params* p, float x, float y, float z, int* ring, int* sector, float* height);
/* SC.cpp:151-195.  pts: n points, stride bytes apart, each starting with float x,y,z.  out: R*S doubles,
 * column-major (element (r,c) at c*R+r) like Eigen::MatrixXd. */
void sco_make_sc(const sco_params* p, const void* pts, size_t n, size_t stride, double* out);
/* SC.cpp:198-211, 214-227 */
void sco_ringkey(const sco_params* p, const double* sc, double* out_R);
void sco_sectorkey(const sco_params* p, const double* sc, double* out_S);
/* SC.cpp:93-113 */
int sco_fast_align(const sco_params* p, const double* vkey1, const double* vkey2);
/* SC.cpp:69-90 applied to (sc1, circshift(sc2, shift)) (SC.cpp:39-59) */
double sco_dist_direct_shifted(const sco_params* p, const double* sc1, const double* sc2, int shift);
/* SC.cpp:116-148 */
void sco_distance(const sco_params* p, const double* sc1, const double* sc2, double* dist, int* shift);
/* nf.hpp:383-408 */
float sco_key_dist2(const float* a, const float* b, int dim);
/* Exact K nearest ring keys by brute force in the canonical (dist asc, idx asc) order, with the result-set
 * initial state of SC.cpp:283-284 + nf.hpp:159-165 for unfilled slots (idx 0; dist 0, last slot FLT_MAX).
 * Returns min(n, K). */
int sco_knn(const sco_params* p, const float* keys, size_t n, const float* query, uint64_t* idx, float* d2);

/* The growing database + detect state of SCManager (SC.h:96-106). */
typedef struct sco_db sco_db;
sco_db* sco_db_create(const sco_params* p);
void sco_db_destroy(sco_db* db);
size_t sco_db_size(const sco_db* db);
/* SC.cpp:230-244 */
void sco_db_append_scan(sco_db* db, const void* pts, size_t n, size_t stride);
void sco_db_append_desc(sco_db* db, const double* sc);
void sco_db_get_entry(const sco_db* db, size_t i, double* sc, double* ringkey, double* sectorkey, float* ringkey_f);
/* SC.cpp:247-338 with the periodic tree snapshot of SC.cpp:264-276 (n_tree = size - exclude at the last
 * eligible call whose counter % period == 0).  cand_* (optional, K entries) receive the kNN candidates and
 * their (SC distance, shift); *n_tree the snapshot size.  Returns K, or 0 on the early return of SC.cpp:257. */
int sco_db_detect(sco_db* db, int* loop_id, float* yaw, double* min_dist, uint64_t* cand_idx, float* cand_d2,
                  double* cand_dist, int* cand_shift, uint64_t* n_tree);
/* Score query against entries [0, n): strict-min in index order (forward before column-reversed when
 * flipped != 0; the reversed pass is the reference's distanceBtnScanContext on the candidate with its
 * columns reversed -- composed, the reference has no such function). */
void sco_db_exhaustive(const sco_db* db, const double* query_sc, size_t n, int flipped, double* best_dist,
                       int* best_shift, int64_t* best_idx, int* best_flip);

/* vectorised helpers for the parity tests */
void sco_atanf_many(const float* x, size_t n, float* out_port, float* out_libm);
void sco_bin_points(const sco_params* p, const float* xyz, size_t n, int32_t* bin, float* height, float* theta);

#ifdef __cplusplus
}
#endif
#endif
