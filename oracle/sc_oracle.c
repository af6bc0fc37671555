/* oracle/sc_oracle.c -- TEST INFRASTRUCTURE ONLY (the checker; never linked into or called by the product).
 *
 * Plain-C restatement of the reference's Scan Context path.  See sc_oracle.h for the citation key.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -march, no fast-math), like the reference build
 * (LeGO-LOAM/CMakeLists.txt:4-5): every float/double operation below is one IEEE operation.
 */
#define _GNU_SOURCE /* M_PI */
#include "sc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void sco_default_params(sco_params* p) {
  p->R = 20;
  p->S = 60;
  p->lidar_height = 2.0;
  p->max_radius = 80.0;
  p->exclude_recent = 50;
  p->num_candidates = 10;
  p->search_ratio = 0.1;
  p->dist_thres = 0.5;
  p->tree_period = 10;
}

/* ---------------------------------------------------------------------------------------------
 * atanf.  The reference calls the C library (SC.cpp:26-35 -> atanf because <math.h> is in scope in the
 * ROS build).  glibc 2.39's binary32 atanf is the classic fdlibm algorithm (argument reduction to
 * [0,7/16], odd/even split degree-11 polynomial, hi/lo table) evaluated in plain binary32; it is NOT
 * correctly rounded, so neither (float)atan(double) nor CUDA's atanf can stand in for it.
 * ------------------------------------------------------------------------------------------- */
static float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

float sco_atanf(float x) {
  static const uint32_t HI[4] = {0x3eed6338u, 0x3f490fdau, 0x3f7b985eu, 0x3fc90fdau};
  static const uint32_t LO[4] = {0x31ac3769u, 0x33222168u, 0x33140fb4u, 0x33a22168u};
  static const uint32_t AT[11] = {0x3eaaaaabu, 0xbe4ccccdu, 0x3e124925u, 0xbde38e38u, 0x3dba2e6eu, 0xbd9d8795u,
                                  0x3d886b35u, 0xbd6ef16bu, 0x3d4bda59u, 0xbd15a221u, 0x3c8569d7u};
  const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
  int id;
  if (ix >= 0x4c000000u) { /* |x| >= 2^25, inf, NaN */
    if (ix > 0x7f800000u) return x + x;
    float r = u2f(HI[3]) + u2f(LO[3]);
    return (hx >> 31) ? -r : r;
  }
  if (ix < 0x3ee00000u) { /* |x| < 7/16 */
    if (ix < 0x31000000u) return x; /* |x| < 2^-29 */
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000u) {   /* |x| < 19/16 */
      if (ix < 0x3f300000u) { /* 7/16 <= |x| < 11/16 */
        id = 0;
        x = (2.0f * x - 1.0f) / (2.0f + x);
      } else {
        id = 1;
        x = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000u) { /* |x| < 39/16 */
        id = 2;
        x = (x - 1.5f) / (1.0f + 1.5f * x);
      } else {
        id = 3;
        x = -1.0f / x;
      }
    }
  }
  float aT[11];
  for (int i = 0; i < 11; ++i) aT[i] = u2f(AT[i]);
  const float z = x * x;
  const float w = z * z;
  const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = u2f(HI[id]) - ((x * (s1 + s2) - u2f(LO[id])) - x);
  return (hx >> 31) ? -r : r;
}

/* SC.cpp:23-36.  (180/M_PI) is an int/double expression = a double constant; the float atanf result is
 * widened, scaled and offset in double, and the double is narrowed to float by the return.  The four tests
 * use >= / < exactly as written, so -0.0 counts as ">= 0".  With a NaN coordinate the reference falls off
 * the end of the function (undefined); the restatement returns NaN. */
float sco_xy2theta(float x, float y) {
  const double k = 180 / M_PI;
  if ((x >= 0) & (y >= 0)) return (float)(k * (double)sco_atanf(y / x));
  if ((x < 0) & (y >= 0)) return (float)(180 - (k * (double)sco_atanf(y / (-x))));
  if ((x < 0) & (y < 0)) return (float)(180 + (k * (double)sco_atanf(y / x)));
  if ((x >= 0) & (y < 0)) return (float)(360 - (k * (double)sco_atanf((-y) / x)));
  return NAN;
}

/* int(double) as the x86-64 build does it (cvttsd2si): NaN and out-of-range give INT_MIN. */
static int d2i_x86(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return (-2147483647 - 1);
  return (int)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* SC.cpp:166-179 */
int sco_bin_point(const sco_params* p, float x, float y, float z, int* ring, int* sector, float* height) {
  const float h = (float)((double)z + p->lidar_height);      /* SC.cpp:168: float = float + double */
  const float azim_range = sqrtf(x * x + y * y);             /* SC.cpp:171 */
  const float azim_angle = sco_xy2theta(x, y);               /* SC.cpp:172 */
  if ((double)azim_range > p->max_radius) return 0;          /* SC.cpp:175 */
  *ring = imax(imin(p->R, d2i_x86(ceil(((double)azim_range / p->max_radius) * (double)p->R))), 1);  /* :178 */
  *sector = imax(imin(p->S, d2i_x86(ceil(((double)azim_angle / 360.0) * (double)p->S))), 1);        /* :179 */
  *height = h;
  return 1;
}

/* SC.cpp:151-195 */
void sco_make_sc(const sco_params* p, const void* pts, size_t n, size_t stride, double* out) {
  const int R = p->R, S = p->S;
  const double NO_POINT = -1000;
  for (int i = 0; i < R * S; ++i) out[i] = NO_POINT * 1.0; /* SC.cpp:159 */
  const unsigned char* b = (const unsigned char*)pts;
  for (size_t i = 0; i < n; ++i) {
    float xyz[3];
    memcpy(xyz, b + i * stride, sizeof xyz);
    int ring, sector;
    float h;
    if (!sco_bin_point(p, xyz[0], xyz[1], xyz[2], &ring, &sector, &h)) continue;
    double* cell = &out[(size_t)(sector - 1) * R + (ring - 1)];
    if (*cell < (double)h) *cell = (double)h; /* SC.cpp:182-183 */
  }
  for (int i = 0; i < R * S; ++i)
    if (out[i] == NO_POINT) out[i] = 0; /* SC.cpp:187-190 */
}

/* Eigen 3.3 Core/Redux.h, LinearVectorizedTraversal + NoUnrolling, Packet2d, alignedStart = 0: the order in
 * which sum()/mean()/norm()/dot() add their terms in the reference's default (SSE2) build. */
static double redux_sum(const double* v, int size) {
  const int packetSize = 2;
  const int alignedSize2 = (size / (2 * packetSize)) * (2 * packetSize);
  const int alignedSize = (size / packetSize) * packetSize;
  if (size == 0) return 0.0;
  double res;
  if (alignedSize) {
    double p00 = v[0], p01 = v[1];
    if (alignedSize > packetSize) {
      double p10 = v[2], p11 = v[3];
      for (int i = 2 * packetSize; i < alignedSize2; i += 2 * packetSize) {
        p00 = p00 + v[i];
        p01 = p01 + v[i + 1];
        p10 = p10 + v[i + 2];
        p11 = p11 + v[i + 3];
      }
      p00 = p00 + p10;
      p01 = p01 + p11;
      if (alignedSize > alignedSize2) {
        p00 = p00 + v[alignedSize2];
        p01 = p01 + v[alignedSize2 + 1];
      }
    }
    res = p00 + p01;
    for (int i = alignedSize; i < size; ++i) res = res + v[i];
  } else {
    res = v[0];
    for (int i = 1; i < size; ++i) res = res + v[i];
  }
  return res;
}

#define SCO_MAX_DIM 1024

/* SC.cpp:198-211: row(r).mean() */
void sco_ringkey(const sco_params* p, const double* sc, double* out_R) {
  double tmp[SCO_MAX_DIM];
  for (int r = 0; r < p->R; ++r) {
    for (int c = 0; c < p->S; ++c) tmp[c] = sc[(size_t)c * p->R + r];
    out_R[r] = redux_sum(tmp, p->S) / (double)p->S;
  }
}

/* SC.cpp:214-227: col(c).mean() */
void sco_sectorkey(const sco_params* p, const double* sc, double* out_S) {
  for (int c = 0; c < p->S; ++c) out_S[c] = redux_sum(sc + (size_t)c * p->R, p->R) / (double)p->R;
}

static double col_norm(const double* col, int R) {
  double tmp[SCO_MAX_DIM];
  for (int r = 0; r < R; ++r) tmp[r] = col[r] * col[r];
  return sqrt(redux_sum(tmp, R));
}

/* SC.cpp:93-113 (+ circshift SC.cpp:39-59 on the 1 x S key: shifted[j] = vkey2[(j - s) mod S]) */
int sco_fast_align(const sco_params* p, const double* vkey1, const double* vkey2) {
  const int S = p->S;
  int argmin = 0;
  double best = 10000000;
  double tmp[SCO_MAX_DIM];
  for (int s = 0; s < S; ++s) {
    for (int j = 0; j < S; ++j) {
      const double d = vkey1[j] - vkey2[(j - s + S) % S];
      tmp[j] = d * d;
    }
    const double nrm = sqrt(redux_sum(tmp, S));
    if (nrm < best) {
      argmin = s;
      best = nrm;
    }
  }
  return argmin;
}

/* SC.cpp:69-90 on (sc1, circshift(sc2, shift)) */
double sco_dist_direct_shifted(const sco_params* p, const double* sc1, const double* sc2, int shift) {
  const int R = p->R, S = p->S;
  int num_eff_cols = 0;
  double sum_sim = 0;
  double tmp[SCO_MAX_DIM];
  for (int j = 0; j < S; ++j) {
    const double* a = sc1 + (size_t)j * R;
    const double* b = sc2 + (size_t)((j - shift + S) % S) * R;
    const double na = col_norm(a, R), nb = col_norm(b, R);
    if ((na == 0) | (nb == 0)) continue;
    for (int r = 0; r < R; ++r) tmp[r] = a[r] * b[r];
    const double sim = redux_sum(tmp, R) / (na * nb);
    sum_sim = sum_sim + sim;
    num_eff_cols = num_eff_cols + 1;
  }
  const double sc_sim = sum_sim / num_eff_cols;
  return 1.0 - sc_sim;
}

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

/* SC.cpp:116-148 */
void sco_distance(const sco_params* p, const double* sc1, const double* sc2, double* dist, int* shift) {
  const int S = p->S;
  double v1[SCO_MAX_DIM], v2[SCO_MAX_DIM];
  sco_sectorkey(p, sc1, v1);
  sco_sectorkey(p, sc2, v2);
  const int a = sco_fast_align(p, v1, v2);
  const int radius = (int)round(0.5 * p->search_ratio * S); /* SC.cpp:123 */
  int space[2 * SCO_MAX_DIM + 1];
  int ns = 0;
  space[ns++] = a;
  for (int ii = 1; ii < radius + 1; ++ii) {
    space[ns++] = (a + ii + S) % S;
    space[ns++] = (a - ii + S) % S;
  }
  qsort(space, (size_t)ns, sizeof(int), cmp_int);
  int argmin_shift = 0;
  double min_d = 10000000;
  for (int i = 0; i < ns; ++i) {
    const double d = sco_dist_direct_shifted(p, sc1, sc2, space[i]);
    if (d < min_d) {
      argmin_shift = space[i];
      min_d = d;
    }
  }
  *dist = min_d;
  *shift = argmin_shift;
}

/* nf.hpp:383-408: groups of four squared differences added left to right, then added to the running sum. */
float sco_key_dist2(const float* a, const float* b, int dim) {
  float result = 0.0f;
  int d = 0;
  while (d + 3 < dim) {
    const float d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2], d3 = a[d + 3] - b[d + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    d += 4;
  }
  while (d < dim) {
    const float d0 = a[d] - b[d];
    result += d0 * d0;
    ++d;
  }
  return result;
}

/* Brute-force replacement of SC.cpp:283-289 (nanoflann kNN).  Scanning idx ascending and inserting AFTER
 * equal distances (nf.hpp:175-202 without NANOFLANN_FIRST_MATCH) yields the (dist, idx) total order; a point
 * is accepted only if strictly below the current K-th distance (nf.hpp:1360). */
int sco_knn(const sco_params* p, const float* keys, size_t n, const float* query, uint64_t* idx, float* d2) {
  const int K = p->num_candidates, R = p->R;
  for (int i = 0; i < K; ++i) {
    idx[i] = 0;  /* SC.cpp:283-284: value-initialised vectors */
    d2[i] = 0.0f;
  }
  if (K > 0) d2[K - 1] = FLT_MAX; /* nf.hpp:163-164 */
  int count = 0;
  for (size_t j = 0; j < n; ++j) {
    const float dist = sco_key_dist2(query, keys + j * (size_t)R, R);
    if (!(dist < d2[K - 1])) continue;
    int i;
    for (i = count; i > 0; --i) {
      if (d2[i - 1] > dist) {
        if (i < K) {
          d2[i] = d2[i - 1];
          idx[i] = idx[i - 1];
        }
      } else
        break;
    }
    if (i < K) {
      d2[i] = dist;
      idx[i] = j;
    }
    if (count < K) count++;
  }
  return count;
}

/* ------------------------------------------------------------------------------------------- */
struct sco_db {
  sco_params p;
  size_t n, cap;
  double* sc;        /* n * R*S */
  double* ringkey;   /* n * R  (polarcontext_invkeys_) */
  double* sectorkey; /* n * S  (polarcontext_vkeys_) */
  float* ringkey_f;  /* n * R  (polarcontext_invkeys_mat_) */
  int counter;       /* tree_making_period_conter */
  size_t n_tree;     /* polarcontext_invkeys_to_search_.size() */
};

sco_db* sco_db_create(const sco_params* p) {
  sco_db* db = (sco_db*)calloc(1, sizeof(sco_db));
  db->p = *p;
  return db;
}

void sco_db_destroy(sco_db* db) {
  if (!db) return;
  free(db->sc);
  free(db->ringkey);
  free(db->sectorkey);
  free(db->ringkey_f);
  free(db);
}

size_t sco_db_size(const sco_db* db) { return db->n; }

static void db_reserve(sco_db* db, size_t want) {
  if (want <= db->cap) return;
  size_t cap = db->cap ? db->cap * 2 : 256;
  while (cap < want) cap *= 2;
  const size_t RS = (size_t)db->p.R * db->p.S;
  db->sc = (double*)realloc(db->sc, cap * RS * sizeof(double));
  db->ringkey = (double*)realloc(db->ringkey, cap * db->p.R * sizeof(double));
  db->sectorkey = (double*)realloc(db->sectorkey, cap * db->p.S * sizeof(double));
  db->ringkey_f = (float*)realloc(db->ringkey_f, cap * db->p.R * sizeof(float));
  db->cap = cap;
}

/* SC.cpp:233-240 */
void sco_db_append_desc(sco_db* db, const double* sc) {
  db_reserve(db, db->n + 1);
  const int R = db->p.R, S = db->p.S;
  const size_t i = db->n;
  memcpy(db->sc + i * R * S, sc, sizeof(double) * R * S);
  sco_ringkey(&db->p, sc, db->ringkey + i * R);
  sco_sectorkey(&db->p, sc, db->sectorkey + i * S);
  for (int r = 0; r < R; ++r) db->ringkey_f[i * R + r] = (float)db->ringkey[i * R + r]; /* SC.cpp:62-66 */
  db->n = i + 1;
}

void sco_db_append_scan(sco_db* db, const void* pts, size_t n, size_t stride) {
  double sc[SCO_MAX_DIM * 8];
  double* buf = sc;
  const size_t RS = (size_t)db->p.R * db->p.S;
  if (RS > sizeof sc / sizeof sc[0]) buf = (double*)malloc(RS * sizeof(double));
  sco_make_sc(&db->p, pts, n, stride, buf);
  sco_db_append_desc(db, buf);
  if (buf != sc) free(buf);
}

void sco_db_get_entry(const sco_db* db, size_t i, double* sc, double* ringkey, double* sectorkey, float* ringkey_f) {
  const int R = db->p.R, S = db->p.S;
  if (sc) memcpy(sc, db->sc + i * R * S, sizeof(double) * R * S);
  if (ringkey) memcpy(ringkey, db->ringkey + i * R, sizeof(double) * R);
  if (sectorkey) memcpy(sectorkey, db->sectorkey + i * S, sizeof(double) * S);
  if (ringkey_f) memcpy(ringkey_f, db->ringkey_f + i * R, sizeof(float) * R);
}

/* SC.cpp:12-20 */
static float deg2rad_f(float degrees) { return (float)((double)degrees * M_PI / 180.0); }

/* SC.cpp:247-338 */
int sco_db_detect(sco_db* db, int* loop_id, float* yaw, double* min_dist_out, uint64_t* cand_idx, float* cand_d2,
                  double* cand_dist, int* cand_shift, uint64_t* n_tree) {
  const sco_params* p = &db->p;
  const int R = p->R, S = p->S, K = p->num_candidates;
  *loop_id = -1;
  *yaw = 0.0f;
  if (min_dist_out) *min_dist_out = 10000000;
  if (db->n < (size_t)p->exclude_recent + 1) return 0; /* SC.cpp:257-261 */
  if (db->counter % p->tree_period == 0) db->n_tree = db->n - (size_t)p->exclude_recent; /* SC.cpp:264-275 */
  db->counter = db->counter + 1;                                                          /* SC.cpp:276 */
  const float* curr_key = db->ringkey_f + (db->n - 1) * R;
  const double* curr_desc = db->sc + (db->n - 1) * (size_t)R * S;
  uint64_t* ci = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)K);
  float* cd = (float*)malloc(sizeof(float) * (size_t)K);
  sco_knn(p, db->ringkey_f, db->n_tree, curr_key, ci, cd);
  double min_dist = 10000000;
  int nn_align = 0, nn_idx = 0;
  for (int i = 0; i < K; ++i) { /* SC.cpp:296-311 */
    double d;
    int sh;
    sco_distance(p, curr_desc, db->sc + ci[i] * (size_t)R * S, &d, &sh);
    if (cand_idx) cand_idx[i] = ci[i];
    if (cand_d2) cand_d2[i] = cd[i];
    if (cand_dist) cand_dist[i] = d;
    if (cand_shift) cand_shift[i] = sh;
    if (d < min_dist) {
      min_dist = d;
      nn_align = sh;
      nn_idx = (int)ci[i];
    }
  }
  free(ci);
  free(cd);
  if (min_dist < p->dist_thres) *loop_id = nn_idx; /* SC.cpp:317-330 */
  const double unit_sector_angle = 360.0 / (double)S; /* SC.h:82 */
  *yaw = deg2rad_f((float)(nn_align * unit_sector_angle)); /* SC.cpp:333 */
  if (min_dist_out) *min_dist_out = min_dist;
  if (n_tree) *n_tree = db->n_tree;
  return K;
}

void sco_db_exhaustive(const sco_db* db, const double* query_sc, size_t n, int flipped, double* best_dist,
                       int* best_shift, int64_t* best_idx, int* best_flip) {
  const sco_params* p = &db->p;
  const int R = p->R, S = p->S;
  double bd = 10000000;
  int bs = 0, bf = 0;
  int64_t bi = 0;
  double* rev = flipped ? (double*)malloc(sizeof(double) * (size_t)R * S) : NULL;
  for (size_t i = 0; i < n && i < db->n; ++i) {
    const double* c = db->sc + i * (size_t)R * S;
    double d;
    int sh;
    sco_distance(p, query_sc, c, &d, &sh);
    if (d < bd) {
      bd = d;
      bs = sh;
      bi = (int64_t)i;
      bf = 0;
    }
    if (flipped) {
      for (int j = 0; j < S; ++j) memcpy(rev + (size_t)j * R, c + (size_t)(S - 1 - j) * R, sizeof(double) * R);
      sco_distance(p, query_sc, rev, &d, &sh);
      if (d < bd) {
        bd = d;
        bs = sh;
        bi = (int64_t)i;
        bf = 1;
      }
    }
  }
  free(rev);
  *best_dist = bd;
  *best_shift = bs;
  *best_idx = bi;
  *best_flip = bf;
}

/* ---- vectorised helpers for the parity tests ------------------------------------------------ */
void sco_atanf_many(const float* x, size_t n, float* out_port, float* out_libm) {
  for (size_t i = 0; i < n; ++i) {
    if (out_port) out_port[i] = sco_atanf(x[i]);
    if (out_libm) out_libm[i] = atanf(x[i]);
  }
}

/* bin = (sector-1)*R + (ring-1), or -1 when the point is not binned (outside the ROI / NaN coordinate or height,
 * which the product kernel also drops); height / theta as SC.cpp:168 / 172. */
void sco_bin_points(const sco_params* p, const float* xyz, size_t n, int32_t* bin, float* height, float* theta) {
  for (size_t i = 0; i < n; ++i) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    int ring, sector;
    float h = (float)((double)z + p->lidar_height);
    int ok = 0;
    if (x == x && y == y) ok = sco_bin_point(p, x, y, z, &ring, &sector, &h);
    if (ok && !(h == h)) ok = 0;
    bin[i] = ok ? (sector - 1) * p->R + (ring - 1) : -1;
    height[i] = h;
    theta[i] = sco_xy2theta(x, y);
  }
}
