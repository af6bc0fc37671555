/* scangen.c -- deterministic synthetic LiDAR scans / descriptors for parity tests and the benchmark.
 *
 * Host-only data generator (no Scan Context arithmetic lives here).  Shapes follow SURVEY.md 8(d):
 * HDL-64 = 64 beams x 1875 azimuth steps = 120,000 points; OS1-64 = 64 x 1024 = 65,536 points
 * (reference utility.h:101-102).  A "world" is a set of places, each a polar height field on a
 * 2 m x 3 deg grid with a place-specific per-ring base height (so ring keys discriminate places);
 * a trajectory visits every place once and then revisits random places with a random yaw and a
 * lateral offset, which is what makes loops exist.  Everything is a pure function of (seed, indices)
 * through a counter-based hash, so any scan can be generated independently and identically anywhere.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define SG_CELLS_R 60  /* 2 m radial cells out to 120 m */
#define SG_CELLS_A 120 /* 3 deg sectors */

typedef struct {
  uint64_t seed;
  int n_beams;        /* 64 */
  int n_azim;         /* 1875 (HDL-64) or 1024 (OS1-64) */
  int n_places;       /* scans [0, n_places) are first visits, later ones are revisits */
  float sensor_h;     /* ground is at z = -sensor_h in the sensor frame */
  float max_range;    /* beams sample ground ranges in [1.5, max_range]; > 80 exercises the ROI skip */
  float jitter;       /* lateral offset half-width of revisits [m] */
  float range_sigma;  /* range noise [m] */
} scangen_cfg;

static uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static uint64_t h4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  return mix64(mix64(mix64(mix64(seed) ^ a) ^ (b * 0x100000001b3ull)) ^ (c * 0xc2b2ae3d27d4eb4full));
}
static float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

/* height of the structure in world cell (k, s) of place p; 0 = bare ground */
static float place_height(uint64_t seed, int p, int k, int s) {
  const float base = 6.0f * u01(h4(seed, 1, (uint64_t)p, (uint64_t)k));
  const uint64_t h = h4(seed, 2, (uint64_t)p, (uint64_t)(k * SG_CELLS_A + s));
  if (u01(h) >= 0.12f) return 0.0f; /* 12 % of the 2 m x 3 deg cells carry a structure (~40 % of 20x60 bins) */
  return base + 3.0f * u01(mix64(h));
}

void scangen_default_cfg(scangen_cfg* c, int hdl64) {
  c->seed = 20181001ull;
  c->n_beams = 64;
  c->n_azim = hdl64 ? 1875 : 1024;
  c->n_places = 1000;
  c->sensor_h = hdl64 ? 1.73f : 2.0f;
  c->max_range = 110.0f;
  c->jitter = 0.5f;
  c->range_sigma = 0.02f;
}

/* pose of scan i: which place, yaw [rad], lateral offset */
void scangen_pose(const scangen_cfg* c, uint64_t i, int* place, float* yaw, float* dx, float* dy) {
  const uint64_t h = h4(c->seed, 3, i, 0);
  if (i < (uint64_t)c->n_places) {
    *place = (int)i;
    *dx = 0.f;
    *dy = 0.f;
  } else {
    *place = (int)(mix64(h) % (uint64_t)c->n_places);
    *dx = c->jitter * (2.f * u01(h4(c->seed, 3, i, 1)) - 1.f);
    *dy = c->jitter * (2.f * u01(h4(c->seed, 3, i, 2)) - 1.f);
  }
  *yaw = 6.2831853f * u01(h);
}

/* One scan: n_beams*n_azim points written stride bytes apart (float x,y,z first; the rest of each record,
 * if any, is zeroed).  Sensor frame. */
void scangen_scan(const scangen_cfg* c, uint64_t scan_index, void* out, size_t stride) {
  int place;
  float yaw, dx, dy;
  scangen_pose(c, scan_index, &place, &yaw, &dx, &dy);
  /* cache the place's height field */
  static _Thread_local float H[SG_CELLS_R][SG_CELLS_A];
  for (int k = 0; k < SG_CELLS_R; ++k)
    for (int s = 0; s < SG_CELLS_A; ++s) H[k][s] = place_height(c->seed, place, k, s);
  unsigned char* o = (unsigned char*)out;
  const float two_pi = 6.2831853f;
  for (int e = 0; e < c->n_beams; ++e) {
    for (int a = 0; a < c->n_azim; ++a) {
      const uint64_t h = h4(c->seed ^ 0x5ca9ull, scan_index, (uint64_t)e, (uint64_t)a);
      const uint64_t h2 = mix64(h), h3 = mix64(h2);
      const float ub = ((float)e + u01(h)) / (float)c->n_beams;
      float rho = 1.5f + (c->max_range - 1.5f) * ub * ub;
      /* sum of 4 uniforms ~ normal enough for 2 cm of range noise */
      const float g = (u01(h2) + u01(h2 << 24 | h2 >> 40) + u01(h3) + u01(h3 << 24 | h3 >> 40) - 2.0f) * 1.7320508f;
      rho += c->range_sigma * g;
      const float phi = two_pi * ((float)a + u01(mix64(h3))) / (float)c->n_azim;
      const float xs = rho * cosf(phi), ys = rho * sinf(phi);
      /* world coordinates relative to the place centre */
      const float xw = dx + rho * cosf(phi + yaw), yw = dy + rho * sinf(phi + yaw);
      const float rw = sqrtf(xw * xw + yw * yw);
      float aw = atan2f(yw, xw);
      if (aw < 0.f) aw += two_pi;
      int k = (int)(rw * 0.5f), s = (int)(aw * ((float)SG_CELLS_A / two_pi));
      if (k >= SG_CELLS_R) k = SG_CELLS_R - 1;
      if (s >= SG_CELLS_A) s = SG_CELLS_A - 1;
      const float uz = sqrtf(u01(mix64(h3 ^ 0x77ull)));
      const float z = H[k][s] * uz - c->sensor_h;
      float rec[3] = {xs, ys, z};
      memset(o, 0, stride);
      memcpy(o, rec, sizeof rec);
      o += stride;
    }
  }
}

/* A descriptor generated directly in descriptor space (R x S, column-major floats): the place's height
 * field resampled on the R x S grid, rotated by the visit's yaw, with per-visit perturbations.  Used to
 * fill large databases without generating 100k scans. */
void scangen_desc(const scangen_cfg* c, uint64_t scan_index, int R, int S, float* out) {
  int place;
  float yaw, dx, dy;
  scangen_pose(c, scan_index, &place, &yaw, &dx, &dy);
  const int shift = (int)(yaw / 6.2831853f * (float)S) % S;
  const int kr = 40 / R > 0 ? 40 / R : 1, ks = SG_CELLS_A / S > 0 ? SG_CELLS_A / S : 1; /* world cells per bin */
  for (int col = 0; col < S; ++col) {
    const int ws0 = ((col + shift) % S) * SG_CELLS_A / S;
    for (int r = 0; r < R; ++r) {
      const int wk0 = r * 40 / R; /* 80 m = 40 world cells */
      float v = 0.f;
      for (int a = 0; a < kr; ++a)
        for (int b = 0; b < ks; ++b) {
          const float hcell = place_height(c->seed, place, wk0 + a, (ws0 + b) % SG_CELLS_A);
          if (hcell > v) v = hcell;
        }
      const uint64_t h = h4(c->seed ^ 0xde5cull, scan_index, (uint64_t)col, (uint64_t)r);
      if (v > 0.f) {
        v += 2.0f - 0.3f * u01(mix64(h));          /* max over the bin's returns + LIDAR_HEIGHT-like offset */
        if (u01(mix64(h ^ 1)) < 0.03f) v = 0.27f;  /* structure occasionally missed */
      } else {
        v = (u01(h) < 0.75f) ? 0.27f : 0.f;        /* ground return, or an empty bin (zeros ~ 15 % overall) */
      }
      out[(size_t)col * R + r] = v;
    }
  }
}
