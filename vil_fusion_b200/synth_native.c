/* Synthetic spinning-lidar scan generator (test / bench INPUT only, not part of the product path).
 * Analytic ray cast of a ground plane + axis-aligned boxes, Gaussian range noise from a counter-based
 * RNG (so a scan depends only on (seed, frame), not on thread count).  Built by __graft_entry__.build()
 * into vil_fusion_b200/libvilf_synth.so; vil_fusion_b200/synth.py is the Python front end.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline double u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* boxes: [nb][6] xmin ymin zmin xmax ymax zmax (world frame); R row-major 3x3, t: sensor pose in the world.
 * elev_rad: [n_rings]; azimuth of column a = -pi + (a + 0.5) * 2 pi / n_az.
 * order 0: ring-major, azimuth ascending inside a ring; order 1: azimuth-major (firing order).
 * Outputs hold up to n_rings * n_az points; returns the number of returns kept. */
int synth_scan(const double* boxes, int nb, double ground_z, const double* R, const double* t, const double* elev_rad, int n_rings, int n_az,
               double noise, double max_range, uint64_t seed, uint64_t frame, int order, float* xyzi_out, uint16_t* ring_out) {
  const int n = n_rings * n_az;
  float* tmp = (float*)malloc((size_t)n * 4 * sizeof(float));
  unsigned char* ok = (unsigned char*)malloc((size_t)n);
  if (!tmp || !ok) { free(tmp); free(ok); return -1; }
  /* boxes that can be hit inside max_range */
  int* near = (int*)malloc((size_t)(nb > 0 ? nb : 1) * sizeof(int));
  int nn = 0;
  for (int j = 0; j < nb; ++j) {
    const double* b = boxes + 6 * (size_t)j;
    double cx = t[0] < b[0] ? b[0] : (t[0] > b[3] ? b[3] : t[0]);
    double cy = t[1] < b[1] ? b[1] : (t[1] > b[4] ? b[4] : t[1]);
    cx -= t[0]; cy -= t[1];
    if (cx * cx + cy * cy < max_range * max_range) near[nn++] = j;
  }
  const double two_pi = 6.283185307179586476925286766559;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    const int r = i / n_az, a = i % n_az;
    const double e = elev_rad[r];
    const double az = -3.14159265358979323846 + (a + 0.5) * (two_pi / n_az);
    const double ce = cos(e), se = sin(e), ca = cos(az), sa = sin(az);
    const double ds[3] = {ce * ca, ce * sa, se};            /* sensor frame */
    double d[3];
    for (int k = 0; k < 3; ++k) d[k] = R[k * 3 + 0] * ds[0] + R[k * 3 + 1] * ds[1] + R[k * 3 + 2] * ds[2];
    double best = INFINITY;
    if (d[2] < -1e-9) {
      const double sg = (ground_z - t[2]) / d[2];
      if (sg > 0) best = sg;
    }
    const double i0 = 1.0 / d[0], i1 = 1.0 / d[1], i2 = 1.0 / d[2];
    for (int q = 0; q < nn; ++q) {
      const double* b = boxes + 6 * (size_t)near[q];
      double lo = (b[0] - t[0]) * i0, hi = (b[3] - t[0]) * i0;
      double tmin = lo < hi ? lo : hi, tmax = lo < hi ? hi : lo;
      lo = (b[1] - t[1]) * i1; hi = (b[4] - t[1]) * i1;
      double a0 = lo < hi ? lo : hi, a1 = lo < hi ? hi : lo;
      if (a0 > tmin) tmin = a0;
      if (a1 < tmax) tmax = a1;
      lo = (b[2] - t[2]) * i2; hi = (b[5] - t[2]) * i2;
      a0 = lo < hi ? lo : hi; a1 = lo < hi ? hi : lo;
      if (a0 > tmin) tmin = a0;
      if (a1 < tmax) tmax = a1;
      if (tmax >= tmin && tmin > 0.0 && tmin < best) best = tmin;
    }
    int keep = isfinite(best) && best < max_range;
    const uint64_t h0 = mix64(mix64(seed * 0x100000001B3ull + frame) + (uint64_t)i * 3u);
    const uint64_t h1 = mix64(h0 + 1u), h2 = mix64(h0 + 2u);
    const double g = sqrt(-2.0 * log(u01(h0))) * cos(two_pi * u01(h1));  /* Box-Muller */
    const double s = best + noise * g;
    keep = keep && s > 0.5;
    ok[i] = (unsigned char)keep;
    if (keep) {
      tmp[4 * (size_t)i + 0] = (float)(ds[0] * s);
      tmp[4 * (size_t)i + 1] = (float)(ds[1] * s);
      tmp[4 * (size_t)i + 2] = (float)(ds[2] * s);
      tmp[4 * (size_t)i + 3] = (float)u01(h2);
    }
  }
  int m = 0;
  if (order == 0) {
    for (int i = 0; i < n; ++i)
      if (ok[i]) { for (int k = 0; k < 4; ++k) xyzi_out[4 * (size_t)m + k] = tmp[4 * (size_t)i + k]; ring_out[m++] = (uint16_t)(i / n_az); }
  } else {
    for (int a = 0; a < n_az; ++a)
      for (int r = 0; r < n_rings; ++r) {
        const int i = r * n_az + a;
        if (ok[i]) { for (int k = 0; k < 4; ++k) xyzi_out[4 * (size_t)m + k] = tmp[4 * (size_t)i + k]; ring_out[m++] = (uint16_t)r; }
      }
  }
  free(tmp); free(ok); free(near);
  return m;
}
