/* synth_raycast.c — C restatement of vloam_b200/synth.py:Scene.raycast (test / bench infrastructure only).
 *
 * The numpy ray caster costs ~1.2 s per 64 x 2048 scan; the benchmark's forward trajectories and the parity sweeps need
 * hundreds of scans.  Every operation below is the same IEEE double operation, in the same order, as the numpy code
 * (compile with -ffp-contract=off: no FMA), so both produce identical bits; tests/test_synth_and_bench_helpers.py
 * checks that, and synth.py falls back to numpy when this library is missing or disagrees.
 */
#include <math.h>
#include <stddef.h>

static inline double nmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }   /* np.minimum */
static inline double nmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }   /* np.maximum */

/* o[3]; d[M][3]; box_min/box_max [nb][3]; pole_xy [np][2]; pole_r, pole_top [np]; best[M] out */
void synth_raycast(const double* o, const double* d, long M, const double* box_min, const double* box_max, int nb,
                   const double* pole_xy, const double* pole_r, const double* pole_top, int np_, double ground_z,
                   double* best) {
  for (long i = 0; i < M; ++i) {
    const double dx = d[3 * i], dy = d[3 * i + 1], dz = d[3 * i + 2];
    double b = INFINITY;
    /* ground */
    {
      double tg = (ground_z - o[2]) / dz;
      tg = ((dz < 0) && (tg > 0)) ? tg : INFINITY;
      b = nmin(b, tg);
    }
    /* boxes (slab test) */
    const double ix = 1.0 / dx, iy = 1.0 / dy, iz = 1.0 / dz;
    for (int k = 0; k < nb; ++k) {
      const double* mn = box_min + 3 * k;
      const double* mx = box_max + 3 * k;
      const double t0x = (mn[0] - o[0]) * ix, t0y = (mn[1] - o[1]) * iy, t0z = (mn[2] - o[2]) * iz;
      const double t1x = (mx[0] - o[0]) * ix, t1y = (mx[1] - o[1]) * iy, t1z = (mx[2] - o[2]) * iz;
      const double tmin = nmax(nmax(nmin(t0x, t1x), nmin(t0y, t1y)), nmin(t0z, t1z));
      const double tmax = nmin(nmin(nmax(t0x, t1x), nmax(t0y, t1y)), nmax(t0z, t1z));
      const int hit = (tmax >= tmin) && (tmax > 0);
      const double t = (tmin > 0) ? tmin : tmax;
      if (hit && (t < b)) b = t;
    }
    /* poles (vertical cylinders) */
    const double a = dx * dx + dy * dy;
    for (int p = 0; p < np_; ++p) {
      const double ocx = o[0] - pole_xy[2 * p], ocy = o[1] - pole_xy[2 * p + 1];
      const double bq = dx * ocx + dy * ocy;
      const double cq = (ocx * ocx + ocy * ocy) - pole_r[p] * pole_r[p];
      const double disc = bq * bq - a * cq;
      const int ok = disc > 0;
      const double sq = sqrt(ok ? disc : 0.0);
      const double t = (-bq - sq) / a;
      const double z = o[2] + t * dz;
      const int hit = ok && (t > 0) && (z <= pole_top[p]) && (z >= ground_z);
      if (hit && (t < b)) b = t;
    }
    best[i] = b;
  }
}
