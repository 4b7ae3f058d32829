/* pair_epilogue.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (CPU arm of bench.py; never linked into the product).
 *
 * Fused, OpenMP-parallel element-wise tails of the CPU pair tests in oracle/cpu_arm.py, so that the CPU baseline is
 * bound by its FP64 DGEMM (OpenBLAS) and not by NumPy's single-threaded element-wise passes.
 *   fantasy_epilogue:   c = sf2*exp(-dist/2) - acc ; mu' = m_z + c*a_x ; s2' = max(s_z - c^2 b_x, 0) ;
 *                       ok &= mu' - beta*sqrt(s2') >= 0        (oracle/gp_oracle.py fantasy_counts, same operations)
 *   lipschitz_epilogue: hit_x |= any_z ( ucb_x - L*sqrt(max(d2,0)) >= 0 )   (models/SafeOpt.py:85-88 of the reference)
 * Build: gcc -O2 -fopenmp -shared -fPIC (oracle/Makefile). */
#include <math.h>
#include <stdint.h>
#include <omp.h>

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets the thread count explicitly */
void epilogue_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int epilogue_max_threads(void) { return omp_get_max_threads(); }

void fantasy_epilogue(int64_t nz, int64_t nx, const double* dist, const double* acc, const double* mz, const double* sz,
                      const double* ax, const double* bx, double sf2, double beta, uint8_t* ok) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nz; ++i) {
    const double m = mz[i], s = sz[i];
    for (int64_t j = 0; j < nx; ++j) {
      const double c = sf2 * exp(-0.5 * dist[i * nx + j]) - acc[i * nx + j];
      const double mu = m + c * ax[j];
      double s2 = s - c * c * bx[j];
      if (s2 < 0.0) s2 = 0.0;
      if (!(mu - beta * sqrt(s2) >= 0.0)) ok[i * nx + j] = 0;
    }
  }
}

/* d2[nx][nz] squared raw distances (dot form); r[nx] = ucb_x / L (negative: never reaches); hit[nx] in/out */
void lipschitz_epilogue(int64_t nx, int64_t nz, const double* d2, const double* ucb, double L, uint8_t* hit) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nx; ++i) {
    if (hit[i]) continue;
    const double u = ucb[i];
    uint8_t h = 0;
    for (int64_t j = 0; j < nz && !h; ++j) {
      const double q = d2[i * nz + j];
      if (u - L * sqrt(q > 0.0 ? q : 0.0) >= 0.0) h = 1;
    }
    hit[i] = h;
  }
}
