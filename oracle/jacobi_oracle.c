/*
 * CPU oracle (plain C + OpenMP) for fpie's Jacobi Poisson hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the
 * product package.  Users: tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs (as the checker / the CPU baseline).
 *
 * Restates the reference's TRUE-Jacobi iteration for sizes where the numpy
 * restatement (oracle/np_oracle.py) is too slow.  Parity is PINNED: the test
 * suite checks every routine here bit-for-bit against oracle/np_oracle.py,
 * which in turn is pinned to golden vectors produced by the reference's own
 * numpy backend (tests/golden/fpie_numpy_golden.npz).
 *
 * Reference lines followed (paths relative to the reference checkout):
 *   Equ sweep      fpie/np_solver.py:33-41   ((((B+X[a0])+X[a1])+X[a2])+X[a3])/4
 *   Equ residual   fpie/np_solver.py:42-50
 *   Grid sweep     fpie/np_solver.py:81-88   ((((g+up)+down)+left)+right)/4, masked px only
 *   Grid residual  fpie/np_solver.py:90-96   ((((4t-g)-up)-down)-left)-right
 *   clip + u8      fpie/core/openmp/equ.cc:120-123, grid.cc:105-108 (truncation)
 *   row bands      fpie/core/mpi/grid.cc:27-31 (offset rule; see np_oracle.band_offsets)
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off and the absence of -ffast-math are REQUIRED: results must
 * be the fp32 round-to-nearest value of each individual add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- GridSolver ------------------------------------------------------ */

/* One sweep src -> dst over rows [r0, r1).  Frame pixels are never masked
 * (caller guarantee, fpie/process.py:342-351); to stay memory-safe for hostile
 * input we skip them explicitly. */
static void grid_sweep_rows(int64_t n, int64_t m, const int32_t *mask, const float *g,
                            const float *src, float *dst, int64_t r0, int64_t r1) {
  const int64_t m3 = m * 3;
  for (int64_t r = r0; r < r1; ++r) {
    const float *row = src + r * m3;
    float *out = dst + r * m3;
    if (r == 0 || r == n - 1) {
      memcpy(out, row, sizeof(float) * (size_t)m3);
      continue;
    }
    const int32_t *mk = mask + r * m;
    const float *gr = g + r * m3;
    for (int64_t c = 0; c < m; ++c) {
      const int64_t e = c * 3;
      if (mk[c] && c > 0 && c < m - 1) {
        for (int ch = 0; ch < 3; ++ch) {
          float s = gr[e + ch] + row[e + ch - m3]; /* + up    */
          s = s + row[e + ch + m3];                /* + down  */
          s = s + row[e + ch - 3];                 /* + left  */
          s = s + row[e + ch + 3];                 /* + right */
          out[e + ch] = s / 4.0f;
        }
      } else {
        out[e] = row[e];
        out[e + 1] = row[e + 1];
        out[e + 2] = row[e + 2];
      }
    }
  }
}

/* iters true-Jacobi sweeps in place on tgt[n][m][3]; scratch has the same size. */
API void oracle_grid_sweeps(int64_t n, int64_t m, const int32_t *mask, float *tgt,
                            const float *grad, float *scratch, int iters, int threads) {
  float *a = tgt, *b = scratch;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  for (int it = 0; it < iters; ++it) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) grid_sweep_rows(n, m, mask, grad, a, b, r, r + 1);
    float *t = a;
    a = b;
    b = t;
  }
  if (a != tgt) memcpy(tgt, a, sizeof(float) * (size_t)(n * m * 3));
}

/* err32: fp32 accumulation in row-major order of masked pixels (what numpy's
 * sum(axis=0) over the [K,3] term array does); err64: same fp32 terms summed
 * in double. */
API void oracle_grid_residual(int64_t n, int64_t m, const int32_t *mask, const float *tgt,
                              const float *grad, float *err32, double *err64) {
  const int64_t m3 = m * 3;
  float s32[3] = {0.f, 0.f, 0.f};
  double s64[3] = {0., 0., 0.};
  for (int64_t r = 1; r + 1 < n; ++r) {
    for (int64_t c = 1; c + 1 < m; ++c) {
      if (!mask[r * m + c]) continue;
      const float *p = tgt + r * m3 + c * 3;
      const float *gp = grad + r * m3 + c * 3;
      for (int ch = 0; ch < 3; ++ch) {
        float t = 4.0f * p[ch] - gp[ch];
        t = t - p[ch - m3];
        t = t - p[ch + m3];
        t = t - p[ch - 3];
        t = t - p[ch + 3];
        t = fabsf(t);
        s32[ch] += t;
        s64[ch] += (double)t;
      }
    }
  }
  for (int ch = 0; ch < 3; ++ch) {
    if (err32) err32[ch] = s32[ch];
    if (err64) err64[ch] = s64[ch];
  }
}

/* ---- EquSolver ------------------------------------------------------- */

/* iters true-Jacobi sweeps in place on X[n][3]; A[n][4] (up,down,left,right;
 * 0 = the constant-zero row), B[n][3]; scratch[n][3]. Row 0 maps to itself. */
API void oracle_equ_sweeps(int64_t n, const int32_t *A, float *X, const float *B,
                           float *scratch, int iters, int threads) {
  float *a = X, *b = scratch;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  for (int it = 0; it < iters; ++it) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const int32_t *nb = A + i * 4;
      for (int ch = 0; ch < 3; ++ch) {
        float s = B[i * 3 + ch] + a[(int64_t)nb[0] * 3 + ch];
        s = s + a[(int64_t)nb[1] * 3 + ch];
        s = s + a[(int64_t)nb[2] * 3 + ch];
        s = s + a[(int64_t)nb[3] * 3 + ch];
        b[i * 3 + ch] = s / 4.0f;
      }
    }
    float *t = a;
    a = b;
    b = t;
  }
  if (a != X) memcpy(X, a, sizeof(float) * (size_t)(n * 3));
}

API void oracle_equ_residual(int64_t n, const int32_t *A, const float *X, const float *B,
                             float *err32, double *err64) {
  float s32[3] = {0.f, 0.f, 0.f};
  double s64[3] = {0., 0., 0.};
  for (int64_t i = 0; i < n; ++i) {
    const int32_t *nb = A + i * 4;
    for (int ch = 0; ch < 3; ++ch) {
      float s = B[i * 3 + ch] + X[(int64_t)nb[0] * 3 + ch];
      s = s + X[(int64_t)nb[1] * 3 + ch];
      s = s + X[(int64_t)nb[2] * 3 + ch];
      s = s + X[(int64_t)nb[3] * 3 + ch];
      s = s - 4.0f * X[i * 3 + ch];
      s = fabsf(s);
      s32[ch] += s;
      s64[ch] += (double)s;
    }
  }
  for (int ch = 0; ch < 3; ++ch) {
    if (err32) err32[ch] = s32[ch];
    if (err64) err64[ch] = s64[ch];
  }
}

/* ---- output ---------------------------------------------------------- */

API void oracle_clip_u8(int64_t count, const float *v, uint8_t *out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < count; ++i) {
    float x = v[i];
    out[i] = x < 0.f ? 0 : (x > 255.f ? 255 : (uint8_t)x);
  }
}
