/* CPU ORACLE (test infrastructure / CPU baseline only — never linked into the product).
 *
 * Plain C + OpenMP restatement of the three memory-bound kernels of the hot path, used to time
 * the CPU baseline with all host cores (bench.py cpu_baseline / --impl reference) and
 * cross-checked against the numpy oracle in tests/test_oracle_cport.py.
 *
 *   oracle_patch_apply  — PCApply_PATCH additive (alfi/solver.py:318-324):
 *                         y += sum_i R_i^T Ainv_i R_i x, patches of one colour in parallel
 *   oracle_bsr_spmv     — MatMult_SeqBAIJ on the baij velocity block (alfi/solver.py:512)
 *   oracle_csr_apply    — scalar CSR (x) I_bs, the standard prolongation/restriction
 *                         (alfi/transfer.py:284-290)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* inv: row-major n_i x n_i blocks, patch p at inv + inv_off[p]; order/colours as in the library */
void oracle_patch_apply(int32_t npatch, const int64_t* off, const int32_t* dofs, int32_t norder,
                        const int32_t* order, const int32_t* colours, int32_t ncolour,
                        const int64_t* inv_off, const double* inv, const double* x, double* y) {
  (void)npatch;
  for (int32_t col = 0; col < ncolour; ++col) {
#pragma omp parallel
    {
      double* r = NULL;
      int64_t cap = 0;
#pragma omp for schedule(dynamic, 1)
      for (int32_t q = 0; q < norder; ++q) {
        const int32_t p = order[q];
        if (colours[p] != col) continue;
        const int64_t o = off[p];
        const int64_t n = off[p + 1] - o;
        if (n == 0) continue;
        if (n > cap) {
          free(r);
          r = (double*)malloc(sizeof(double) * (size_t)n);
          cap = n;
        }
        const int32_t* I = dofs + o;
        for (int64_t j = 0; j < n; ++j) r[j] = x[I[j]];
        const double* A = inv + inv_off[p];
        for (int64_t i = 0; i < n; ++i) {
          const double* row = A + i * n;
          double s = 0.0;
          for (int64_t j = 0; j < n; ++j) s += row[j] * r[j];
          y[I[i]] += s;            /* patches of one colour share no dof: race free */
        }
      }
      free(r);
    }
  }
}

void oracle_bsr_spmv(int32_t nbrows, int32_t bs, const int32_t* rowptr, const int32_t* colidx,
                     const double* vals, const double* x, double* y) {
  const int b2 = bs * bs;
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < nbrows; ++i) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const double* v = vals + (int64_t)k * b2;
      const double* xc = x + (int64_t)colidx[k] * bs;
      for (int r = 0; r < bs; ++r)
        for (int c = 0; c < bs; ++c) acc[r] += v[r * bs + c] * xc[c];
    }
    for (int r = 0; r < bs; ++r) y[(int64_t)i * bs + r] = acc[r];
  }
}

void oracle_csr_apply(int32_t nrows, int32_t bs, const int32_t* rowptr, const int32_t* colidx,
                      const double* vals, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < nrows; ++i) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      for (int r = 0; r < bs; ++r) acc[r] += vals[k] * x[(int64_t)colidx[k] * bs + r];
    for (int r = 0; r < bs; ++r) y[(int64_t)i * bs + r] = acc[r];
  }
}
