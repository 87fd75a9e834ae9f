// The pieces that bracket fieldsplit_0 in alfi's outer solver, on the device so that one outer
// Krylov iteration — or the whole linear solve of a Newton step — costs one PCIe round trip instead of one per
// velocity-block application:
//
//   * PCFIELDSPLIT schur, full factorisation (alfi/solver.py:405-421): y1 = A^-1 r_u ; y_p = S^-1 (r_p - B y1) ;
//     y_u = A^-1 (r_u - B^T y_p), with A^-1 = the multigrid cycle (cycle.cu) applied twice;
//   * fieldsplit_1 = alfi.solver.DGMassInv (solver.py:15-38): S^-1 = -(nu + gamma) M_p^-1, a block-diagonal matvec
//     for the discontinuous pressure spaces of both element pairs, followed by the removal of the constant-pressure
//     nullspace (alfi/problem.py:33-38);
//   * the B / B^T products of the saddle-point Jacobian [A B^T; B 0] (MatMult of the nest blocks);
//   * the outer KSP: right-preconditioned FGMRES(restart) with classical Gram-Schmidt (solver.py:463-474), the
//     Hessenberg least-squares update on the host (k + 2 doubles cross PCIe per iteration).
//
// Everything here is HBM-bound BLAS-1 / SpMV work around the two cycle applications, which are >= 95 % of the time.
#include <algorithm>
#include <cmath>

#include "alfib_internal.h"

namespace {

constexpr int ST = 256, SGRID = 296;      // fixed reduction grid => fixed summation order

__global__ void __launch_bounds__(ST) sum_partial_kernel(int n, const double* __restrict__ v, double* __restrict__ partial) {
  __shared__ double red[ST / 32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x; i < n; i += (int64_t)SGRID * ST) s += v[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < ST / 32 ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}

// v -= mean(v): every block forms the mean from the SGRID partial sums in the same fixed order
__global__ void __launch_bounds__(ST) subtract_mean_kernel(int n, double* __restrict__ v, const double* __restrict__ partial) {
  __shared__ double mean;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int b = threadIdx.x; b < SGRID; b += 32) s += partial[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) mean = s / (double)n;
  }
  __syncthreads();
  const double m = mean;
  for (int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x; i < n; i += (int64_t)gridDim.x * ST) v[i] -= m;
}

void csr(alfib_ctx* c, int nrows, const DBuf<int32_t>& rp, const DBuf<int32_t>& ci, const DBuf<double>& v, int64_t nnz,
         const double* x, double* y) {
  launch_csr_apply(c, nrows, 1, rp.p, ci.p, v.p, x, y, nnz);
}

Level& finest(alfib_ctx* c) {
  ALFIB_REQUIRE(c->nlevels >= 1, "alfib_cycle_setup first");
  return *c->levels[c->nlevels - 1];
}

}  // namespace

// y = P^-1 r for r = [r_u; r_p] (device pointers, no aliasing)
void schur_apply_device(alfib_ctx* c, double nu, double gamma, const double* r, double* y) {
  Schur& S = c->schur;
  ALFIB_REQUIRE(S.on, "alfib_schur_set has not been called");
  ALFIB_REQUIRE(r != y, "schur apply: input and output must not alias");
  const int nu_ = S.nu, np_ = S.np;
  const double* ru = r;
  const double* rp = r + nu_;
  double* yu = y;
  double* yp = y + nu_;
  krylov_reserve(c);
  cycle_apply_device(c, ru, S.y1.p);                                   // y1 = A^-1 r_u
  csr(c, np_, S.b_rowptr, S.b_colidx, S.b_vals, S.b_nnz, S.y1.p, S.tp.p);       // B y1
  launch_sub(c, np_, rp, S.tp.p, S.tp2.p);                             // r_p - B y1
  csr(c, np_, S.mi_rowptr, S.mi_colidx, S.mi_vals, S.mi_nnz, S.tp2.p, S.tp.p);  // M_p^-1 (.)
  launch_axpby(c, np_, -(nu + gamma), S.tp.p, 0.0, yp);                // DGMassInv.apply (solver.py:32-35)
  if (S.remove_mean) {
    sum_partial_kernel<<<SGRID, ST, 0, c->stream>>>(np_, yp, c->partial.p);
    subtract_mean_kernel<<<SGRID, ST, 0, c->stream>>>(np_, yp, c->partial.p);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  csr(c, nu_, S.bt_rowptr, S.bt_colidx, S.bt_vals, S.b_nnz, yp, S.tu.p);        // B^T y_p
  launch_sub(c, nu_, ru, S.tu.p, S.y1.p);                              // r_u - B^T y_p
  cycle_apply_device(c, S.y1.p, yu);                                   // y_u = A^-1 (.)
}

// out = [A z_u + B^T z_p ; B z_u]
void jacobian_apply_device(alfib_ctx* c, const double* z, double* out) {
  Schur& S = c->schur;
  ALFIB_REQUIRE(S.on, "alfib_schur_set has not been called");
  ALFIB_REQUIRE(z != out, "jacobian apply: input and output must not alias");
  Level& L = finest(c);
  ALFIB_REQUIRE(L.has_values, "finest level has no values");
  launch_bsr_spmv(c, L, L.vals.p, z, out, nullptr);
  csr(c, S.nu, S.bt_rowptr, S.bt_colidx, S.bt_vals, S.b_nnz, z + S.nu, S.tu.p);
  launch_axpby(c, S.nu, 1.0, S.tu.p, 1.0, out);
  csr(c, S.np, S.b_rowptr, S.b_colidx, S.b_vals, S.b_nnz, z, out + S.nu);
}

// KSPFGMRES as alfi configures the outer solver (solver.py:463-474; PETSc defaults: right preconditioning, classical
// Gram-Schmidt without refinement, restart 30, zero initial guess, convergence on the recurrence residual against
// max(rtol |b|, atol)).  Vectors and BLAS-1 on the device; the (restart + 1) x restart Hessenberg problem on the host.
void outer_solve_device(alfib_ctx* c, double nu, double gamma, const double* b, double* x, double rtol, double atol,
                        int maxit, int restart, int* iterations, double* history, int nhistory) {
  Schur& S = c->schur;
  ALFIB_REQUIRE(S.on, "alfib_schur_set has not been called");
  ALFIB_REQUIRE(restart >= 1 && restart <= ALFIB_MAX_KRYLOV, "restart out of range");
  ALFIB_REQUIRE(maxit >= 0, "negative iteration limit");
  const int n = S.nu + S.np;
  cudaStream_t s = c->stream;
  if (S.restart < restart) {
    S.V.alloc((size_t)(restart + 1) * n);
    S.Z.alloc((size_t)restart * n);
    S.restart = restart;
  }
  S.w.alloc(n);
  S.r.alloc(n);
  S.hd.alloc(ALFIB_MAX_KRYLOV + 4);
  krylov_reserve(c);
  double* V = S.V.p;
  double* Z = S.Z.p;
  double* w = S.w.p;
  double* hd = S.hd.p;                       // [0, restart]: h ; [restart + 1]: norm ; [restart + 2]: 1 / norm
  double* d_nrm = hd + ALFIB_MAX_KRYLOV + 1;
  double* d_inv = hd + ALFIB_MAX_KRYLOV + 2;
  int nh = 0;
  auto record = [&](double v) {
    if (history && nh < nhistory) history[nh] = v;
    ++nh;
  };
  auto norm_to_host = [&](double* vec) {     // |vec| (device inverse left in d_inv)
    launch_maxpy_norm(c, n, 0, hd, 1.0, V, n, vec, d_nrm, d_inv);
    double v = 0.0;
    CUDA_TRY(cudaMemcpyAsync(&v, d_nrm, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return v;
  };
  CUDA_TRY(cudaMemsetAsync(x, 0, sizeof(double) * n, s));
  CUDA_TRY(cudaMemcpyAsync(S.r.p, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
  double beta = norm_to_host(S.r.p);
  const double r0 = beta;
  record(beta);
  int its = 0;
  if (iterations) *iterations = 0;
  if (!(beta > std::max(atol, 0.0))) return;
  const double target = std::max(rtol * r0, atol);
  std::vector<double> H, cs, sn, g, y, hcol;
  while (its < maxit) {
    const int m = std::min(restart, maxit - its);
    H.assign((size_t)(m + 1) * m, 0.0);      // column-major, ld = m + 1
    cs.assign(m, 0.0);
    sn.assign(m, 0.0);
    g.assign(m + 1, 0.0);
    g[0] = beta;
    launch_scale_by(c, n, d_inv, S.r.p, V);                                   // v_0 = r / beta
    int kdone = 0;
    bool converged = false;
    for (int k = 0; k < m; ++k) {
      double* zk = Z + (size_t)k * n;
      schur_apply_device(c, nu, gamma, V + (size_t)k * n, zk);                // z_k = P^-1 v_k
      jacobian_apply_device(c, zk, w);                                        // w = J z_k
      launch_multi_dot(c, n, k + 1, V, n, w, hd);                             // h = V^T w
      launch_maxpy_norm(c, n, k + 1, hd, -1.0, V, n, w, hd + k + 1, d_inv);   // w -= V h ; |w|
      hcol.assign(k + 2, 0.0);
      CUDA_TRY(cudaMemcpyAsync(hcol.data(), hd, sizeof(double) * (k + 2), cudaMemcpyDeviceToHost, s));
      launch_scale_by(c, n, d_inv, w, V + (size_t)(k + 1) * n);               // v_{k+1} = w / |w|  (0 on breakdown)
      CUDA_TRY(cudaStreamSynchronize(s));
      double* Hk = H.data() + (size_t)k * (m + 1);
      for (int j = 0; j <= k + 1; ++j) Hk[j] = hcol[j];
      for (int j = 0; j < k; ++j) {                                           // previous rotations
        const double t = cs[j] * Hk[j] + sn[j] * Hk[j + 1];
        Hk[j + 1] = -sn[j] * Hk[j] + cs[j] * Hk[j + 1];
        Hk[j] = t;
      }
      const double rr = std::hypot(Hk[k], Hk[k + 1]);
      cs[k] = rr == 0.0 ? 1.0 : Hk[k] / rr;
      sn[k] = rr == 0.0 ? 0.0 : Hk[k + 1] / rr;
      Hk[k] = rr;
      Hk[k + 1] = 0.0;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      ++its;
      kdone = k + 1;
      const double res = std::fabs(g[k + 1]);
      record(res);
      if (res <= target) {
        converged = true;
        break;
      }
    }
    y.assign(ALFIB_MAX_KRYLOV + 1, 0.0);
    for (int j = kdone - 1; j >= 0; --j) {                                    // back substitution
      double v = g[j];
      for (int k = j + 1; k < kdone; ++k) v -= H[(size_t)k * (m + 1) + j] * y[k];
      const double d = H[(size_t)j * (m + 1) + j];
      y[j] = d == 0.0 ? 0.0 : v / d;
    }
    CUDA_TRY(cudaMemcpyAsync(hd, y.data(), sizeof(double) * std::max(kdone, 1), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));                                       // y is a local: copied before it goes
    launch_maxpy_norm(c, n, kdone, hd, 1.0, Z, n, x, nullptr, nullptr);       // x += Z y
    if (converged) break;
    jacobian_apply_device(c, x, w);                                           // r = b - J x
    launch_sub(c, n, b, w, S.r.p);
    beta = norm_to_host(S.r.p);
    if (beta <= target) break;
  }
  if (iterations) *iterations = its;
  CUDA_TRY(cudaStreamSynchronize(s));
}
