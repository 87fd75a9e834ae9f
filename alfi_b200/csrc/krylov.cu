// Level smoother: FGMRES(m) exactly as alfi configures the mg_levels KSP (alfi/solver.py:313-317):
// right-preconditioned, classical Gram-Schmidt without refinement, `convergence_test skip`
// => exactly m iterations, nonzero initial guess.  Everything stays on the
// device — dots are two-pass (fixed-order, bitwise reproducible) reductions, the (m+1) x m
// Hessenberg least-squares problem is solved by one thread with Givens rotations — so a whole
// smoother call is a fixed sequence of launches with no host synchronisation.
//
// HBM-bound BLAS-1; algorithmic bytes per call ~ 8 N (m^2 + 8 m)  (SURVEY §8d).
#include <cstdlib>

#include "alfib_internal.h"

namespace {

constexpr int RT = 256;            // threads per reduction block
constexpr int RGRID = 592;         // fixed reduction grid (4 x 148 SMs) => fixed summation order
constexpr int MAXV = ALFIB_MAX_KRYLOV + 1;

// partial[j * RGRID + block] = sum over the block's slice of V_j[i] * w[i],  j < nv
__global__ void __launch_bounds__(RT) multi_dot_kernel(int n, int nv, const double* __restrict__ V, int64_t ldv,
                                                       const double* __restrict__ w, double* __restrict__ partial) {
  __shared__ double red[RT / 32];
  const int tid = threadIdx.x;
  for (int j0 = 0; j0 < nv; j0 += 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * RT + tid; i < n; i += (int64_t)RGRID * RT) {
      const double wi = w[i];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        if (j0 + jj < nv) acc[jj] = fma(V[(int64_t)(j0 + jj) * ldv + i], wi, acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (j0 + jj >= nv) break;
      double v = acc[jj];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red[tid >> 5] = v;
      __syncthreads();
      if (tid < 32) {
        v = tid < RT / 32 ? red[tid] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (tid == 0) partial[(int64_t)(j0 + jj) * RGRID + blockIdx.x] = v;
      }
      __syncthreads();
    }
  }
}

// Second pass of a dot by one warp: lane-strided sum of the RGRID partials, then the shuffle tree (valid in lane 0).
// All loads are issued before the first add — the plain loop `v += p[b]` is a chain of 19 dependent-latency loads,
// which made the one-warp finalize kernel the longest of the small kernels of a Krylov iteration (10-12 us in the
// launch list of round 2 against 3-4 us for an empty-ish kernel).  Same summation order as that loop.
__device__ __forceinline__ double warp_sum_partials(const double* __restrict__ p, int lane) {
  constexpr int K = (RGRID + 31) / 32;
  double a[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int b = lane + 32 * k;
    a[k] = b < RGRID ? p[b] : 0.0;
  }
  double v = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) v += a[k];
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// out[j] = op(sum_b partial[j*RGRID + b]);  one warp per j, fixed order.  mode 0: plain sum,
// mode 1: sqrt (norm).  Optionally also writes 1/out (0 if out == 0) to inv.
__global__ void finalize_kernel(int nv, const double* __restrict__ partial, double* __restrict__ out, int mode,
                                double* __restrict__ inv) {
  const int j = blockIdx.x, lane = threadIdx.x;
  if (j >= nv) return;
  double v = warp_sum_partials(partial + (int64_t)j * RGRID, lane);
  if (lane == 0) {
    if (mode == 1) v = sqrt(v);
    out[j] = v;
    if (inv) inv[j] = v > 0.0 ? 1.0 / v : 0.0;
  }
}

// w[i] += sign * sum_j coef[j] V_j[i];  if partial != nullptr also accumulates |w|^2 partials.
// hpart != nullptr: the coefficients are still first-pass partials of multi_dot_kernel (hpart[j * RGRID + block]); every
// block runs their second pass itself (warp j, the arithmetic of finalize_kernel: all blocks get the same bits) and
// block 0 stores them to hout — one launch and one dependent-latency kernel less per Gram-Schmidt step.  `partial`
// must not overlap hpart (other blocks may still be reading it).
__global__ void __launch_bounds__(RT) maxpy_kernel(int n, int nv, const double* __restrict__ coef,
                                                   const double* __restrict__ hpart, double* __restrict__ hout, double sign,
                                                   const double* __restrict__ V, int64_t ldv, double* __restrict__ w,
                                                   double* __restrict__ partial) {
  __shared__ double red[RT / 32];
  __shared__ double cf[MAXV];
  const int tid = threadIdx.x;
  if (hpart) {
    for (int j = tid >> 5; j < nv; j += RT / 32) {
      const double h = warp_sum_partials(hpart + (int64_t)j * RGRID, tid & 31);
      if ((tid & 31) == 0) {
        cf[j] = sign * h;
        if (blockIdx.x == 0) hout[j] = h;
      }
    }
  } else if (tid < nv) {
    cf[tid] = sign * coef[tid];
  }
  __syncthreads();
  double nrm = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RT + tid; i < n; i += (int64_t)RGRID * RT) {
    double v = w[i];
    for (int j = 0; j < nv; ++j) v = fma(cf[j], V[(int64_t)j * ldv + i], v);
    w[i] = v;
    nrm = fma(v, v, nrm);
  }
  if (partial) {
#pragma unroll
    for (int o = 16; o; o >>= 1) nrm += __shfl_down_sync(0xffffffffu, nrm, o);
    if ((tid & 31) == 0) red[tid >> 5] = nrm;
    __syncthreads();
    if (tid < 32) {
      nrm = tid < RT / 32 ? red[tid] : 0.0;
#pragma unroll
      for (int o = 16; o; o >>= 1) nrm += __shfl_down_sync(0xffffffffu, nrm, o);
      if (tid == 0) partial[blockIdx.x] = nrm;
    }
  }
}

// out[i] = scale[0] * in[i]
__global__ void scale_kernel(int n, const double* __restrict__ scale, const double* __restrict__ in,
                             double* __restrict__ out) {
  const double s = scale[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = s * in[i];
}

// out[i] = in[i] / nrm with nrm = sqrt(sum of the RGRID |.|^2 partials): the second pass of the norm (finalize_kernel,
// mode 1, same bits in every block) folded into the normalisation; block 0 stores nrm and 1 / nrm (0 if nrm = 0).
// n = 0 with one block: the norm alone (last Arnoldi step, whose basis vector nobody reads).
__global__ void __launch_bounds__(RT) scale_norm_kernel(int n, const double* __restrict__ npart, double* __restrict__ nrm_out,
                                                        double* __restrict__ inv_out, const double* __restrict__ in,
                                                        double* __restrict__ out) {
  __shared__ double s_inv;
  if (threadIdx.x < 32) {
    double v = warp_sum_partials(npart, threadIdx.x);
    if (threadIdx.x == 0) {
      v = sqrt(v);
      const double inv = v > 0.0 ? 1.0 / v : 0.0;
      s_inv = inv;
      if (blockIdx.x == 0) {
        nrm_out[0] = v;
        inv_out[0] = inv;
      }
    }
  }
  __syncthreads();
  const double s = s_inv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = s * in[i];
}

// Solve min || beta e1 - H y || for the (m+1) x m Hessenberg matrix (column-major, ld = MAXV)
// by Givens rotations — the update PETSc's FGMRES performs.  A zero pivot (happy breakdown)
// gives y_j = 0.
__global__ void hessenberg_solve_kernel(int m, const double* __restrict__ H, const double* __restrict__ beta,
                                        double* __restrict__ y) {
  // the (m+1) x m matrix is staged in shared memory by the whole warp: the sequential update below is a chain of
  // dependent accesses, which in global memory cost a load latency each (18 us per call in the launch list of round 2)
  __shared__ double Hs[MAXV * MAXV];
  __shared__ double ys[MAXV];
  for (int idx = threadIdx.x; idx < (m + 1) * m; idx += blockDim.x) {
    const int col = idx / (m + 1), row = idx - col * (m + 1);
    Hs[row + col * MAXV] = H[row + col * MAXV];
  }
  __syncthreads();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double g[MAXV];
  for (int i = 0; i <= m; ++i) g[i] = 0.0;
  g[0] = beta[0];
  for (int j = 0; j < m; ++j) {
    const double a = Hs[j + j * MAXV], b = Hs[j + 1 + j * MAXV];
    const double r = hypot(a, b);
    const double c = r == 0.0 ? 1.0 : a / r, s = r == 0.0 ? 0.0 : b / r;
    for (int col = j; col < m; ++col) {
      const double t0 = Hs[j + col * MAXV], t1 = Hs[j + 1 + col * MAXV];
      Hs[j + col * MAXV] = c * t0 + s * t1;
      Hs[j + 1 + col * MAXV] = -s * t0 + c * t1;
    }
    const double g0 = g[j], g1 = g[j + 1];
    g[j] = c * g0 + s * g1;
    g[j + 1] = -s * g0 + c * g1;
  }
  for (int j = m - 1; j >= 0; --j) {
    double v = g[j];
    for (int k = j + 1; k < m; ++k) v -= Hs[j + k * MAXV] * ys[k];
    const double d = Hs[j + j * MAXV];
    ys[j] = d == 0.0 ? 0.0 : v / d;
    y[j] = ys[j];
  }
}

}  // namespace

// ---- the same BLAS-1 pieces for the outer solver (outer.cu): fixed-order two-pass reductions over n entries --------
void krylov_reserve(alfib_ctx* c) {
  c->partial.alloc((size_t)(MAXV + 1) * RGRID);        // MAXV simultaneous dots + the norm partials of the fused update
  c->scal.alloc(MAXV * MAXV + 2 + 2 * MAXV);
}

// out[j] = <V_j, w>, j < nv <= ALFIB_MAX_KRYLOV + 1 (device pointers)
void launch_multi_dot(alfib_ctx* c, int n, int nv, const double* V, int64_t ldv, const double* w, double* out) {
  ALFIB_REQUIRE(nv >= 1 && nv <= MAXV, "too many simultaneous dots");
  krylov_reserve(c);
  multi_dot_kernel<<<RGRID, RT, 0, c->stream>>>(n, nv, V, ldv, w, c->partial.p);
  finalize_kernel<<<nv, 32, 0, c->stream>>>(nv, c->partial.p, out, 0, nullptr);
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
}

// w += sign * sum_j coef[j] V_j ; if nrm: nrm[0] = |w| afterwards, inv[0] = 1 / |w| (0 if w = 0)
void launch_maxpy_norm(alfib_ctx* c, int n, int nv, const double* coef, double sign, const double* V, int64_t ldv,
                       double* w, double* nrm, double* inv) {
  ALFIB_REQUIRE(nv >= 0 && nv <= MAXV, "too many vectors in one update");
  krylov_reserve(c);
  maxpy_kernel<<<RGRID, RT, 0, c->stream>>>(n, nv, coef, nullptr, nullptr, sign, V, ldv, w, nrm ? c->partial.p : nullptr);
  c->launches += 1;
  if (nrm) {
    finalize_kernel<<<1, 32, 0, c->stream>>>(1, c->partial.p, nrm, 1, inv);
    c->launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
}

// out = scale[0] * in
void launch_scale_by(alfib_ctx* c, int n, const double* scale, const double* in, double* out) {
  scale_kernel<<<RGRID, RT, 0, c->stream>>>(n, scale, in, out);
  c->launches += 1;
  CUDA_TRY(cudaGetLastError());
}

// scal layout: [0, MAXV*MAXV) H ; then beta, inv, y[MAXV], h[MAXV]
void fgmres_device(alfib_ctx* c, Level& L, int level, int m, const double* b, double* x) {
  ALFIB_REQUIRE(m >= 1 && m <= ALFIB_MAX_KRYLOV, "smoothing iterations out of range");
  // A level with a halo (alfib_level_set_halo) holds local vectors: n entries per basis vector, of which the
  // first `no` are owned.  BLAS-1 runs over the owned entries, every dot / norm is completed by one small
  // all-reduce (KSPFGMRES on an MPI Vec: VecMDot / VecNorm), the ghosts a gather reads are refreshed by the
  // consumer (patch_apply_sum, launch_bsr_spmv).
  const int n = L.n;
  const int no = L.halo.on ? L.n_owned : L.n;
  const bool dist = L.halo.on && c->nranks > 1;
  if (L.krylov_m < m) {
    L.V.alloc((size_t)(m + 1) * n);
    L.Z.alloc((size_t)m * n);
    if (L.halo.on) {                 // ghost parts are read by the exchanges' pack kernels before they are written
      CUDA_TRY(cudaMemsetAsync(L.V.p, 0, sizeof(double) * (size_t)(m + 1) * n, c->stream));
      CUDA_TRY(cudaMemsetAsync(L.Z.p, 0, sizeof(double) * (size_t)m * n, c->stream));
    }
    L.krylov_m = m;
  }
  L.w.alloc(n);
  krylov_reserve(c);
  double* H = c->scal.p;
  double* beta = H + MAXV * MAXV;
  double* inv = beta + 1;
  double* y = inv + 1;
  double* V = L.V.p;
  double* Z = L.Z.p;
  double* w = L.w.p;
  cudaStream_t s = c->stream;
  CUDA_TRY(cudaMemsetAsync(H, 0, sizeof(double) * MAXV * MAXV, s));

  // out[j] = <V_j, w> over all ranks, j < nv
  auto mdot = [&](int nv, const double* Vp, const double* wp, double* out) {
    multi_dot_kernel<<<RGRID, RT, 0, s>>>(no, nv, Vp, n, wp, c->partial.p);
    c->launches += 1;
    if (dist && comm_small_allreduce_partials(c, c->partial.p, RGRID, out, nv, 0, nullptr)) return;   // second pass + exchange: one kernel
    finalize_kernel<<<nv, 32, 0, s>>>(nv, c->partial.p, out, 0, nullptr);
    c->launches += 1;
    if (dist) comm_small_allreduce(c, out, nv, 0, nullptr);
  };
  // out = sqrt(sum of the |.|^2 partials over all ranks), invp = 1 / out
  auto norm_of_partials = [&](double* out, double* invp) {
    if (!dist) {
      finalize_kernel<<<1, 32, 0, s>>>(1, c->partial.p, out, 1, invp);
      c->launches += 1;
    } else {
      if (comm_small_allreduce_partials(c, c->partial.p, RGRID, out, 1, 1, invp)) return;
      finalize_kernel<<<1, 32, 0, s>>>(1, c->partial.p, out, 0, nullptr);
      c->launches += 1;
      comm_small_allreduce(c, out, 1, 1, invp);
    }
  };

  // Single rank (default; ALFIB_FUSE_DOTS=0 keeps the separate second-pass kernels): the second pass of every dot / norm
  // runs inside its consumer — h inside the Gram-Schmidt update, the norm inside the normalisation — so an Arnoldi
  // step is 3 BLAS-1 launches instead of 5, and the last basis vector (never read) is not written.
  static const bool fuse_env = !(std::getenv("ALFIB_FUSE_DOTS") && std::getenv("ALFIB_FUSE_DOTS")[0] == '0');
  const bool fused = fuse_env && !dist;
  double* npart = c->partial.p + (size_t)MAXV * RGRID;       // norm partials of the fused update (not the dots' rows)

  // r0 = b - A x ; beta = |r0| ; v0 = r0 / beta
  {
    ScopedEvent ev(c, ALFIB_EV_MATMULT, level);
    launch_bsr_spmv(c, L, L.vals.p, x, w, b);
  }
  {
    ScopedEvent ev(c, ALFIB_EV_KSP_GMRES_ORTHOG, level);
    multi_dot_kernel<<<RGRID, RT, 0, s>>>(no, 1, w, n, w, c->partial.p);
    c->launches += 1;
    if (fused) {
      scale_norm_kernel<<<RGRID, RT, 0, s>>>(no, c->partial.p, beta, inv, w, V);
    } else {
      norm_of_partials(beta, inv);
      scale_kernel<<<RGRID, RT, 0, s>>>(no, inv, w, V);
    }
    c->launches += 1;
  }
  for (int k = 0; k < m; ++k) {
    double* vk = V + (size_t)k * n;
    double* zk = Z + (size_t)k * n;
    smoother_apply_device(c, L, level, vk, zk);                      // z_k = M^-1 v_k
    {
      ScopedEvent ev(c, ALFIB_EV_MATMULT, level);
      launch_bsr_spmv(c, L, L.vals.p, zk, w, nullptr);        // w = A z_k
    }
    ScopedEvent ev(c, ALFIB_EV_KSP_GMRES_ORTHOG, level);
    double* hcol = H + (size_t)k * MAXV;
    if (fused) {
      multi_dot_kernel<<<RGRID, RT, 0, s>>>(no, k + 1, V, n, w, c->partial.p);                           // h = V^T w  (CGS), first pass
      maxpy_kernel<<<RGRID, RT, 0, s>>>(no, k + 1, nullptr, c->partial.p, hcol, -1.0, V, n, w, npart);   // h; w -= V h; |w|^2
      if (k + 1 < m)
        scale_norm_kernel<<<RGRID, RT, 0, s>>>(no, npart, hcol + k + 1, inv, w, V + (size_t)(k + 1) * n);
      else
        scale_norm_kernel<<<1, RT, 0, s>>>(0, npart, hcol + k + 1, inv, nullptr, nullptr);                          // H[m][m-1] only
      c->launches += 3;
      continue;
    }
    mdot(k + 1, V, w, hcol);                                                          // h = V^T w  (CGS)
    maxpy_kernel<<<RGRID, RT, 0, s>>>(no, k + 1, hcol, nullptr, nullptr, -1.0, V, n, w, c->partial.p);   // w -= V h, |w|^2
    c->launches += 1;
    norm_of_partials(hcol + k + 1, inv);
    scale_kernel<<<RGRID, RT, 0, s>>>(no, inv, w, V + (size_t)(k + 1) * n);
    c->launches += 1;
  }
  ScopedEvent ev(c, ALFIB_EV_KSP_GMRES_ORTHOG, level);
  hessenberg_solve_kernel<<<1, 32, 0, s>>>(m, H, beta, y);
  maxpy_kernel<<<RGRID, RT, 0, s>>>(no, m, y, nullptr, nullptr, 1.0, Z, n, x, nullptr);          // x += Z y
  c->launches += 2;
  CUDA_TRY(cudaGetLastError());
}
