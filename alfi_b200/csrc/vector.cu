// Small vector kernels used by the transfers, the smoother's boundary fix-up and the coarse solve.
#include "alfib_internal.h"

namespace {

__global__ void set_rows_kernel(double* __restrict__ y, const double* __restrict__ x,
                                const int32_t* __restrict__ idx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int k = idx[i];
    y[k] = x ? x[k] : 0.0;
  }
}

__global__ void axpby_kernel(int n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}

__global__ void sub_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}

// one warp per block row: scatter the row's blocks into a zeroed column-major dense matrix
template <int BS>
__global__ void bsr_to_dense_kernel(int nbrows, const int32_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                                    double* __restrict__ dense, int64_t n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nbrows) return;
  for (int k = rowptr[warp] + lane; k < rowptr[warp + 1]; k += 32) {
    const int64_t cn = colidx[k];
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
      for (int cc = 0; cc < BS; ++cc)
        dense[((int64_t)warp * BS + r) + (cn * BS + cc) * n] = vals[(int64_t)k * BS * BS + r * BS + cc];
  }
}

}  // namespace

void launch_set_rows(alfib_ctx* c, double* y, const double* x, const int32_t* idx, int nidx) {
  if (nidx == 0) return;
  set_rows_kernel<<<cdiv(nidx, 256), 256, 0, c->stream>>>(y, x, idx, nidx);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void launch_axpby(alfib_ctx* c, int n, double a, const double* x, double b, double* y) {
  if (n == 0) return;
  axpby_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(n, a, x, b, y);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void launch_sub(alfib_ctx* c, int n, const double* a, const double* b, double* out) {
  if (n == 0) return;
  sub_kernel<<<cdiv(n, 256), 256, 0, c->stream>>>(n, a, b, out);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void launch_bsr_to_dense(alfib_ctx* c, const Level& L, double* dense) {
  const int64_t n = L.n;
  CUDA_TRY(cudaMemsetAsync(dense, 0, (size_t)n * n * sizeof(double), c->stream));
  const int blocks = cdiv((int64_t)L.n_nodes * 32, 256);
  if (L.bs == 2)
    bsr_to_dense_kernel<2><<<blocks, 256, 0, c->stream>>>(L.n_nodes, L.rowptr.p, L.colidx.p, L.vals.p, dense, n);
  else
    bsr_to_dense_kernel<3><<<blocks, 256, 0, c->stream>>>(L.n_nodes, L.rowptr.p, L.colidx.p, L.vals.p, dense, n);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}
