// Sparse kernels: BSR residual/SpMV of the level operator, scalar-CSR (x) I_bs prolongation.
//
// Reference: PETSc MatMult_SeqBAIJ on the `baij` velocity block (alfi/solver.py:512) and the
// firedrake.mg prolong/restrict kernels behind `standard_transfer` (alfi/transfer.py:284-290).
//
// BSR SpMV: half a warp per block row (kernel below).  HBM-bound: algorithmic bytes
// nnzb*(8 bs^2 + 4) + 4(nbrows+1) + 16 N  (SURVEY §8d).  Measured (ncu, profiles/): 2.1x faster
// than a warp-per-row kernel walking the row's flat value run with 128-bit loads, because the
// 13 independent loads per lane hide the colidx -> x dependency; a 72-byte 3x3 block cannot be
// 16-byte aligned for every block, so the values are read as 8-byte loads that coalesce in L1.
#include <cstdlib>

#include "alfib_internal.h"

namespace {

// 16 lanes per block row, one whole bs x bs block per lane and iteration.  All loads of
// an iteration (1 column index, bs x-entries, bs*bs values) are independent except colidx -> x,
// so every lane keeps 1 + bs + bs*bs requests in flight; the strided 8-byte value loads of a
// half-warp cover one contiguous 16*bs*bs*8-byte run and coalesce in L1.
template <int BS>
__global__ void __launch_bounds__(256) bsr_spmv_block_kernel(int row0, int nbrows, const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const double* __restrict__ vals,
                                                             const double* __restrict__ x, PeerOut yout,
                                                             const double* __restrict__ b) {
  constexpr int B2 = BS * BS;
  double* __restrict__ y = resolve(yout);
  const int row = row0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
  const int l16 = threadIdx.x & 15;
  double acc[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) acc[r] = 0.0;
  if (row < nbrows) {
    const int k1 = __ldg(rowptr + row + 1);
#pragma unroll 2
    for (int k = __ldg(rowptr + row) + l16; k < k1; k += 16) {
      const double* __restrict__ v = vals + (int64_t)k * B2;
      const double* __restrict__ xc = x + (int64_t)__ldg(colidx + k) * BS;
      double vv[B2], xx[BS];
#pragma unroll
      for (int i = 0; i < B2; ++i) vv[i] = __ldg(v + i);
#pragma unroll
      for (int i = 0; i < BS; ++i) xx[i] = __ldg(xc + i);
#pragma unroll
      for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int cc = 0; cc < BS; ++cc) acc[r] = fma(vv[r * BS + cc], xx[cc], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int o = 8; o; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o, 16);
  if (row < nbrows && l16 == 0) {
#pragma unroll
    for (int r = 0; r < BS; ++r) {
      const int64_t i = (int64_t)row * BS + r;
      y[i] = b ? b[i] - acc[r] : acc[r];
    }
  }
}

// y[node, :] = sum_k vals[k] * x[col[k], :]   (scalar CSR acting on bs-interleaved vectors)
template <int BS>
__global__ void csr_apply_kernel(int nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const double* __restrict__ vals, const double* __restrict__ x,
                                 double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double acc[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) acc[r] = 0.0;
  for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
    const double v = vals[k];
    const double* xc = x + (int64_t)colidx[k] * BS;
#pragma unroll
    for (int r = 0; r < BS; ++r) acc[r] = fma(v, xc[r], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < BS; ++r) y[(int64_t)i * BS + r] = acc[r];
}

// Long rows (P_H^T on the coarser levels has hundreds of entries per row and few rows): one warp per row,
// lanes stride over the entries, fixed-order shuffle reduction (deterministic).
template <int BS>
__global__ void csr_apply_warp_kernel(int nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                      const double* __restrict__ vals, const double* __restrict__ x,
                                      double* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  double acc[BS];
#pragma unroll
  for (int r = 0; r < BS; ++r) acc[r] = 0.0;
  const int k1 = __ldg(rowptr + row + 1);
  for (int k = __ldg(rowptr + row) + lane; k < k1; k += 32) {
    const double v = __ldg(vals + k);
    const double* __restrict__ xc = x + (int64_t)__ldg(colidx + k) * BS;
#pragma unroll
    for (int r = 0; r < BS; ++r) acc[r] = fma(v, __ldg(xc + r), acc[r]);
  }
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < BS; ++r) y[(int64_t)row * BS + r] = acc[r];
  }
}

template <int BS>
__global__ void transpose_blocks_kernel(double* vals, int64_t nnzb) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnzb) return;
  double* v = vals + k * BS * BS;
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int c = r + 1; c < BS; ++c) {
      const double t = v[r * BS + c];
      v[r * BS + c] = v[c * BS + r];
      v[c * BS + r] = t;
    }
}

}  // namespace

void launch_bsr_spmv(alfib_ctx* c, const Level& L, const double* vals, const double* x, double* y,
                     const double* b) {
  const int threads = 256;
  // multi-GPU: this rank computes its own block rows, then the owned rows are broadcast
  // distributed vectors (alfib_level_set_halo): refresh the ghosts of x, then the owned block rows only —
  // x is a local work vector there, its ghost part is scratch
  const bool sharded = c->nranks > 1 && !L.row_start.empty() && !L.halo.on;
  if (L.halo.on) halo_update(c, const_cast<Level&>(L).halo, const_cast<double*>(x), L.index);
  const int row0 = sharded ? (int)L.row_start[c->rank] : 0;
  const int row1 = sharded ? (int)L.row_start[c->rank + 1] : (L.halo.on ? L.n_owned / L.bs : L.n_nodes);
  static const bool debug_skip_kernel = std::getenv("ALFIB_DEBUG_SKIP_SPMV") != nullptr;   // timing experiments only
  const int blocks = debug_skip_kernel ? 0 : cdiv((int64_t)(row1 - row0) * 16, threads);
  const bool peer = sharded && c->peers_open;
  const PeerOut out = peer ? comm_peer_out(c) : plain_out(y);
  if (blocks > 0) {
    if (L.bs == 2)
      bsr_spmv_block_kernel<2><<<blocks, threads, 0, c->stream>>>(row0, row1, L.rowptr.p, L.colidx.p, vals, x, out, b);
    else if (L.bs == 3)
      bsr_spmv_block_kernel<3><<<blocks, threads, 0, c->stream>>>(row0, row1, L.rowptr.p, L.colidx.p, vals, x, out, b);
    else
      throw DeviceError{ALFIB_EINVAL, "block size must be 2 or 3"};
    c->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if (peer) {                        // pull every rank's rows from its symmetric slot (NVLink peer loads)
    long long lo[ALFIB_MAX_RANKS], hi[ALFIB_MAX_RANKS];
    for (int r = 0; r < c->nranks; ++r) { lo[r] = L.dof_start[r]; hi[r] = L.dof_start[r + 1]; }
    comm_peer_reduce(c, L.n, -1, lo, hi, y);
  } else if (sharded) {
    comm_allgather_rows(c, y, L.dof_start);
  }
}

namespace {
// r[row] = x[row] - (A y)[row] for the block rows of a list; one warp per block row, lanes over its blocks
template <int BS>
__global__ void __launch_bounds__(256) bsr_residual_rows_kernel(int nrows, const int32_t* __restrict__ rows,
                                                                const int32_t* __restrict__ rowptr,
                                                                const int32_t* __restrict__ colidx,
                                                                const double* __restrict__ vals, const double* __restrict__ x,
                                                                const double* __restrict__ y, double* __restrict__ r) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nrows) return;
  const int row = rows[w];
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) acc[i] = 0.0;
  for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
    const double* __restrict__ B = vals + (int64_t)k * BS * BS;
    const double* __restrict__ yc = y + (int64_t)colidx[k] * BS;
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) acc[i] = fma(B[i * BS + j], yc[j], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < BS; ++i)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (lane < BS) r[(int64_t)row * BS + lane] = x[(int64_t)row * BS + lane] - acc[lane];
}
}  // namespace

void launch_bsr_residual_rows(alfib_ctx* c, const Level& L, const double* vals, const int32_t* rows, int nrows,
                              const double* x, const double* y, double* r) {
  if (nrows <= 0) return;
  const int blocks = cdiv((int64_t)nrows * 32, 256);
  if (L.bs == 2)
    bsr_residual_rows_kernel<2><<<blocks, 256, 0, c->stream>>>(nrows, rows, L.rowptr.p, L.colidx.p, vals, x, y, r);
  else
    bsr_residual_rows_kernel<3><<<blocks, 256, 0, c->stream>>>(nrows, rows, L.rowptr.p, L.colidx.p, vals, x, y, r);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void launch_csr_apply(alfib_ctx* c, int nrows, int bs, const int32_t* rowptr, const int32_t* colidx,
                      const double* vals, const double* x, double* y, int64_t nnz) {
  const int threads = 256;
  if (nrows == 0) return;
  // mean row length >= 16: a warp per row (restriction on the coarser levels); otherwise a thread per row
  if (nnz >= (int64_t)16 * nrows) {
    const int wblocks = cdiv((int64_t)nrows * 32, threads);
    if (bs == 1)
      csr_apply_warp_kernel<1><<<wblocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
    else if (bs == 2)
      csr_apply_warp_kernel<2><<<wblocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
    else if (bs == 3)
      csr_apply_warp_kernel<3><<<wblocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
    else
      throw DeviceError{ALFIB_EINVAL, "block size must be 1, 2 or 3"};
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return;
  }
  const int blocks = cdiv(nrows, threads);
  if (blocks == 0) return;
  if (bs == 1)
    csr_apply_kernel<1><<<blocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
  else if (bs == 2)
    csr_apply_kernel<2><<<blocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
  else if (bs == 3)
    csr_apply_kernel<3><<<blocks, threads, 0, c->stream>>>(nrows, rowptr, colidx, vals, x, y);
  else
    throw DeviceError{ALFIB_EINVAL, "block size must be 1, 2 or 3"};
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

void launch_transpose_blocks(alfib_ctx* c, double* vals, int64_t nnzb, int bs) {
  const int threads = 256;
  const int blocks = cdiv(nnzb, threads);
  if (blocks == 0) return;
  if (bs == 2)
    transpose_blocks_kernel<2><<<blocks, threads, 0, c->stream>>>(vals, nnzb);
  else
    transpose_blocks_kernel<3><<<blocks, threads, 0, c->stream>>>(vals, nnzb);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}
