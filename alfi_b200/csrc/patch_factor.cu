// Per-Newton-step patch setup (PCSetUp_PATCH): gather A_i = A[I_i, I_i] from the level's BSR
// values and invert it — `patch_pc_patch_save_operators`, `patch_sub_pc_type lu` and
// `patch_pc_patch_dense_inverse` of alfi/solver.py:320,327,599-602 (the reference does this
// with LAPACK getrf+getri through PETSc).
//
// One persistent CTA per SM takes patches from an atomic counter (largest first).  The patch
// matrix lives column-major in a per-CTA global workspace slot (13 MB for n = 1275, so L2/HBM
// resident) and is inverted in place by *blocked Gauss-Jordan with partial row pivoting*:
//   for each panel K of NB columns
//     1. panel -> shared memory; NB unblocked Gauss-Jordan steps on all n rows of the panel,
//        pivot = first max |.| among the not-yet-pivoted rows (LAPACK idamax rule);
//     2. the NB row swaps are applied to every other column, R = W[K, :] is set aside and
//        W[K, :] zeroed;
//     3. rank-NB update W[:, J] += N * R with N = the transformed panel (one row per thread,
//        N[r, 0:NB] in registers, R staged through shared memory 64 columns at a time).
//   A^{-1} = W with the column swaps undone in reverse order; that permutation is folded into
//   the final pass that writes the 64-row-tiled apply layout (patch_apply.cu).
// Flops 2 n^3 per patch; the update is the only O(n^3) part and streams W once per panel.
#include <algorithm>
#include <climits>
#include <numeric>

#include "alfib_internal.h"

namespace {

constexpr int FT = 512;        // threads per CTA
constexpr int CC = 64;         // columns per staged R chunk
constexpr int UC = 8;          // columns whose loads are issued together in the update

struct FactorArgs {
  int npatch;
  const int32_t* forder;       // patches, largest first
  const int64_t* poff;
  const int32_t* pdofs;
  const int32_t* sorted;       // patch dofs sorted ascending ...
  const int32_t* sperm;        // ... and their local indices
  const int64_t* soff;
  double* store;
  // level operator
  int bs;
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  // workspace
  double* work;
  int64_t slot_elems;          // per CTA: W (maxn*ld) + R (maxn*NB)
  int maxn;
  int* counter;
  int* info;                   // 0 or (1 + index of a singular patch)
};

__device__ __forceinline__ int lookup(const int32_t* sd, int n, int key) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = sd[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

template <int NB>
__global__ void __launch_bounds__(FT, 1) patch_factor_kernel(FactorArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int ldp_max = (a.maxn + 1) & ~1;
  // shared memory carve-up
  double* panel = reinterpret_cast<double*>(smem_raw);                 // ldp_max * NB
  double* Rs = panel + (size_t)ldp_max * NB;                           // CC * NB
  double* pr = Rs + CC * NB;                                           // NB
  double* red_v = pr + NB;                                             // 33
  int* red_i = reinterpret_cast<int*>(red_v + 34);                     // 34
  int* ipiv = red_i + 34;                                              // maxn
  int* s_next = ipiv + a.maxn;                                         // 1
  // the gather's lookup tables alias the panel (used before the factorisation starts)
  int* sd = reinterpret_cast<int*>(panel);
  int* sp = sd + a.maxn;

  double* W = a.work + (size_t)blockIdx.x * a.slot_elems;
  double* Rg = W + (size_t)a.maxn * ldp_max;

  for (;;) {
    if (tid == 0) *s_next = atomicAdd(a.counter, 1);
    __syncthreads();
    const int q = *s_next;
    __syncthreads();
    if (q >= a.npatch) break;
    const int p = a.forder[q];
    const int64_t o = a.poff[p];
    const int n = (int)(a.poff[p + 1] - o);
    if (n == 0) continue;
    const int ld = (n + 1) & ~1;
    const int ldp = ld;
    const int32_t* I = a.pdofs + o;

    // ---- gather A[I, I] into W (column-major) ----------------------------------------------
    for (int64_t i = tid; i < (int64_t)n * ld; i += FT) W[i] = 0.0;
    for (int i = tid; i < n; i += FT) {
      sd[i] = a.sorted[o + i];
      sp[i] = a.sperm[o + i];
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31, bs = a.bs, b2 = bs * bs;
      for (int r = warp; r < n; r += FT / 32) {
        const int g = I[r];
        const int node = g / bs, comp = g - node * bs;
        const int k1 = a.rowptr[node + 1];
        for (int k = a.rowptr[node] + lane; k < k1; k += 32) {
          const int cn = a.colidx[k];
          for (int c2 = 0; c2 < bs; ++c2) {
            const int hit = lookup(sd, n, cn * bs + c2);
            if (hit >= 0) W[r + (size_t)sp[hit] * ld] = a.vals[(int64_t)k * b2 + comp * bs + c2];
          }
        }
      }
    }
    __syncthreads();

    // ---- blocked Gauss-Jordan ---------------------------------------------------------------
    for (int k0 = 0; k0 < n; k0 += NB) {
      const int nb = (n - k0) < NB ? (n - k0) : NB;
      // 1. panel to shared memory
      for (int c = 0; c < nb; ++c)
        for (int r = tid; r < n; r += FT) panel[r + c * ldp] = W[r + (size_t)(k0 + c) * ld];
      __syncthreads();
      for (int j = 0; j < nb; ++j) {
        const int kj = k0 + j;
        double best = -1.0;
        int bi = INT_MAX;
        for (int r = kj + tid; r < n; r += FT) {
          const double v = fabs(panel[r + j * ldp]);
          if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
          const double ov = __shfl_down_sync(0xffffffffu, best, off);
          const int oi = __shfl_down_sync(0xffffffffu, bi, off);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid < 32) {
          best = tid < FT / 32 ? red_v[tid] : -1.0;
          bi = tid < FT / 32 ? red_i[tid] : INT_MAX;
#pragma unroll
          for (int off = 16; off; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          if (tid == 0) {
            if (!(best > 0.0)) { bi = kj; atomicCAS(a.info, 0, p + 1); }
            red_i[32] = bi;
            red_v[32] = best;
            ipiv[kj] = bi;
          }
        }
        __syncthreads();
        const int pv = red_i[32];
        const bool singular = !(red_v[32] > 0.0);
        if (tid < nb && pv != kj) {
          const double t0 = panel[kj + tid * ldp];
          panel[kj + tid * ldp] = panel[pv + tid * ldp];
          panel[pv + tid * ldp] = t0;
        }
        __syncthreads();
        const double d = singular ? 0.0 : 1.0 / panel[kj + j * ldp];
        if (tid < NB) pr[tid] = (tid == j || tid >= nb) ? 0.0 : panel[kj + tid * ldp] * d;
        __syncthreads();
        for (int r = tid; r < n; r += FT) {
          if (r == kj) {
#pragma unroll
            for (int jj = 0; jj < NB; ++jj)
              if (jj < nb) panel[kj + jj * ldp] = (jj == j) ? d : pr[jj];
          } else {
            const double f = panel[r + j * ldp];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj)
              if (jj < nb) panel[r + jj * ldp] = fma(-f, pr[jj], panel[r + jj * ldp]);
            panel[r + j * ldp] = -f * d;
          }
        }
        __syncthreads();
      }
      // 2. row swaps on the other columns, set R aside, zero the pivot rows
      for (int c = tid; c < n; c += FT) {
        if (c >= k0 && c < k0 + nb) continue;
        double* col = W + (size_t)c * ld;
        for (int j = 0; j < nb; ++j) {
          const int kj = k0 + j, pv = ipiv[kj];
          if (pv != kj) {
            const double t0 = col[kj];
            col[kj] = col[pv];
            col[pv] = t0;
          }
        }
#pragma unroll
        for (int t = 0; t < NB; ++t) {
          double v = 0.0;
          if (t < nb) { v = col[k0 + t]; col[k0 + t] = 0.0; }
          Rg[(size_t)c * NB + t] = v;
        }
      }
      __syncthreads();
      // 3. rank-NB update of all other columns, R staged CC columns at a time
      for (int cc0 = 0; cc0 < n; cc0 += CC) {
        const int ccn = (n - cc0) < CC ? (n - cc0) : CC;
        for (int i = tid; i < ccn * NB; i += FT) Rs[i] = Rg[(size_t)cc0 * NB + i];
        __syncthreads();
        for (int r = tid; r < n; r += FT) {
          double N[NB];
#pragma unroll
          for (int t = 0; t < NB; ++t) N[t] = (t < nb) ? panel[r + t * ldp] : 0.0;
          double* Wr = W + r;
          // UC columns at a time: all loads first (memory-level parallelism — with one load in
          // flight per thread the update ran at ~7 GB/s per SM), then the FMAs, then the stores
          for (int ci0 = 0; ci0 < ccn; ci0 += UC) {
            double w[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const int cg = cc0 + ci0 + u;
              const bool ok = (ci0 + u < ccn) && !(cg >= k0 && cg < k0 + nb);
              w[u] = ok ? Wr[(size_t)cg * ld] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const double* __restrict__ rs = Rs + (ci0 + u) * NB;
#pragma unroll
              for (int t = 0; t < NB; ++t) w[u] = fma(N[t], rs[t], w[u]);
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const int cg = cc0 + ci0 + u;
              if ((ci0 + u < ccn) && !(cg >= k0 && cg < k0 + nb)) Wr[(size_t)cg * ld] = w[u];
            }
          }
        }
        __syncthreads();
      }
      // panel back
      for (int c = 0; c < nb; ++c)
        for (int r = tid; r < n; r += FT) W[r + (size_t)(k0 + c) * ld] = panel[r + c * ldp];
      __syncthreads();
    }

    // ---- undo the pivoting (column swaps in reverse) and write the tiled apply layout --------
    int* src = sd;               // the panel region is free again
    if (tid == 0) {
      for (int c = 0; c < n; ++c) src[c] = c;
      for (int k = n - 1; k >= 0; --k) {
        const int pv = ipiv[k];
        if (pv != k) { const int t0 = src[k]; src[k] = src[pv]; src[pv] = t0; }
      }
    }
    __syncthreads();
    {
      double* out = a.store + a.soff[p];
      const int ntile = (n + ALFIB_TILE_ROWS - 1) / ALFIB_TILE_ROWS;
      for (int t = 0; t < ntile; ++t) {
        const int row0 = t * ALFIB_TILE_ROWS;
        int rows = n - row0;
        rows = rows > ALFIB_TILE_ROWS ? ALFIB_TILE_ROWS : ((rows + 1) & ~1);
        double* tile = out + (size_t)row0 * n;
        for (int64_t i = tid; i < (int64_t)rows * n; i += FT) {
          const int c = (int)(i / rows), rr = (int)(i - (int64_t)c * rows);
          const int r = row0 + rr;
          tile[i] = (r < n) ? W[r + (size_t)src[c] * ld] : 0.0;
        }
      }
    }
    __syncthreads();
  }
}

size_t factor_smem_bytes(int maxn, int NB) {
  const size_t ldp = (maxn + 1) & ~1;
  size_t doubles = ldp * NB + CC * NB + NB + 34;
  size_t ints = 34 + maxn + 2;
  size_t panel_bytes = ldp * NB * sizeof(double);
  size_t bytes = doubles * sizeof(double) + ints * sizeof(int);
  // lookup tables alias the panel: need 2*maxn ints inside it
  if (panel_bytes < 2 * (size_t)maxn * sizeof(int)) bytes += 2 * (size_t)maxn * sizeof(int) - panel_bytes;
  return bytes + 16;
}

template <int NB>
void run_factor(alfib_ctx* c, FactorArgs a, size_t smem, int grid) {
  CUDA_TRY(cudaFuncSetAttribute(patch_factor_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  patch_factor_kernel<NB><<<grid, FT, smem, c->stream>>>(a);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

}  // namespace

void launch_patch_factor(alfib_ctx* c, const Level& L, PatchSet& ps, const double* vals) {
  if (ps.npatch == 0 || ps.maxn == 0) { ps.factored = true; return; }
  const int maxn = ps.maxn;
  const size_t budget = 200 * 1024;
  int NB = 32;
  while (NB > 4 && factor_smem_bytes(maxn, NB) > budget) NB >>= 1;
  if (factor_smem_bytes(maxn, NB) > budget)
    throw DeviceError{ALFIB_EINVAL, "patch too large for the shared-memory panel (n = " + std::to_string(maxn) + ")"};
  const size_t smem = factor_smem_bytes(maxn, NB);
  const int ldmax = roundup2(maxn);
  const int64_t slot = (int64_t)maxn * ldmax + (int64_t)maxn * NB;
  // resident CTAs: one per SM for large panels, more when shared memory allows
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
  int grid = std::min(ps.npatch, c->num_sms * per_sm);
  c->fwork.alloc((size_t)grid * slot);
  c->finfo.alloc(2);
  CUDA_TRY(cudaMemsetAsync(c->finfo.p, 0, 2 * sizeof(int), c->stream));

  FactorArgs a;
  a.npatch = ps.npatch;
  a.forder = ps.forder.p;
  a.poff = ps.off.p;
  a.pdofs = ps.dofs.p;
  a.sorted = ps.sorted.p;
  a.sperm = ps.sperm.p;
  a.soff = ps.soff.p;
  a.store = ps.store;
  a.bs = L.bs;
  a.rowptr = L.rowptr.p;
  a.colidx = L.colidx.p;
  a.vals = vals;
  a.work = c->fwork.p;
  a.slot_elems = slot;
  a.maxn = maxn;
  a.counter = c->finfo.p + 1;
  a.info = c->finfo.p;
  switch (NB) {
    case 32: run_factor<32>(c, a, smem, grid); break;
    case 16: run_factor<16>(c, a, smem, grid); break;
    case 8: run_factor<8>(c, a, smem, grid); break;
    default: run_factor<4>(c, a, smem, grid); break;
  }
  int info[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(info, c->finfo.p, sizeof(info), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (info[0] != 0)
    throw DeviceError{ALFIB_ESINGULAR, "patch " + std::to_string(info[0] - 1) + " is singular"};
  ps.factored = true;
}

void patch_extract_inverse(alfib_ctx* c, const PatchSet& ps, int patch, double* host_out) {
  ALFIB_REQUIRE(patch >= 0 && patch < ps.npatch, "patch index out of range");
  ALFIB_REQUIRE(ps.factored, "patches not factored");
  const int n = (int)(ps.h_off[patch + 1] - ps.h_off[patch]);
  if (n == 0) return;
  const int64_t elems = (int64_t)n * roundup2(n);
  std::vector<double> tmp(elems);
  CUDA_TRY(cudaMemcpyAsync(tmp.data(), ps.store + ps.h_soff[patch], elems * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int ntile = (n + ALFIB_TILE_ROWS - 1) / ALFIB_TILE_ROWS;
  for (int t = 0; t < ntile; ++t) {
    const int row0 = t * ALFIB_TILE_ROWS;
    int rows = n - row0;
    rows = rows > ALFIB_TILE_ROWS ? ALFIB_TILE_ROWS : roundup2(rows);
    const double* tile = tmp.data() + (int64_t)row0 * n;
    for (int cidx = 0; cidx < n; ++cidx)
      for (int rr = 0; rr < rows && row0 + rr < n; ++rr)
        host_out[(int64_t)(row0 + rr) * n + cidx] = tile[(int64_t)cidx * rows + rr];
  }
}
