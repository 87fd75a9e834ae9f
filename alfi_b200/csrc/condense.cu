// Condensed patch inverses: the block/separator ("static condensation") form of A_i^{-1}.
//
// A macro-star patch of the Scott-Vogelius discretisation on a barycentrically refined mesh
// (alfi/relaxation.py:168-177, alfi/bary.py) is mostly *interior* dofs of its macro cells: in 3-D
// (k = 3) 24 cells x 45 dofs = 1080 of the 1275 patch dofs, and the interiors of different macro
// cells do not couple.  With the patch dofs split into decoupled blocks B_k and a separator S
// (everything else), and N_k the separator dofs block k couples to,
//
//     A^{-1} = blockdiag(D_k) + [-W; I] X_SS [-V, I],   D_k = A_kk^{-1}, W_k = D_k A_kN, V_k = A_Nk D_k,
//
// where X_SS = (A^{-1})[S, S] is the inverse of the Schur complement.  Stored per patch:
// X_SS (|S|^2), and per block V_k (|N_k| x |B_k|) and [D_k | -W_k] (|B_k| x (|B_k| + |N_k|)):
// 153 k doubles instead of 1.63 M for the interior 3-D patch.  PCApply_PATCH streams the factors
// once per application, so the apply is ~10x fewer HBM bytes.
//
// Numerics (measured in numpy before this was written): forming the Schur
// complement A_SS - sum_k A_Sk D_k A_kS in FP64 is *not* accurate enough for augmented-Lagrangian
// patch matrices (kappa(A_kk) ~ gamma/nu: 1e-3 relative error at gamma = 1e4, Re = 5000), whereas
// X_SS taken from the pivoted dense inverse of the whole patch together with FP64 D_k, V_k, W_k
// reproduces the dense-inverse apply to its own accuracy (~kappa * eps).  So the per-Newton-step
// setup still inverts the whole patch with the blocked Gauss-Jordan kernel (patch_factor.cu) in
// its workspace, but keeps only X_SS; the block factors are computed by `condense_blocks_kernel`.
//
// Apply = four flat kernels over all patches of the set (no CTA barriers, warps independent):
//   K1  tile ops   g1[q]   = V_q x[B_q]                              (one warp per block)
//   K2  sep rhs    rs[p,s] = x[S_p[s]] - sum_{q, j: N_q[j] = s} g1[q][j]   (fixed order)
//   K3  tile ops   us[p]   = X_SS rs[p];  y[S_p] += us[p]            (one warp per 64-row tile)
//   K4  tile ops   y[B_q] += [D_q | -W_q] [x[B_q]; us[p][N_q]]       (one warp per block)
// K3 and K4 scatter into y with atomics, or colour by colour with plain stores in deterministic
// mode, exactly like patch_apply.cu.  Roofline: HBM; algorithmic bytes = stored factors + indices.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <numeric>

#include "alfib_internal.h"

namespace {

// ------------------------------------------------------------------------------------------------
// generic tile op:  dst[rows] (+)= M * src[cols],  M column-major with roundup2(nrows) rows per column
// Lanes own row pairs (double2 loads); when a tile has <= 32 (<= 16) rows the warp is split into
// 2 (4) column groups so that all lanes carry loads, and the groups are summed with shuffles.
__device__ __forceinline__ double fetch_src(const int32_t* __restrict__ ci, int cpos, int n,
                                            const double* __restrict__ srcA, const double* __restrict__ srcB) {
  if (cpos >= n) return 0.0;
  const int e = __ldg(ci + cpos);
  return e >= 0 ? __ldg(srcA + e) : __ldg(srcB + (~e));
}

template <int G, bool ATOMIC, bool ACCUM>
__device__ __forceinline__ void tile_op_body(const double* __restrict__ mat, const int32_t* __restrict__ ci,
                                             const int32_t* __restrict__ ri, double* __restrict__ priv,
                                             const int nrows, const int n, const double* __restrict__ srcA,
                                             const double* __restrict__ srcB, double* __restrict__ y, int lane) {
  constexpr int LPG = 32 / G;                       // lanes per column group
  const int half = (nrows + 1) >> 1;                // double2 per column
  const int grp = lane / LPG, l = lane - grp * LPG;
  const bool active = l < half;
  const double2* __restrict__ T = reinterpret_cast<const double2*>(mat) + (active ? l : 0);
  double acc0 = 0.0, acc1 = 0.0;
  double xv = fetch_src(ci, lane, n, srcA, srcB);
  for (int c0 = 0; c0 < n; c0 += 32) {
    const double xnext = fetch_src(ci, c0 + 32 + lane, n, srcA, srcB);
    const int cnt = (n - c0) < 32 ? (n - c0) : 32;
    const double2* __restrict__ Tc = T + (int64_t)c0 * half;
    if (cnt == 32) {
#pragma unroll 8
      for (int jj = 0; jj < LPG; ++jj) {
        const int j = jj * G + grp;
        const double xc = __shfl_sync(0xffffffffu, xv, j);
        if (active) {
          const double2 a = __ldcs(Tc + (int64_t)j * half);
          acc0 = fma(a.x, xc, acc0);
          acc1 = fma(a.y, xc, acc1);
        }
      }
    } else {
      for (int jj = 0; jj * G < cnt; ++jj) {
        const int j = jj * G + grp;
        const double xc = __shfl_sync(0xffffffffu, xv, j & 31);
        if (active && j < cnt) {
          const double2 a = __ldcs(Tc + (int64_t)j * half);
          acc0 = fma(a.x, xc, acc0);
          acc1 = fma(a.y, xc, acc1);
        }
      }
    }
    xv = xnext;
  }
#pragma unroll
  for (int off = 16; off >= LPG; off >>= 1) {       // sum the column groups (fixed order)
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
  }
  if (grp == 0 && active) {
    const int r = 2 * l;
    if (priv) {
      if (ACCUM) {                                   // lists with column chunks of wide tiles: partial products
        atomicAdd(priv + r, acc0);
        if (r + 1 < nrows) atomicAdd(priv + r + 1, acc1);
      } else {
        priv[r] = acc0;
        if (r + 1 < nrows) priv[r + 1] = acc1;
      }
    }
    if (ri) {
      if (ATOMIC) {
        atomicAdd(y + ri[r], acc0);
        if (r + 1 < nrows) atomicAdd(y + ri[r + 1], acc1);
      } else {
        y[ri[r]] += acc0;
        if (r + 1 < nrows) y[ri[r + 1]] += acc1;
      }
    }
  }
}

// Second version of the tile op (default; ALFIB_TILE_V1=1 selects the first).  The ops are small (a few
// KB to 40 KB), so what counts is how soon the matrix stream starts and that it never waits on a
// dependent load.  Each lane keeps a ring of UB column loads in flight: a slot is refilled with the
// column UB*G further on as soon as it has been consumed, so the first UB columns are requested before
// the gathered source has arrived, a partial last batch is just predicated loads, and there is no
// drain between batches.  The source index list is read two 32-column chunks ahead and the source
// values one chunk ahead (no dependent wait at a chunk switch).
// ALFIB_FUSE_INDEX=1: the two index kernels are folded into the source fetch of the op that consumes their output
// (same summation order, so the results are bitwise those of the separate kernels):
//   MODE 1 (K3, X_SS ops): the entry e >= 0 of the separator right-hand side is formed on the fly,
//                          x[idx[e]] - sum_j g[lst[j]], j in [ptr[e], ptr[e+1])          (= sep_rhs_kernel)
//   MODE 2 (K4, shared form): the entry ~e of z is formed on the fly, sum_j g[lst[j]]     (= slot_sum_kernel)
struct FusedSrc {
  const int32_t* idx;
  const int32_t* ptr;
  const int32_t* lst;
  const double* g;
};

template <int G, bool ATOMIC, bool ACCUM, int MODE = 0>
__device__ __forceinline__ void tile_op_body_v2(const double* __restrict__ mat, const int32_t* __restrict__ ci,
                                                const int32_t* __restrict__ ri, double* __restrict__ priv,
                                                const int nrows, const int n, const double* __restrict__ srcA,
                                                const double* __restrict__ srcB, double* __restrict__ y, int lane,
                                                const FusedSrc fs = FusedSrc{nullptr, nullptr, nullptr, nullptr}) {
  constexpr int LPG = 32 / G, UB = 8, CPB = UB * G;   // lanes per column group; ring slots; columns per round
  const int half = (nrows + 1) >> 1;
  const int grp = lane / LPG, l = lane - grp * LPG;
  const bool active = l < half;
  const double2* __restrict__ T = reinterpret_cast<const double2*>(mat) + (active ? l : 0);
  auto column = [&](int c) { return (active && c < n) ? __ldcs(T + c * half) : make_double2(0.0, 0.0); };
  auto index_at = [&](int pos) { return pos < n ? __ldg(ci + pos) : 0; };
  auto value_at = [&](int e, int pos) {
    if (pos >= n) return 0.0;
    if (MODE == 1 && e >= 0) {
      double v = __ldg(srcA + __ldg(fs.idx + e));
      const int j1 = __ldg(fs.ptr + e + 1);
      for (int j = __ldg(fs.ptr + e); j < j1; ++j) v -= fs.g[__ldg(fs.lst + j)];
      return v;
    }
    if (MODE == 2 && e < 0) {
      const int i = ~e;
      double v = 0.0;
      const int j1 = __ldg(fs.ptr + i + 1);
      for (int j = __ldg(fs.ptr + i); j < j1; ++j) v += fs.g[__ldg(fs.lst + j)];
      return v;
    }
    return e >= 0 ? __ldg(srcA + e) : __ldg(srcB + (~e));
  };
  const int e0 = index_at(lane), e1 = index_at(32 + lane);
  double2 a[UB];
#pragma unroll
  for (int u = 0; u < UB; ++u) a[u] = column(u * G + grp);
  double xv = value_at(e0, lane);
  double xnext = value_at(e1, 32 + lane);
  int enext = index_at(64 + lane);
  double acc0 = 0.0, acc1 = 0.0, bcc0 = 0.0, bcc1 = 0.0;   // even / odd slots: two independent FMA chains
  for (int cb = 0; cb < n; cb += CPB) {               // CPB divides 32: a round never straddles a chunk
    if (cb > 0 && (cb & 31) == 0) {
      xv = xnext;
      xnext = value_at(enext, cb + 32 + lane);
      enext = index_at(cb + 64 + lane);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int c = cb + u * G + grp;
      const double xc = __shfl_sync(0xffffffffu, xv, c & 31);
      if (u & 1) {
        bcc0 = fma(a[u].x, xc, bcc0);
        bcc1 = fma(a[u].y, xc, bcc1);
      } else {
        acc0 = fma(a[u].x, xc, acc0);
        acc1 = fma(a[u].y, xc, acc1);
      }
      a[u] = column(c + CPB);                         // predicated off behind the last column
    }
  }
  acc0 += bcc0;
  acc1 += bcc1;
#pragma unroll
  for (int off = 16; off >= LPG; off >>= 1) {       // sum the column groups (fixed order)
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
    acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
  }
  if (grp == 0 && active) {
    const int r = 2 * l;
    if (priv) {
      if (ACCUM) {                                   // lists with column chunks of wide tiles: partial products
        atomicAdd(priv + r, acc0);
        if (r + 1 < nrows) atomicAdd(priv + r + 1, acc1);
      } else {
        priv[r] = acc0;
        if (r + 1 < nrows) priv[r + 1] = acc1;
      }
    }
    if (ri) {
      if (ATOMIC) {
        atomicAdd(y + ri[r], acc0);
        if (r + 1 < nrows) atomicAdd(y + ri[r + 1], acc1);
      } else {
        y[ri[r]] += acc0;
        if (r + 1 < nrows) y[ri[r + 1]] += acc1;
      }
    }
  }
}

template <bool ATOMIC, bool ACCUM = false, int MODE = 0>
__global__ void __launch_bounds__(128) tile_ops_kernel_v2(const TileOp* __restrict__ ops, int nops,
                                                          const int32_t* __restrict__ cidx,
                                                          const double* __restrict__ store,
                                                          const double* __restrict__ srcA,
                                                          const double* __restrict__ srcB, PeerOut yout,
                                                          double* __restrict__ dstB,
                                                          const FusedSrc fs = FusedSrc{nullptr, nullptr, nullptr, nullptr}) {
  double* __restrict__ y = resolve(yout);
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nops) return;
  const TileOp* __restrict__ op = ops + w;
  const double* __restrict__ mat = store + op->mat;
  const int32_t* __restrict__ ci = cidx + op->col;
  const long long row = op->row, pv = op->priv;
  const int32_t* __restrict__ ri = row >= 0 ? cidx + row : nullptr;
  double* __restrict__ priv = pv >= 0 ? dstB + pv : nullptr;
  const int nrows = op->nrows, ncols = op->ncols;
  const int half = (nrows + 1) >> 1;
  if (half <= 8)
    tile_op_body_v2<4, ATOMIC, ACCUM, MODE>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane, fs);
  else if (half <= 16)
    tile_op_body_v2<2, ATOMIC, ACCUM, MODE>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane, fs);
  else
    tile_op_body_v2<1, ATOMIC, ACCUM, MODE>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane, fs);
}

#include "tile_tma.cuh"

template <bool ATOMIC, bool ACCUM, int MODE, int CFG, bool RED>
void launch_tile_ops_tma_cfg(alfib_ctx* c, const TileOp* ops, int nops, const int32_t* cidx, const double* store,
                             const double* srcA, const double* srcB, PeerOut y, double* dstB, const FusedSrc& fs) {
  static bool configured = false;       // per instantiation
  constexpr int WARPS = tma::Cfg<CFG>::WARPS;
  const size_t smem = tma::smem_bytes<CFG>();
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tile_ops_kernel_tma<ATOMIC, ACCUM, MODE, CFG, RED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    configured = true;
  }
  if (!c->tile_counter.p) {
    c->tile_counter.alloc(4);
    CUDA_TRY(cudaMemsetAsync(c->tile_counter.p, 0, 4 * sizeof(unsigned), c->stream));
  }
  const int grid = std::min(c->num_sms, cdiv(nops, WARPS));
  tile_ops_kernel_tma<ATOMIC, ACCUM, MODE, CFG, RED><<<grid, WARPS * 32, smem, c->stream>>>(ops, nops, cidx, store, srcA, srcB, y,
                                                                                            dstB, fs, c->tile_counter.p);
}

// ALFIB_TILE_TMA_CFG (measured on cfg5's finest level, ms per application, profiles/apply_variants_r2.txt; v2 = 0.567):
//   0 = 8 warps x 8 KB stages, read-modify-write scatter (0.501)   1 = 16 warps x 4 KB (0.509)
//   2 = 16 warps x 4 KB + RED scatter (0.506)                      3 = 8 warps x 8 KB + RED scatter (0.496, default)
template <bool ATOMIC, bool ACCUM, int MODE>
void launch_tile_ops_tma(alfib_ctx* c, const TileOp* ops, int nops, const int32_t* cidx, const double* store,
                         const double* srcA, const double* srcB, PeerOut y, double* dstB, const FusedSrc& fs) {
  const char* env_cfg = std::getenv("ALFIB_TILE_TMA_CFG");
  const int cfg = env_cfg ? std::atoi(env_cfg) : 3;
  if (cfg == 2)
    launch_tile_ops_tma_cfg<ATOMIC, ACCUM, MODE, 1, true>(c, ops, nops, cidx, store, srcA, srcB, y, dstB, fs);
  else if (cfg == 1)
    launch_tile_ops_tma_cfg<ATOMIC, ACCUM, MODE, 1, false>(c, ops, nops, cidx, store, srcA, srcB, y, dstB, fs);
  else if (cfg == 3)
    launch_tile_ops_tma_cfg<ATOMIC, ACCUM, MODE, 0, true>(c, ops, nops, cidx, store, srcA, srcB, y, dstB, fs);
  else
    launch_tile_ops_tma_cfg<ATOMIC, ACCUM, MODE, 0, false>(c, ops, nops, cidx, store, srcA, srcB, y, dstB, fs);
}

template <bool ATOMIC, bool ACCUM = false>
__global__ void __launch_bounds__(128) tile_ops_kernel(const TileOp* __restrict__ ops, int nops,
                                                       const int32_t* __restrict__ cidx,
                                                       const double* __restrict__ store,
                                                       const double* __restrict__ srcA,
                                                       const double* __restrict__ srcB, PeerOut yout,
                                                       double* __restrict__ dstB) {
  double* __restrict__ y = resolve(yout);
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nops) return;
  const TileOp* __restrict__ op = ops + w;
  const double* __restrict__ mat = store + op->mat;
  const int32_t* __restrict__ ci = cidx + op->col;
  const long long row = op->row, pv = op->priv;
  const int32_t* __restrict__ ri = row >= 0 ? cidx + row : nullptr;
  double* __restrict__ priv = pv >= 0 ? dstB + pv : nullptr;
  const int nrows = op->nrows, ncols = op->ncols;
  const int half = (nrows + 1) >> 1;
  if (half <= 8)
    tile_op_body<4, ATOMIC, ACCUM>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane);
  else if (half <= 16)
    tile_op_body<2, ATOMIC, ACCUM>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane);
  else
    tile_op_body<1, ATOMIC, ACCUM>(mat, ci, ri, priv, nrows, ncols, srcA, srcB, y, lane);
}

// K2: separator right-hand sides, rs[e] = x[sepdofs[e]] - sum_j g1[cg1[j]], j in [cptr[e], cptr[e+1])
__global__ void sep_rhs_kernel(int64_t nsep, const int32_t* __restrict__ sepdofs, const int32_t* __restrict__ cptr,
                               const int32_t* __restrict__ cg1, const double* __restrict__ x,
                               const double* __restrict__ g1, double* __restrict__ rs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsep) return;
  double v = __ldg(x + sepdofs[e]);
  // four index -> value pairs in flight per thread (the plain loop is a chain of dependent loads: long_scoreboard was
  // its only stall); subtracted in list order, and v - 0.0 == v, so the bits are those of the plain loop
  const int j1 = cptr[e + 1];
  for (int j = cptr[e]; j < j1; j += 4) {
    double g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = (j + u < j1) ? g1[cg1[j + u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) v -= g[u];
  }
  rs[e] = v;
}

// K3b (shared form): z[e] = sum_j us[zsrc[j]], j in [zptr[e], zptr[e+1]) — the separator solutions of the
// visited patches gathered to the positions of each distinct block's neighbourhood U_k, fixed order
__global__ void slot_sum_kernel(int64_t nslots, const int32_t* __restrict__ zptr, const int32_t* __restrict__ zsrc,
                                const double* __restrict__ us, double* __restrict__ z) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nslots) return;
  double v = 0.0;
  const int j1 = zptr[e + 1];
  for (int j = zptr[e]; j < j1; j += 4) {          // as in sep_rhs_kernel: independent loads, list order
    double g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = (j + u < j1) ? us[zsrc[j + u]] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) v += g[u];
  }
  z[e] = v;
}

// ------------------------------------------------------------------------------------------------
// Per-Newton-step block setup: one CTA per (patch, block).  Gathers A_kk, A_kN, A_Nk from the BSR
// values, inverts A_kk in shared memory by Gauss-Jordan with partial (row) pivoting (first-max
// rule, like the patch kernel and LAPACK idamax) and writes the V and [D | -W] tiles.
constexpr int CT = 128;

struct CondenseArgs {
  int64_t nblocks;
  const BlockDesc* blocks;
  const int32_t* bdofs;
  const int32_t* bkeys;
  const int32_t* bperm;
  int bs;
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  double* store;
  int maxb, maxm;
  int* info;
  // Schur-complement setup: A_kN is carried through the elimination (Ws = A_kk^-1 A_kN with solve accuracy), the
  // W tile is -Ws and C = A_Nk Ws goes to cbuf + coff[q] (m x m, column-major); null: the legacy products
  double* cbuf;
  const int64_t* coff;
};

__device__ __forceinline__ int bsearch_i32(const int32_t* a, int n, int key) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = a[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

__global__ void __launch_bounds__(CT) condense_blocks_kernel(CondenseArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ldk_max = a.maxb + 1;
  double* Akk = reinterpret_cast<double*>(smem_raw);          // b x b, column-major, ld = b + 1
  double* AkN = Akk + (size_t)a.maxb * ldk_max;               // b x m, column-major, ld = b
  double* ANk = AkN + (size_t)a.maxb * a.maxm;                // m x b, column-major, ld = m
  double* prow = ANk + (size_t)a.maxb * a.maxm;               // b
  double* fcol = prow + a.maxb;                               // b
  int* piv = reinterpret_cast<int*>(fcol + a.maxb);           // b
  int* src = piv + a.maxb;                                    // b
  int* keys = src + a.maxb;                                   // b + m
  int* perm = keys + a.maxb + a.maxm;                         // b + m
  int* s_flag = perm + a.maxb + a.maxm;                       // [0] pivot row, [1] singular
  double* prowN = reinterpret_cast<double*>(s_flag + 4);      // m (Schur setup: pivot row of the carried A_kN)
  const bool schur = a.cbuf != nullptr;

  for (int64_t q = blockIdx.x; q < a.nblocks; q += gridDim.x) {
    const BlockDesc d = a.blocks[q];
    const int b = d.b, m = d.m, ld = b + 1, bm = b + m;
    const int32_t* dofs = a.bdofs + d.dofs;
    __syncthreads();                                         // previous block's readers are done
    for (int i = tid; i < b * ld; i += CT) Akk[i] = 0.0;
    for (int i = tid; i < b * m; i += CT) { AkN[i] = 0.0; ANk[i] = 0.0; }
    for (int i = tid; i < bm; i += CT) { keys[i] = a.bkeys[d.keys + i]; perm[i] = a.bperm[d.keys + i]; }
    __syncthreads();
    // ---- gather: one warp per row of [B; N], lanes over the BSR row's blocks ------------------
    {
      const int bs = a.bs, b2 = bs * bs;
      for (int rp = warp; rp < bm; rp += CT / 32) {
        const int g = dofs[rp];
        const int node = g / bs, comp = g - node * bs;
        const int k1 = a.rowptr[node + 1];
        for (int k = a.rowptr[node] + lane; k < k1; k += 32) {
          const int cn = a.colidx[k];
          for (int c2 = 0; c2 < bs; ++c2) {
            const int hit = bsearch_i32(keys, bm, cn * bs + c2);
            if (hit < 0) continue;
            const int cp = perm[hit];
            const double v = a.vals[(int64_t)k * b2 + comp * bs + c2];
            if (rp < b) {
              if (cp < b) Akk[rp + cp * ld] = v; else AkN[rp + (cp - b) * b] = v;
            } else if (cp < b) {
              ANk[(rp - b) + cp * m] = v;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- in-place Gauss-Jordan inversion of Akk with partial pivoting --------------------------
    for (int j = 0; j < b; ++j) {
      if (warp == 0) {
        double best = -1.0;
        int bi = INT_MAX;
        for (int r = j + lane; r < b; r += 32) {
          const double v = fabs(Akk[r + j * ld]);
          if (v > best) { best = v; bi = r; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
          const double ov = __shfl_down_sync(0xffffffffu, best, off);
          const int oi = __shfl_down_sync(0xffffffffu, bi, off);
          if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) {
          const bool singular = !(best > 0.0);
          if (singular) { bi = j; atomicCAS(a.info, 0, (int)(q % INT_MAX) + 1); }
          piv[j] = bi;
          s_flag[0] = bi;
          s_flag[1] = singular ? 1 : 0;
        }
      }
      __syncthreads();
      const int pv = s_flag[0];
      const bool singular = s_flag[1] != 0;
      if (pv != j && tid < b) {
        const double t0 = Akk[j + tid * ld];
        Akk[j + tid * ld] = Akk[pv + tid * ld];
        Akk[pv + tid * ld] = t0;
      }
      if (schur && pv != j && tid < m) {                     // the same row swap on the carried right-hand sides
        const double t0 = AkN[j + tid * b];
        AkN[j + tid * b] = AkN[pv + tid * b];
        AkN[pv + tid * b] = t0;
      }
      __syncthreads();
      const double dinv = singular ? 0.0 : 1.0 / Akk[j + j * ld];
      if (tid < b) {
        prow[tid] = (tid == j) ? 0.0 : Akk[j + tid * ld] * dinv;
        fcol[tid] = Akk[tid + j * ld];
      }
      if (schur && tid < m) prowN[tid] = AkN[j + tid * b] * dinv;
      __syncthreads();
      if (schur)
        for (int i = tid; i < b * m; i += CT) {
          const int cc = i / b, r = i - cc * b;
          AkN[i] = (r == j) ? prowN[cc] : fma(-fcol[r], prowN[cc], AkN[i]);
        }
      for (int i = tid; i < b * b; i += CT) {
        const int cc = i / b, r = i - cc * b;
        double v;
        if (r == j)
          v = (cc == j) ? dinv : prow[cc];
        else if (cc == j)
          v = -fcol[r] * dinv;
        else
          v = fma(-fcol[r], prow[cc], Akk[r + cc * ld]);
        Akk[r + cc * ld] = v;
      }
      __syncthreads();
    }
    // undo the row pivoting: A^-1 = M with the column swaps applied in reverse order
    if (tid == 0) {
      for (int cc = 0; cc < b; ++cc) src[cc] = cc;
      for (int k = b - 1; k >= 0; --k) {
        const int pv = piv[k];
        if (pv != k) { const int t0 = src[k]; src[k] = src[pv]; src[pv] = t0; }
      }
    }
    __syncthreads();
    // D[r][c] = Akk[r + src[c] * ld]
    // ---- V = A_Nk D  (m x b), tile with roundup2(m) rows per column ------------------------------
    {
      const int mr = (m + 1) & ~1;
      double* Vt = a.store + d.voff;
      for (int i = tid; i < mr * b; i += CT) {
        const int cc = i / mr, r = i - cc * mr;
        double v = 0.0;
        if (r < m) {
          const double* Dc = Akk + (size_t)src[cc] * ld;
          for (int k = 0; k < b; ++k) v = fma(ANk[r + k * m], Dc[k], v);
        }
        Vt[i] = v;
      }
    }
    // ---- [D | -W], W = D A_kN  (b x m), tile with roundup2(b) rows per column ---------------------
    {
      const int br = (b + 1) & ~1;
      double* Dt = a.store + d.dwoff;
      for (int i = tid; i < br * b; i += CT) {
        const int cc = i / br, r = i - cc * br;
        Dt[i] = (r < b) ? Akk[r + (size_t)src[cc] * ld] * d.dscale : 0.0;
      }
      double* Wt = Dt + (size_t)br * b;
      for (int i = tid; i < br * m; i += CT) {
        const int cc = i / br, r = i - cc * br;
        double v = 0.0;
        if (r < b) {
          if (schur) {
            v = AkN[r + (size_t)cc * b];                     // Ws: already A_kk^-1 A_kN
          } else {
            const double* Ac = AkN + (size_t)cc * b;
            for (int k = 0; k < b; ++k) v = fma(Akk[r + (size_t)src[k] * ld], Ac[k], v);
          }
        }
        Wt[i] = -v;
      }
    }
    // ---- Schur setup: C = A_Nk Ws  (m x m) ------------------------------------------------------------
    if (schur) {
      double* C = a.cbuf + a.coff[q];
      for (int i = tid; i < m * m; i += CT) {
        const int cc = i / m, r = i - cc * m;
        const double* Wc = AkN + (size_t)cc * b;
        double v = 0.0;
        for (int k = 0; k < b; ++k) v = fma(ANk[r + k * m], Wc[k], v);
        C[i] = v;
      }
    }
  }
}

size_t condense_smem_bytes(int maxb, int maxm) {
  const size_t doubles = (size_t)maxb * (maxb + 1) + 2 * (size_t)maxb * maxm + 2 * (size_t)maxb + (size_t)maxm;
  const size_t ints = 2 * (size_t)maxb + 2 * (size_t)(maxb + maxm) + 4;
  return doubles * sizeof(double) + ints * sizeof(int) + 16;
}

}  // namespace

// ---- host: block structure -> op lists (condense_host.h), uploaded here ----------------------------
void condense_setup(alfib_ctx* c, Level& L, PatchSet& ps, const int32_t* block_of_dof) {
  ALFIB_REQUIRE(!L.h_rowptr.empty(), "set the BSR pattern before the patch blocks");
  Condensed& cd = ps.cond;
  cd.release();
  cd.on = false;
  PatchView pv;
  pv.npatch = ps.npatch;
  pv.ncolour = ps.ncolour;
  pv.bs = L.bs;
  pv.ndofs = L.n;
  pv.off = ps.h_off.data();
  pv.dofs = ps.h_dofs.data();
  pv.order = &ps.h_order;
  pv.colour = ps.h_colour.data();
  pv.rowptr = L.h_rowptr.data();
  pv.colidx = L.h_colidx.data();
  // ALFIB_CONDENSE_SHARED=0 keeps one V / [D | -W] pair per (patch, block) even where blocks could be shared
  const char* env_shared = std::getenv("ALFIB_CONDENSE_SHARED");
  const bool allow_shared = !(env_shared && env_shared[0] == '0');
  try {
    // ALFIB_SPLIT_COLS=n: X_SS tiles wider than 2n columns are cut into n-column chunks (default 256: only coarse
    // levels and literal 3-D macro stars; e.g. 64 also cuts the 195-column tiles of the open macro star in three)
    const char* env_split = std::getenv("ALFIB_SPLIT_COLS");
    const int split_cols = env_split ? std::atoi(env_split) : ALFIB_SPLIT_COLS;
    build_condensed_host(pv, block_of_dof, cd.h, allow_shared, /*split_wide=*/!c->deterministic, split_cols);
  } catch (const std::runtime_error& e) {
    cd.h = CondensedHost();
    throw DeviceError{ALFIB_EINVAL, e.what()};
  }
  const CondensedHost& h = cd.h;
  ps.store_elems = h.store_elems;
  ps.store = nullptr;                 // storage requirement changed: (re)allocated on the next factor
  ps.store_buf.release();
  ps.store_owned = false;
  ps.factored = false;
  cd.sepoff.upload(h.sepoff.data(), h.sepoff.size(), c->stream);
  cd.ssoff.upload(h.ssoff.data(), h.ssoff.size(), c->stream);
  cd.seplocal.upload(h.seplocal.data(), h.seplocal.size(), c->stream);
  cd.sepdofs.upload(h.sepdofs.data(), h.sepdofs.size(), c->stream);
  cd.cidx.upload(h.cidx.data(), h.cidx.size(), c->stream);
  if (h.shared) {                       // block setup over the distinct blocks
    cd.bdofs.upload(h.sdofs.data(), h.sdofs.size(), c->stream);
    cd.bkeys.upload(h.skeys.data(), h.skeys.size(), c->stream);
    cd.bperm.upload(h.sperm.data(), h.sperm.size(), c->stream);
    cd.blocks.upload(h.sblocks.data(), h.sblocks.size(), c->stream);
    cd.zptr.upload(h.zptr.data(), h.zptr.size(), c->stream);
    cd.zsrc.upload(h.zsrc.data(), h.zsrc.size(), c->stream);
    cd.z.alloc((size_t)std::max<int64_t>(h.g1_total, 1));
  } else {
    cd.bdofs.upload(h.bdofs.data(), h.bdofs.size(), c->stream);
    cd.bkeys.upload(h.bkeys.data(), h.bkeys.size(), c->stream);
    cd.bperm.upload(h.bperm.data(), h.bperm.size(), c->stream);
    cd.blocks.upload(h.blocks.data(), h.blocks.size(), c->stream);
  }
  cd.cptr.upload(h.cptr.data(), h.cptr.size(), c->stream);
  cd.cg1.upload(h.cg1.data(), h.cg1.size(), c->stream);
  cd.opsV.upload(h.opsV.data(), h.opsV.size(), c->stream);
  cd.opsS.upload(h.opsS.data(), h.opsS.size(), c->stream);
  cd.opsDW.upload(h.opsDW.data(), h.opsDW.size(), c->stream);
  cd.g1.alloc((size_t)std::max<int64_t>(h.g1_total, 1));
  cd.rs.alloc((size_t)std::max<int64_t>(h.nsep_total, 1));
  cd.us.alloc((size_t)std::max<int64_t>(h.nsep_total, 1));
  // slots of blocks / patches outside the iteration set are read (K2, K3b) but never written
  CUDA_TRY(cudaMemsetAsync(cd.g1.p, 0, cd.g1.n * sizeof(double), c->stream));
  CUDA_TRY(cudaMemsetAsync(cd.us.p, 0, cd.us.n * sizeof(double), c->stream));
  if (cd.z.p) CUDA_TRY(cudaMemsetAsync(cd.z.p, 0, cd.z.n * sizeof(double), c->stream));
  // X_SS from the Schur complement formed with solves (~200x fewer flops per Newton step on 3-D macro stars; measured
  // on B200: 3.27 s -> 0.28 s per Newton step on cfg5).  Default; ALFIB_SCHUR_SETUP=0 cuts X_SS out of the pivoted
  // inverse of the whole patch instead (round 1's setup).
  const char* env_schur = std::getenv("ALFIB_SCHUR_SETUP");
  cd.schur = !(env_schur && env_schur[0] == '0');
  if (cd.schur) {
    build_schur_lists(h, ps.npatch, cd.sh);
    const SchurHost& sh = cd.sh;
    cd.sepsorted.upload(sh.sepsorted.data(), sh.sepsorted.size(), c->stream);
    cd.sepperm.upload(sh.sepperm.data(), sh.sepperm.size(), c->stream);
    cd.blk_start.upload(h.blk_start.data(), h.blk_start.size(), c->stream);
    cd.nb_off.upload(h.nb_off.data(), h.nb_off.size(), c->stream);
    cd.nb_pos.upload(h.nb_pos.data(), h.nb_pos.size(), c->stream);
    cd.sc_upos.upload(sh.upos.data(), sh.upos.size(), c->stream);
    cd.sc_inst_ld.upload(sh.inst_ld.data(), sh.inst_ld.size(), c->stream);
    cd.sc_inst_c.upload(sh.inst_c.data(), sh.inst_c.size(), c->stream);
    cd.sc_coff.upload(sh.coff.data(), sh.coff.size(), c->stream);
    cd.cbuf.alloc((size_t)std::max<int64_t>(sh.ctotal, 1));
    // separators, largest first
    std::vector<int32_t> order(ps.npatch);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      return h.sepoff[x + 1] - h.sepoff[x] > h.sepoff[y + 1] - h.sepoff[y];
    });
    cd.sforder.upload(order.data(), order.size(), c->stream);
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cd.on = true;
}

// D, V, W of every block from the current operator values (X_SS is written by patch_factor.cu)
void launch_condense_blocks(alfib_ctx* c, const Level& L, PatchSet& ps, const double* vals, bool schur) {
  Condensed& cd = ps.cond;
  if (cd.h.nblocks == 0) return;
  CondenseArgs a;
  a.nblocks = cd.h.shared ? cd.h.ndist : cd.h.nblocks;
  a.blocks = cd.blocks.p;
  a.bdofs = cd.bdofs.p;
  a.bkeys = cd.bkeys.p;
  a.bperm = cd.bperm.p;
  a.bs = L.bs;
  a.rowptr = L.rowptr.p;
  a.colidx = L.colidx.p;
  a.vals = vals;
  a.store = ps.store;
  a.maxb = cd.h.maxb;
  a.maxm = std::max(cd.h.maxm, 1);
  a.cbuf = schur ? cd.cbuf.p : nullptr;
  a.coff = schur ? cd.sc_coff.p : nullptr;
  c->finfo.alloc(2);
  CUDA_TRY(cudaMemsetAsync(c->finfo.p, 0, 2 * sizeof(int), c->stream));
  a.info = c->finfo.p;
  const size_t smem = condense_smem_bytes(a.maxb, a.maxm);
  CUDA_TRY(cudaFuncSetAttribute(condense_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(a.nblocks, (int64_t)c->num_sms * 16);
  condense_blocks_kernel<<<grid, CT, smem, c->stream>>>(a);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
  int info = 0;
  CUDA_TRY(cudaMemcpyAsync(&info, c->finfo.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (info != 0)
    throw DeviceError{ALFIB_ESINGULAR, "interior block " + std::to_string(info - 1) + " of a condensed patch is singular"};
}

void launch_condensed_apply(alfib_ctx* c, const PatchSet& ps, const double* x, PeerOut y) {
  const Condensed& cd = ps.cond;
  const CondensedHost& h = cd.h;
  const int threads = 128, wpb = threads / 32;
  const int nV = (int)h.opsV.size(), nS = (int)h.opsS.size(), nDW = (int)h.opsDW.size();
  const char* env_v1 = std::getenv("ALFIB_TILE_V1");
  const bool v1 = env_v1 && env_v1[0] == '1';
  // K1: V ops, all patches, plain private stores
  const char* env_tma = std::getenv("ALFIB_TILE_TMA");
  const bool tma_on = !v1 && !(env_tma && env_tma[0] == '0');      // default; ALFIB_TILE_TMA=0: the LDG stream of v2
  const char* env_tma_min = std::getenv("ALFIB_TILE_TMA_MIN_OPS");
  const int tma_min_ops = env_tma_min ? std::atoi(env_tma_min) : 1;
  if (nV && tma_on && nV >= tma_min_ops) {
    launch_tile_ops_tma<false, false, 0>(c, cd.opsV.p, nV, cd.cidx.p, ps.store, x, nullptr, plain_out(nullptr), cd.g1.p,
                                         FusedSrc{nullptr, nullptr, nullptr, nullptr});
    c->launches++;
  } else if (nV) {
    if (v1)
      tile_ops_kernel<false><<<cdiv(nV, wpb), threads, 0, c->stream>>>(cd.opsV.p, nV, cd.cidx.p, ps.store, x, nullptr,
                                                                        plain_out(nullptr), cd.g1.p);
    else
      tile_ops_kernel_v2<false><<<cdiv(nV, wpb), threads, 0, c->stream>>>(cd.opsV.p, nV, cd.cidx.p, ps.store, x, nullptr,
                                                                           plain_out(nullptr), cd.g1.p);
    c->launches++;
  }
  // ALFIB_FUSE_INDEX=1 (tile op v2 only): K2 and K3b are folded into the source fetch of K3 / K4.  Off by default:
  // measured on B200 (profiles/bench_r2_*): cfg5's finest level 0.583 ms fused against 0.566 ms, and with the fused form
  // on the SMALL sets only (level 1, coarse patch) a cycle is 1.9 ms slower (PCPATCHApply 16.1 against 14.2 ms, coarse
  // solve 0.63 against 0.54 ms) — under graph replay the two launches cost less than the stalled matrix stream.
  const char* env_fuse = std::getenv("ALFIB_FUSE_INDEX");
  const bool fuse = !v1 && env_fuse && env_fuse[0] == '1';
  const FusedSrc fs_rhs{cd.sepdofs.p, cd.cptr.p, cd.cg1.p, cd.g1.p};
  const FusedSrc fs_z{nullptr, cd.zptr.p, cd.zsrc.p, cd.us.p};
  // K2: separator right-hand sides
  if (h.nsep_total && !fuse) {
    sep_rhs_kernel<<<cdiv(h.nsep_total, 256), 256, 0, c->stream>>>(h.nsep_total, cd.sepdofs.p, cd.cptr.p, cd.cg1.p, x,
                                                                    cd.g1.p, cd.rs.p);
    c->launches++;
  }
  // K3 (separator solve + scatter), K4 (block back-substitution + scatter)
  if (h.any_accum) CUDA_TRY(cudaMemsetAsync(cd.us.p, 0, cd.us.n * sizeof(double), c->stream));
  const bool coloured = c->deterministic && !ps.repeated;
  const bool s_atomic = ps.repeated || h.any_accum;      // column chunks of one tile add to the same rows
  auto run = [&](const TileOp* ops, int nops, const double* srcA, const double* srcB, double* dstB, bool atomic,
                 bool accum = false, int mode = 0) {
    if (nops <= 0) return;
    if (tma_on && nops >= tma_min_ops) {   // v3: TMA bulk copies into per-warp shared-memory rings (tile_tma.cuh)
      const FusedSrc none{nullptr, nullptr, nullptr, nullptr};
      const int32_t* ci = cd.cidx.p;
      if (mode == 1) {
        if (accum) launch_tile_ops_tma<true, true, 1>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, fs_rhs);
        else if (atomic) launch_tile_ops_tma<true, false, 1>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, fs_rhs);
        else launch_tile_ops_tma<false, false, 1>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, fs_rhs);
      } else if (mode == 2) {
        launch_tile_ops_tma<false, false, 2>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, fs_z);
      } else if (accum) {
        launch_tile_ops_tma<true, true, 0>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, none);
      } else if (atomic) {
        launch_tile_ops_tma<true, false, 0>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, none);
      } else {
        launch_tile_ops_tma<false, false, 0>(c, ops, nops, ci, ps.store, srcA, srcB, y, dstB, none);
      }
      c->launches++;
      return;
    }
    if (mode == 1) {                 // fused separator right-hand side: srcA is x
      const int g = cdiv(nops, wpb);
      if (accum)
        tile_ops_kernel_v2<true, true, 1><<<g, threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB, fs_rhs);
      else if (atomic)
        tile_ops_kernel_v2<true, false, 1><<<g, threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB, fs_rhs);
      else
        tile_ops_kernel_v2<false, false, 1><<<g, threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB, fs_rhs);
      c->launches++;
      return;
    }
    if (mode == 2) {                 // fused slot sums (shared form, K4: plain stores)
      tile_ops_kernel_v2<false, false, 2><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y,
                                                                                      dstB, fs_z);
      c->launches++;
      return;
    }
    if (accum) {                     // X_SS lists with column chunks: every private write is an atomicAdd into zeroed us
      if (v1)
        tile_ops_kernel<true, true><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
      else
        tile_ops_kernel_v2<true, true><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
      c->launches++;
      return;
    }
    if (v1) {
      if (atomic)
        tile_ops_kernel<true><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
      else
        tile_ops_kernel<false><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
    } else {
      if (atomic)
        tile_ops_kernel_v2<true><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
      else
        tile_ops_kernel_v2<false><<<cdiv(nops, wpb), threads, 0, c->stream>>>(ops, nops, cd.cidx.p, ps.store, srcA, srcB, y, dstB);
    }
    c->launches++;
  };
  if (h.shared) {
    // shared blocks: K3 as below; K3b sums the separator solutions per distinct block; K4 is one launch
    // over the distinct blocks, which are pairwise disjoint (plain stores, deterministic in every mode)
    if (!coloured && ps.ncolour > 1) {
      run(cd.opsS.p, nS, fuse ? x : cd.rs.p, nullptr, cd.us.p, true, h.any_accum, fuse ? 1 : 0);
    } else {
      for (int col = 0; col < ps.ncolour; ++col) {
        const int s = h.s_colour_start[col], e = h.s_colour_start[col + 1];
        run(cd.opsS.p + s, e - s, fuse ? x : cd.rs.p, nullptr, cd.us.p, s_atomic, h.any_accum, fuse ? 1 : 0);
      }
    }
    if (h.g1_total && !fuse) {
      slot_sum_kernel<<<cdiv(h.g1_total, 256), 256, 0, c->stream>>>(h.g1_total, cd.zptr.p, cd.zsrc.p, cd.us.p, cd.z.p);
      c->launches++;
    }
    run(cd.opsDW.p, nDW, x, cd.z.p, nullptr, false, false, fuse ? 2 : 0);
  } else if (!coloured && ps.ncolour > 1) {
    run(cd.opsS.p, nS, fuse ? x : cd.rs.p, nullptr, cd.us.p, true, h.any_accum, fuse ? 1 : 0);
    run(cd.opsDW.p, nDW, x, cd.us.p, nullptr, true);
  } else {
    for (int col = 0; col < ps.ncolour; ++col) {
      const int s = h.s_colour_start[col], e = h.s_colour_start[col + 1];
      run(cd.opsS.p + s, e - s, fuse ? x : cd.rs.p, nullptr, cd.us.p, s_atomic, h.any_accum, fuse ? 1 : 0);
    }
    for (int col = 0; col < ps.ncolour; ++col) {
      const int s = h.dw_colour_start[col], e = h.dw_colour_start[col + 1];
      run(cd.opsDW.p + s, e - s, x, cd.us.p, nullptr, ps.repeated);
    }
  }
  CUDA_TRY(cudaGetLastError());
}

// tests / debugging: dense inverse of one patch rebuilt on the host from its condensed factors
void condensed_extract_inverse(alfib_ctx* c, const PatchSet& ps, int patch, double* host_out) {
  ALFIB_REQUIRE(patch >= 0 && patch < ps.npatch, "patch index out of range");
  ALFIB_REQUIRE(ps.factored, "patches not factored");
  const int n = (int)(ps.h_off[patch + 1] - ps.h_off[patch]);
  auto download = [&](int64_t off, int64_t count) {
    std::vector<double> t((size_t)std::max<int64_t>(count, 1));
    if (count) CUDA_TRY(cudaMemcpyAsync(t.data(), ps.store + off, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return t;
  };
  condensed_inverse_host(ps.cond.h, patch, n, download, host_out);
}
