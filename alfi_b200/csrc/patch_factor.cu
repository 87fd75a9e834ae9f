// Per-Newton-step patch setup (PCSetUp_PATCH): gather A_i = A[I_i, I_i] from the level's BSR
// values and invert it — `patch_pc_patch_save_operators`, `patch_sub_pc_type lu` and
// `patch_pc_patch_dense_inverse` of alfi/solver.py:320,327,599-602 (the reference does this
// with LAPACK getrf+getri through PETSc).
//
// One persistent CTA per SM takes patches from an atomic counter (largest first).  The patch
// matrix lives column-major in a per-CTA global workspace slot (13 MB for n = 1275) and is
// inverted in place by *two-level blocked Gauss-Jordan with partial row pivoting*:
//
//   for each outer block KO of NBO = 64 columns
//     for each inner panel K of NB (16) columns of KO                      [O(n NBO^2) work]
//       1. panel -> shared memory; NB unblocked Gauss-Jordan steps on all n rows of the panel,
//          pivot = first max |.| among the not-yet-pivoted rows (LAPACK idamax rule);
//       2. the panel's row swaps, R = W[K, .] and the rank-NB update are applied to the other
//          columns *of the outer block only* (one row per thread, N[r, 0:NB] in registers);
//     far update, all columns outside KO                                   [the O(n^3) part]
//       3. all NBO row swaps in order, R = W[KO, far] set aside (k-major), W[KO, far] zeroed;
//       4. W[:, far] += N * R with N = W[:, KO]: a rank-64 FP64 GEMM on the tensor cores
//          (mma.sync.m8n8k4.f64, DMMA): 256 x 64 tiles of W per CTA iteration, a 32 x 32 sub-tile
//          per warp, N tile (135 KB) and double-buffered cp.async R tiles (2 x 37 KB) in shared
//          memory, W read and written once per outer block with 128-bit accesses.
//   A^{-1} = W with the column swaps undone in reverse order; that permutation is folded into
//   the final pass that writes the 64-row-tiled apply layout (patch_apply.cu).
//
// The far columns see the composition of the NBO/NB inner transformations as one rank-NBO
// update  C <- Z_KO C + N_all C[KO, :]  because in-place Gauss-Jordan keeps, in the columns it has
// eliminated, the image of the corresponding unit vectors (validated in numpy first, see
// DESIGN.md §3.2).  Flops 2 n^3 per patch; arithmetic intensity of the far update 8 flop/byte, so
// it is bound by the FP64 pipe, not by HBM (the one-level version streamed W once per 16 columns
// and sat at ~1.5 TB/s aggregate).
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cuda_pipeline.h>
#include <numeric>

#include "alfib_internal.h"

namespace {

constexpr int FT = 512;        // threads per CTA
constexpr int NBO = 64;        // outer block (rank of the far update)
constexpr int TR = 256;        // rows of W per far tile
constexpr int TC = 64;         // columns of W per far tile
constexpr int NS_LD = TR + 8;  // padded shared-memory rows: fragment loads (4 k x 8 indices) hit every bank pair twice
constexpr int RS_LD = TC + 8;
constexpr int UC = 8;          // columns whose loads are issued together in the in-block update

struct FactorArgs {
  int npatch;
  const int32_t* forder;       // patches, largest first
  const int64_t* poff;
  const int32_t* pdofs;
  const int32_t* sorted;       // patch dofs sorted ascending ...
  const int32_t* sperm;        // ... and their local indices
  const int64_t* soff;
  double* store;
  // condensed sets (condense.cu): keep only X[S, S], S = separator dofs (patch-local indices); else null
  const int64_t* sepoff;
  const int32_t* seplocal;
  // level operator
  int bs;
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  // workspace
  double* work;
  int64_t slot_elems;          // per CTA: W (maxn*ld) + R (NBO*maxn)
  int maxn;
  int* counter;
  int* info;                   // 0 or (1 + index of a singular patch)
  long long* timing;           // optional per-phase clock64 totals (ALFIB_FACTOR_TIMING=1), else null
  // Schur-complement setup (condense.cu, condense_host.h build_schur_lists): the "patches" are the separators,
  // and after the gather of A_SS every block instance q of patch p subtracts its C = A_Nk A_kk^-1 A_kN:
  //   W[nb_pos[i], nb_pos[j]] -= C_q[upos[i], upos[j]],  i, j in [nb_off[q], nb_off[q+1]);   null: plain patches
  const int64_t* sc_blk_start;
  const int64_t* sc_nb_off;
  const int32_t* sc_nb_pos;
  const int32_t* sc_upos;
  const int64_t* sc_inst_c;
  const int32_t* sc_inst_ld;
  const double* sc_cbuf;
  // patch operators that are not sub-matrices (alfib_level_set_patch_corrections): after the gather
  //   W[rows[e], cols[e]] += vals[e],  e in [corr_off[p], corr_off[p+1]), entries distinct;   null: none
  const int64_t* corr_off;
  const int32_t* corr_rows;
  const int32_t* corr_cols;
  const double* corr_vals;
};

enum { PH_GATHER, PH_PANEL_LOAD, PH_INNER_GJ, PH_INBLOCK, PH_FAR_SWAP, PH_FAR_NS, PH_FAR_RS, PH_FAR_MMA, PH_PACK, PH_COUNT };

// thread 0 of a CTA accumulates the cycles between phase boundaries (all boundaries follow a barrier)
struct PhaseClock {
  long long t0, acc[PH_COUNT];
  bool on;
  __device__ void start(bool enable) {
    on = enable;
    if (on) {
      for (int i = 0; i < PH_COUNT; ++i) acc[i] = 0;
      t0 = clock64();
    }
  }
  __device__ void mark(int phase) {
    if (on) {
      const long long t = clock64();
      acc[phase] += t - t0;
      t0 = t;
    }
  }
  __device__ void flush(long long* out) {
    if (on)
      for (int i = 0; i < PH_COUNT; ++i) atomicAdd((unsigned long long*)out + i, (unsigned long long)acc[i]);
  }
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

__host__ __device__ inline size_t front_doubles(int maxn, int NB) {
  // the front region holds either the inner panel (+ its R staging) or the far tiles
  const size_t ldp = (size_t)((maxn + 1) & ~1);
  const size_t inner = ldp * NB + (size_t)NBO * NB;
  const size_t far = (size_t)NBO * NS_LD + 2 * (size_t)NBO * RS_LD;   // N tile + double-buffered R tiles
  return inner > far ? inner : far;
}

template <int NB>
__global__ void __launch_bounds__(FT, 1) patch_factor_kernel(FactorArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int ldp_max = (a.maxn + 1) & ~1;
  // shared memory carve-up
  double* front = reinterpret_cast<double*>(smem_raw);
  double* panel = front;                                               // ldp_max * NB
  double* Rs = panel + (size_t)ldp_max * NB;                           // NBO * NB (in-block R)
  double* Ns = front;                                                  // NBO * TR   (far phase)
  double* Rsm = Ns + (size_t)NBO * NS_LD;                              // 2 x NBO * RS_LD (far phase)
  double* pr = front + front_doubles(a.maxn, NB);                      // NB
  double* red_v = pr + NB;                                             // 34
  int* red_i = reinterpret_cast<int*>(red_v + 34);                     // 34
  int* swapA = red_i + 34;                                             // NBO
  int* swapB = swapA + NBO;                                            // NBO
  int* s_misc = swapB + NBO;                                           // [0] next patch, [1] nswap
  int* ipiv = s_misc + 2;                                              // maxn
  // the gather's lookup tables alias the front region (used before the factorisation starts)
  int* sd = reinterpret_cast<int*>(front);
  int* sp = sd + a.maxn;

  double* W = a.work + (size_t)blockIdx.x * a.slot_elems;
  double* Rg = W + (size_t)a.maxn * ldp_max;                           // NBO x ldr, k-major
  const int ldr = (a.maxn + TC - 1) & ~(TC - 1);

  for (;;) {
    if (tid == 0) s_misc[0] = atomicAdd(a.counter, 1);
    __syncthreads();
    const int q = s_misc[0];
    __syncthreads();
    if (q >= a.npatch) break;
    const int p = a.forder[q];
    const int64_t o = a.poff[p];
    const int n = (int)(a.poff[p + 1] - o);
    if (n == 0) continue;
    const int ld = (n + 1) & ~1;
    const int ldp = ld;
    const int32_t* I = a.pdofs + o;
    PhaseClock pc;
    pc.start(a.timing != nullptr && tid == 0);

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
    if (a.sc_blk_start) {
      for (int64_t qi = a.sc_blk_start[p]; qi < a.sc_blk_start[p + 1]; ++qi) {
        const int64_t o2 = a.sc_nb_off[qi];
        const int mq = (int)(a.sc_nb_off[qi + 1] - o2);
        const double* __restrict__ C = a.sc_cbuf + a.sc_inst_c[qi];
        const int ldc = a.sc_inst_ld[qi];
        for (int i = tid; i < mq * mq; i += FT) {
          const int jj = i / mq, ii = i - jj * mq;
          W[a.sc_nb_pos[o2 + ii] + (size_t)a.sc_nb_pos[o2 + jj] * ld] -=
              C[a.sc_upos[o2 + ii] + (size_t)a.sc_upos[o2 + jj] * ldc];
        }
        __syncthreads();                 // the instances of a patch overlap on the separator
      }
    }
    if (a.corr_off) {
      for (int64_t e = a.corr_off[p] + tid; e < a.corr_off[p + 1]; e += FT)
        W[a.corr_rows[e] + (size_t)a.corr_cols[e] * ld] += a.corr_vals[e];
      __syncthreads();
    }
    pc.mark(PH_GATHER);

    // ---- two-level blocked Gauss-Jordan -----------------------------------------------------
    for (int k0 = 0; k0 < n; k0 += NBO) {
      const int nbo = (n - k0) < NBO ? (n - k0) : NBO;
      if (tid == 0) s_misc[1] = 0;
      for (int q0 = k0; q0 < k0 + nbo; q0 += NB) {
        const int nb = (k0 + nbo - q0) < NB ? (k0 + nbo - q0) : NB;
        // 1. inner panel to shared memory, NB Gauss-Jordan steps
        for (int c = 0; c < nb; ++c)
          for (int r = tid; r < n; r += FT) panel[r + c * ldp] = W[r + (size_t)(q0 + c) * ld];
        __syncthreads();
        pc.mark(PH_PANEL_LOAD);
        for (int j = 0; j < nb; ++j) {
          const int kj = q0 + j;
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
              if (bi != kj) {                       // remember the swap for the far columns
                const int s = s_misc[1]++;
                swapA[s] = kj;
                swapB[s] = bi;
              }
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
        pc.mark(PH_INNER_GJ);
        // 2. the other columns of the outer block: this panel's swaps, R aside, zero pivot rows
        if (tid < nbo) {
          const int c = k0 + tid;
          const bool inpanel = c >= q0 && c < q0 + nb;
          double* col = W + (size_t)c * ld;
          if (!inpanel) {
            for (int j = 0; j < nb; ++j) {
              const int kj = q0 + j, pv = ipiv[kj];
              if (pv != kj) {
                const double t0 = col[kj];
                col[kj] = col[pv];
                col[pv] = t0;
              }
            }
          }
#pragma unroll
          for (int t = 0; t < NB; ++t) {
            double v = 0.0;
            if (!inpanel && t < nb) { v = col[q0 + t]; col[q0 + t] = 0.0; }
            Rs[tid * NB + t] = v;
          }
        }
        __syncthreads();
        //    rank-NB update of those columns, one row per thread
        for (int r = tid; r < n; r += FT) {
          double N[NB];
#pragma unroll
          for (int t = 0; t < NB; ++t) N[t] = (t < nb) ? panel[r + t * ldp] : 0.0;
          double* Wr = W + r + (size_t)k0 * ld;
          for (int ci0 = 0; ci0 < nbo; ci0 += UC) {
            double w[UC];
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const int cg = k0 + ci0 + u;
              const bool ok = (ci0 + u < nbo) && !(cg >= q0 && cg < q0 + nb);
              w[u] = ok ? Wr[(size_t)(ci0 + u) * ld] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const double* __restrict__ rs = Rs + (ci0 + u) * NB;
#pragma unroll
              for (int t = 0; t < NB; ++t) w[u] = fma(N[t], rs[t], w[u]);
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) {
              const int cg = k0 + ci0 + u;
              if ((ci0 + u < nbo) && !(cg >= q0 && cg < q0 + nb)) Wr[(size_t)(ci0 + u) * ld] = w[u];
            }
          }
        }
        // panel back
        for (int c = 0; c < nb; ++c)
          for (int r = tid; r < n; r += FT) W[r + (size_t)(q0 + c) * ld] = panel[r + c * ldp];
        __syncthreads();
        pc.mark(PH_INBLOCK);
      }

      // ---- far update: every column outside the outer block ---------------------------------
      const int nfar = n - nbo;
      if (nfar > 0) {
        const int nswap = s_misc[1];
        // 3. swaps in order (rare for these matrices), then R = W[KO, far] set aside k-major and
        //    the pivot rows zeroed.  The 64 x 64 blocks of R are transposed through shared memory
        //    so that both the column-major reads of W and the k-major writes of Rg are coalesced
        //    (a thread-per-column walk cost as much as the whole rank-64 update).
        if (nswap > 0) {
          // the swaps of one column are dependent; a thread interleaves those of 4 columns
          for (int f0 = tid; f0 < nfar; f0 += 4 * FT) {
            double* col[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int f = f0 + u * FT;
              col[u] = f < nfar ? W + (size_t)(f < k0 ? f : f + nbo) * ld : nullptr;
            }
            for (int s = 0; s < nswap; ++s) {
              const int ra = swapA[s], rb = swapB[s];
              double va[4], vb[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (col[u]) { va[u] = col[u][ra]; vb[u] = col[u][rb]; }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (col[u]) { col[u][ra] = vb[u]; col[u][rb] = va[u]; }
            }
          }
          __syncthreads();
        }
        {
          double* Tt = front;                        // 64 x 65 transposition tile
          for (int fc0 = 0; fc0 < nfar; fc0 += TC) {
#pragma unroll
            for (int it = 0; it < NBO * TC / FT; ++it) {
              const int idx = tid + it * FT;
              const int k = idx & (NBO - 1), cc = idx / NBO;
              const int f = fc0 + cc;
              double v = 0.0;
              if (k < nbo && f < nfar) {
                double* ptr = W + (size_t)(f < k0 ? f : f + nbo) * ld + k0 + k;
                v = *ptr;
                *ptr = 0.0;
              }
              Tt[cc * (NBO + 1) + k] = v;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < NBO * TC / FT; ++it) {
              const int idx = tid + it * FT;
              const int cc = idx & (TC - 1), k = idx / TC;
              Rg[(size_t)k * ldr + fc0 + cc] = Tt[cc * (NBO + 1) + k];
            }
            __syncthreads();
          }
        }
        pc.mark(PH_FAR_SWAP);
        // 4. W[:, far] += W[:, KO] * R on the FP64 tensor cores: 256 x 64 tiles of W per CTA
        //    iteration, one 32 x 32 sub-tile per warp = 4 x 4 mma.m8n8k4 tiles, accumulators
        //    (32 doubles per lane) in registers.  The product is formed transposed, D[c][r] =
        //    sum_k R[k][c] N[r][k], so that a lane's two accumulator entries are two consecutive
        //    rows of one column of the column-major W (one 128-bit access).  Shared-memory rows
        //    are padded by 8 doubles: the 4 k x 8 index pattern of a fragment load is then
        //    conflict free.
        const int lane = tid & 31, warp = tid >> 5;
        const int g = lane >> 2, t4 = lane & 3;
        const int wr = warp & 7, wc = warp >> 3;           // 8 row blocks x 2 column blocks of 32
        for (int rt0 = 0; rt0 < n; rt0 += TR) {
          __syncthreads();                           // previous tile's readers are done with Ns
          for (int idx = tid; idx < NBO * TR; idx += FT) {
            const int k = idx / TR, r = idx - k * TR;
            Ns[k * NS_LD + r] = (k < nbo && rt0 + r < n) ? W[rt0 + r + (size_t)(k0 + k) * ld] : 0.0;
          }
          pc.mark(PH_FAR_NS);
          // R tiles are double-buffered: chunk c+1 is fetched with cp.async while chunk c is used
          auto fetch_r = [&](double* dst, int fc) {
#pragma unroll
            for (int it = 0; it < NBO * TC / 2 / FT; ++it) {
              const int idx = tid + it * FT;
              const int k = idx / (TC / 2), c2 = idx - k * (TC / 2);
              __pipeline_memcpy_async(dst + k * RS_LD + 2 * c2, Rg + (size_t)k * ldr + fc + 2 * c2, 16);
            }
            __pipeline_commit();
          };
          fetch_r(Rsm, 0);
          int buf = 0;
          const int row_w = rt0 + wr * 32 + 2 * t4;          // + 8 * nb
          for (int fc0 = 0; fc0 < nfar; fc0 += TC, buf ^= 1) {
            __pipeline_wait_prior(0);                // this thread's part of the current chunk landed
            __syncthreads();                         // everyone's did; Ns written; old readers done
            pc.mark(PH_FAR_RS);
            if (fc0 + TC < nfar) fetch_r(Rsm + (buf ^ 1) * NBO * RS_LD, fc0 + TC);
            double acc[4][4][2];
            double* cptr[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
              const int f = fc0 + wc * 32 + mb * 8 + g;
              cptr[mb] = (f < nfar) ? W + (size_t)(f < k0 ? f : f + nbo) * ld + row_w : nullptr;
#pragma unroll
              for (int nb4 = 0; nb4 < 4; ++nb4) {
                const int r = row_w + 8 * nb4;
                if (cptr[mb] && r + 1 < n) {
                  const double2 v = *reinterpret_cast<const double2*>(cptr[mb] + 8 * nb4);
                  acc[mb][nb4][0] = v.x;
                  acc[mb][nb4][1] = v.y;
                } else {
                  acc[mb][nb4][0] = (cptr[mb] && r < n) ? cptr[mb][8 * nb4] : 0.0;
                  acc[mb][nb4][1] = 0.0;
                }
              }
            }
            // pull the next chunk's W tile towards L2 while this one is multiplied: the update sits at
            // the ridge of the roofline and its loads are otherwise exposed after the barrier
            if (fc0 + TC < nfar) {
              const int f = fc0 + TC + wc * 32 + lane;
              if (f < nfar) {
                const double* nxt = W + (size_t)(f < k0 ? f : f + nbo) * ld + rt0 + wr * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 16));
              }
            }
            const double* __restrict__ rs = Rsm + buf * NBO * RS_LD + t4 * RS_LD + wc * 32 + g;
            const double* __restrict__ ns = Ns + t4 * NS_LD + wr * 32 + g;
#pragma unroll 2
            for (int kk = 0; kk < NBO; kk += 4) {
              double af[4], bf[4];
#pragma unroll
              for (int mb = 0; mb < 4; ++mb) af[mb] = rs[kk * RS_LD + mb * 8];       // R[kk + t4][col]
#pragma unroll
              for (int nb4 = 0; nb4 < 4; ++nb4) bf[nb4] = ns[kk * NS_LD + nb4 * 8];  // N[row][kk + t4]
#pragma unroll
              for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb4 = 0; nb4 < 4; ++nb4)
                  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                               : "+d"(acc[mb][nb4][0]), "+d"(acc[mb][nb4][1])
                               : "d"(af[mb]), "d"(bf[nb4]));
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
              if (!cptr[mb]) continue;
#pragma unroll
              for (int nb4 = 0; nb4 < 4; ++nb4) {
                const int r = row_w + 8 * nb4;
                if (r + 1 < n)
                  *reinterpret_cast<double2*>(cptr[mb] + 8 * nb4) = make_double2(acc[mb][nb4][0], acc[mb][nb4][1]);
                else if (r < n)
                  cptr[mb][8 * nb4] = acc[mb][nb4][0];
              }
            }
            pc.mark(PH_FAR_MMA);
          }
        }
        __syncthreads();
        pc.mark(PH_FAR_MMA);
      }
    }

    // ---- undo the pivoting (column swaps in reverse) and write the tiled apply layout --------
    int* src = sd;               // the front region is free again
    if (tid == 0) {
      for (int c = 0; c < n; ++c) src[c] = c;
      for (int k = n - 1; k >= 0; --k) {
        const int pv = ipiv[k];
        if (pv != k) { const int t0 = src[k]; src[k] = src[pv]; src[pv] = t0; }
      }
    }
    __syncthreads();
    if (a.sepoff) {
      // condensed set: only the separator block of the inverse is kept, in the same tiled layout
      const int64_t so = a.sepoff[p];
      const int ns = (int)(a.sepoff[p + 1] - so);
      const int32_t* __restrict__ sl = a.seplocal + so;
      double* out = a.store + a.soff[p];
      for (int row0 = 0; row0 < ns; row0 += ALFIB_TILE_ROWS) {
        int rows = ns - row0;
        rows = rows > ALFIB_TILE_ROWS ? ALFIB_TILE_ROWS : ((rows + 1) & ~1);
        double* tile = out + (size_t)row0 * ns;
        for (int64_t i = tid; i < (int64_t)rows * ns; i += FT) {
          const int c = (int)(i / rows), rr = (int)(i - (int64_t)c * rows);
          const int r = row0 + rr;
          tile[i] = (r < ns) ? W[sl[r] + (size_t)src[sl[c]] * ld] : 0.0;
        }
      }
    } else {
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
    pc.mark(PH_PACK);
    pc.flush(a.timing);
  }
}

size_t factor_smem_bytes(int maxn, int NB) {
  size_t doubles = front_doubles(maxn, NB) + NB + 34;
  size_t ints = 34 + 2 * NBO + 2 + maxn + 2;
  size_t bytes = doubles * sizeof(double) + ints * sizeof(int);
  // lookup tables alias the front region: 2*maxn ints must fit in it
  const size_t front_bytes = front_doubles(maxn, NB) * sizeof(double);
  if (front_bytes < 2 * (size_t)maxn * sizeof(int)) bytes += 2 * (size_t)maxn * sizeof(int) - front_bytes;
  return bytes + 16;
}

template <int NB>
void run_factor(alfib_ctx* c, FactorArgs a, size_t smem, int grid) {
  static_assert(NBO % NB == 0, "inner panel must divide the outer block");
  CUDA_TRY(cudaFuncSetAttribute(patch_factor_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  patch_factor_kernel<NB><<<grid, FT, smem, c->stream>>>(a);
  c->launches++;
  CUDA_TRY(cudaGetLastError());
}

}  // namespace

void launch_patch_factor(alfib_ctx* c, const Level& L, PatchSet& ps, const double* vals) {
  if (ps.npatch == 0 || ps.maxn == 0) { ps.factored = true; return; }
  // Schur-complement setup of a condensed set: the block kernel first (it leaves C = A_Nk A_kk^-1 A_kN per block),
  // then this kernel on the separators only
  const bool schur = ps.cond.on && ps.cond.schur;
  if (schur) {
    launch_condense_blocks(c, L, ps, vals, true);
    if (ps.cond.h.maxsep == 0) { ps.factored = true; return; }
  }
  const int maxn = schur ? ps.cond.h.maxsep : ps.maxn;
  const size_t budget = 220 * 1024;
  int NB = 16;
  while (NB > 4 && factor_smem_bytes(maxn, NB) > budget) NB >>= 1;
  if (factor_smem_bytes(maxn, NB) > budget)
    throw DeviceError{ALFIB_EINVAL, "patch too large for the shared-memory panel (n = " + std::to_string(maxn) + ")"};
  const size_t smem = factor_smem_bytes(maxn, NB);
  const int ldmax = roundup2(maxn);
  const int64_t slot = (int64_t)maxn * ldmax + (int64_t)NBO * ((maxn + TC - 1) & ~(TC - 1));
  const int grid = std::min(ps.npatch, c->num_sms);          // >= 160 KB of shared memory: one CTA per SM
  if (c->fwork.n < (size_t)grid * slot) c->fwork.alloc((size_t)grid * slot);   // grow-only: shared by all patch sets
  c->finfo.alloc(2);
  CUDA_TRY(cudaMemsetAsync(c->finfo.p, 0, 2 * sizeof(int), c->stream));

  FactorArgs a;
  a.npatch = ps.npatch;
  a.forder = schur ? ps.cond.sforder.p : ps.forder.p;
  a.poff = schur ? ps.cond.sepoff.p : ps.off.p;
  a.pdofs = schur ? ps.cond.sepdofs.p : ps.dofs.p;
  a.sorted = schur ? ps.cond.sepsorted.p : ps.sorted.p;
  a.sperm = schur ? ps.cond.sepperm.p : ps.sperm.p;
  a.soff = ps.cond.on ? ps.cond.ssoff.p : ps.soff.p;
  a.store = ps.store;
  a.sepoff = (ps.cond.on && !schur) ? ps.cond.sepoff.p : nullptr;
  a.seplocal = (ps.cond.on && !schur) ? ps.cond.seplocal.p : nullptr;
  a.sc_blk_start = schur ? ps.cond.blk_start.p : nullptr;
  a.sc_nb_off = ps.cond.nb_off.p;
  a.sc_nb_pos = ps.cond.nb_pos.p;
  a.sc_upos = ps.cond.sc_upos.p;
  a.sc_inst_c = ps.cond.sc_inst_c.p;
  a.sc_inst_ld = ps.cond.sc_inst_ld.p;
  a.sc_cbuf = ps.cond.cbuf.p;
  a.corr_off = nullptr;
  a.corr_rows = a.corr_cols = nullptr;
  a.corr_vals = nullptr;
  if (ps.has_corr) {
    ALFIB_REQUIRE(!ps.cond.on, "patch corrections need dense patch inverses (no patch blocks)");
    ALFIB_REQUIRE(ps.corr_fresh, "alfib_level_set_patch_correction_values has not been called since the values changed");
    a.corr_off = ps.corr_off.p;
    a.corr_rows = ps.corr_rows.p;
    a.corr_cols = ps.corr_cols.p;
    a.corr_vals = ps.corr_vals.p;
  }
  a.bs = L.bs;
  a.rowptr = L.rowptr.p;
  a.colidx = L.colidx.p;
  a.vals = vals;
  a.work = c->fwork.p;
  a.slot_elems = slot;
  a.maxn = maxn;
  a.counter = c->finfo.p + 1;
  a.info = c->finfo.p;
  a.timing = nullptr;
  static const bool want_timing = getenv("ALFIB_FACTOR_TIMING") != nullptr;
  DBuf<long long> tbuf;
  if (want_timing) {
    tbuf.alloc(PH_COUNT);
    CUDA_TRY(cudaMemsetAsync(tbuf.p, 0, PH_COUNT * sizeof(long long), c->stream));
    a.timing = tbuf.p;
  }
  switch (NB) {
    case 16: run_factor<16>(c, a, smem, grid); break;
    case 8: run_factor<8>(c, a, smem, grid); break;
    default: run_factor<4>(c, a, smem, grid); break;
  }
  int info[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(info, c->finfo.p, sizeof(info), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (want_timing) {
    long long t[PH_COUNT];
    CUDA_TRY(cudaMemcpy(t, tbuf.p, sizeof(t), cudaMemcpyDeviceToHost));
    static const char* names[PH_COUNT] = {"gather", "panel_load", "inner_gj", "inblock", "far_swap", "far_Ns",
                                          "far_Rs", "far_mma", "pack"};
    double tot = 0;
    for (int i = 0; i < PH_COUNT; ++i) tot += (double)t[i];
    fprintf(stderr, "[alfib factor timing] npatch=%d maxn=%d NB=%d grid=%d:", ps.npatch, maxn, NB, grid);
    for (int i = 0; i < PH_COUNT; ++i) fprintf(stderr, " %s %.1f%%", names[i], 100.0 * t[i] / tot);
    fprintf(stderr, "  (sum %.1f Mcycles per CTA)\n", tot / grid / 1e6);
    tbuf.release();
  }
  if (info[0] != 0)
    throw DeviceError{ALFIB_ESINGULAR, "patch " + std::to_string(info[0] - 1) + " is singular"};
  if (ps.cond.on && !schur) launch_condense_blocks(c, L, ps, vals);
  ps.factored = true;
}

void patch_extract_inverse(alfib_ctx* c, const PatchSet& ps, int patch, double* host_out) {
  if (ps.cond.on) {
    condensed_extract_inverse(c, ps, patch, host_out);
    return;
  }
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
