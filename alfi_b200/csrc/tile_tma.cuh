// Third version of the tile op: the matrix stream goes through shared memory with 1-D TMA bulk copies.
//
// The condensed PCApply_PATCH (condense.cu) is a pure stream of small column-major tiles (6 ... 100 KB each,
// 2.55 GB per application on the finest ldc3d level).  v1/v2 read them with per-lane 128-bit loads: 80 registers,
// 24 warps per SM, ~40 % warps active and `long_scoreboard` as the only stall — a latency-bound LDG stream at 0.69 of
// the copy bandwidth.  Here the bytes in flight do not live in registers: one persistent CTA per SM, every warp
// owns a ring of STAGES shared-memory stages filled by `cp.async.bulk.shared::cluster.global` (SASS UBLKCP) that
// complete on an mbarrier, so 8 warps x STAGES x 8 KB (~190 KB per SM, 28 MB chip-wide) are always outstanding
// with one issuing lane per warp.  A warp consumes its stages in order (LDS.128, lanes own row pairs exactly as in
// v2, FP64 FMA) and refills a stage as soon as it has read it; the ring runs across op boundaries, so there is no
// ramp-down between ops.  Ops are taken from an atomic counter (the first three per warp statically), the next op's
// descriptor and the op after next's number are requested one op ahead, the gathered source values of a stage are
// requested when the stage is filled (its index list one stage earlier) — nothing in the steady state waits on a
// dependent global load.
//
// A stage holds a run of whole columns of one op: 32 columns if a column is <= 256 bytes (<= 32 rows), else 16
// (<= 64 rows = 512 bytes), so a stage never straddles one of the 32-entry chunks the source values are held in.
// Same arithmetic per (row, column group) as v2 up to the order in which the columns of a row are added:
// v2 adds even/odd ring slots in two chains, v3 adds stage by stage; results agree to rounding, and are bitwise
// reproducible run to run in deterministic mode (static order inside an op; ops write disjoint rows per launch).
#pragma once
// (A second version — ops in static byte-balanced ranges per warp, descriptors fetched 32 at a time, indices two
// stages ahead, 4 FMA chains — measured 0.58-0.69 ms against 0.506 ms for this one on cfg5 and was dropped: the
// larger code did not pay.  profiles/apply_variants_r2.txt)

namespace tma {

constexpr int STAGES = 3;
constexpr int META_INTS = 12;   // per stage: id, c0, cnt, nrows, ncols, flags, row (2), priv (2), pad (2)
// CFG 0: 8 warps x 3 stages x 8 KB (first version); CFG 1: 16 warps x 3 stages x 4 KB — the same bytes in flight per SM
// behind twice as many independent instruction streams (ncu of CFG 0: the warps spend 64 % of their samples in
// long-scoreboard / fixed-latency stalls of their own serial chain, not waiting for the copies)
template <int CFG> struct Cfg;
template <> struct Cfg<0> { static constexpr int WARPS = 8, STAGE_BYTES = 8192; };
template <> struct Cfg<1> { static constexpr int WARPS = 16, STAGE_BYTES = 4096; };
template <int CFG>
constexpr size_t smem_bytes() {
  return (size_t)Cfg<CFG>::WARPS * STAGES * Cfg<CFG>::STAGE_BYTES + (size_t)Cfg<CFG>::WARPS * STAGES * 8 +
         (size_t)Cfg<CFG>::WARPS * STAGES * META_INTS * 4;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

struct OpRegs {                       // the fields of a TileOp a warp keeps in registers
  long long mat, col, row, priv;
  int nrows, ncols, flags;
};
__device__ __forceinline__ OpRegs load_op(const TileOp* __restrict__ ops, int id) {
  const int4* p = reinterpret_cast<const int4*>(ops + id);   // 48 bytes, 16-byte aligned
  const int4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  OpRegs o;
  o.mat = ((long long)(unsigned)a.x) | ((long long)a.y << 32);
  o.col = ((long long)(unsigned)a.z) | ((long long)a.w << 32);
  o.row = ((long long)(unsigned)b.x) | ((long long)b.y << 32);
  o.priv = ((long long)(unsigned)b.z) | ((long long)b.w << 32);
  o.nrows = c.x;
  o.ncols = c.y;
  o.flags = c.z;
  return o;
}

// columns cnt of a stage (in shared memory, column j at j * colbytes) times the source values held by the lanes.
// FULL = the number of columns of a full stage for this row count and stage size: known at compile time, so the
// common case is straight-line code (measured: a run-time trip count costs 10 % of the whole application).
template <int G, int FULL>
__device__ __forceinline__ void consume_stage(const unsigned char* __restrict__ sp, int colbytes, int cnt, double xs,
                                              bool active, int grp, double& acc0, double& acc1) {
  if (cnt == FULL) {
#pragma unroll
    for (int jj = 0; jj < FULL / G; ++jj) {
      const int j = jj * G + grp;
      const double xc = __shfl_sync(0xffffffffu, xs, j);
      if (active) {
        const double2 a = *reinterpret_cast<const double2*>(sp + j * colbytes);
        acc0 = fma(a.x, xc, acc0);
        acc1 = fma(a.y, xc, acc1);
      }
    }
  } else {
    for (int jj = 0; jj * G < cnt; ++jj) {
      const int j = jj * G + grp;
      const double xc = __shfl_sync(0xffffffffu, xs, j & 31);
      if (active && j < cnt) {
        const double2 a = *reinterpret_cast<const double2*>(sp + j * colbytes);
        acc0 = fma(a.x, xc, acc0);
        acc1 = fma(a.y, xc, acc1);
      }
    }
  }
}

}  // namespace tma

template <bool ATOMIC, bool ACCUM, int MODE, int CFG, bool RED>
__global__ void __launch_bounds__(tma::Cfg<CFG>::WARPS * 32, 1)
    tile_ops_kernel_tma(const TileOp* __restrict__ ops, int nops, const int32_t* __restrict__ cidx,
                        const double* __restrict__ store, const double* __restrict__ srcA,
                        const double* __restrict__ srcB, PeerOut yout, double* __restrict__ dstB, const FusedSrc fs,
                        unsigned* __restrict__ counter) {
  using namespace tma;
  constexpr int WARPS = Cfg<CFG>::WARPS, STAGE_BYTES = Cfg<CFG>::STAGE_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  double* __restrict__ y = resolve(yout);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = gridDim.x * WARPS, gw = blockIdx.x * WARPS + warp;
  unsigned char* ring = smem + (size_t)warp * STAGES * STAGE_BYTES;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar_u32 = smem_u32(smem + (size_t)WARPS * STAGES * STAGE_BYTES) + warp * STAGES * 8;
  int* meta = reinterpret_cast<int*>(smem + (size_t)WARPS * STAGES * STAGE_BYTES + (size_t)WARPS * STAGES * 8) +
              warp * STAGES * META_INTS;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar_u32 + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const uint64_t pol = policy_evict_first();

  auto value_at = [&](int e, bool valid) -> double {
    if (!valid) return 0.0;
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
  // columns per stage: the largest of 32 / 16 / 8 that fits (a column is roundup2(nrows) doubles, <= 512 bytes)
  auto stage_cols = [](int nrows) {
    const int colbytes = ((nrows + 1) >> 1) * 16;
    return 32 * colbytes <= STAGE_BYTES ? 32 : (16 * colbytes <= STAGE_BYTES ? 16 : 8);
  };

  // ---- producer state (uniform over the warp) -------------------------------------------------------------
  int id0 = gw, id1 = gw + W;                  // current / next op; the one after next is in flight in lane 0
  int idp = gw + 2 * W;
  OpRegs d0, d1;
  d0.mat = d0.col = d0.row = d0.priv = 0; d0.nrows = d0.ncols = d0.flags = 0;
  d1 = d0;
  if (id0 < nops) d0 = load_op(ops, id0);
  if (id1 < nops) d1 = load_op(ops, id1);
  int p_c = 0;                                 // next column of the current op to be requested
  // source index of this lane for the chunk to be produced next (requested one produce step earlier)
  int e_cur = (id0 < nops && lane < min(stage_cols(d0.nrows), d0.ncols)) ? __ldg(cidx + d0.col + lane) : 0;
  double xs[STAGES];

  auto produce = [&](int s) {
    // op switch: everything needed here was requested one op ago
    // (exactly one counter increment per processed op: atomicInc wraps to 0 after the nops-th, so the counter
    //  resets itself for the next launch)
    if (id0 < nops && p_c >= d0.ncols && !(p_c == 0 && d0.ncols == 0)) {
      id0 = id1;
      d0 = d1;
      id1 = __shfl_sync(0xffffffffu, idp, 0);
      if (id1 < nops) d1 = load_op(ops, id1);
      if (lane == 0) idp = 3 * W + (int)atomicInc(counter, (unsigned)(nops - 1));
      p_c = 0;
    }
    int* m = meta + s * META_INTS;
    if (id0 >= nops) {                          // no more work: sentinel
      if (lane == 0) m[0] = -1;
      xs[s] = 0.0;
      return;
    }
    const int colbytes = ((d0.nrows + 1) >> 1) * 16;
    const int cps = stage_cols(d0.nrows);
    const int c0 = p_c, cnt = min(cps, d0.ncols - c0);
    const uint32_t bytes = (uint32_t)(cnt * colbytes);
    if (lane == 0) {
      m[0] = id0; m[1] = c0; m[2] = cnt; m[3] = d0.nrows; m[4] = d0.ncols; m[5] = d0.flags;
      *reinterpret_cast<long long*>(m + 6) = d0.row;
      *reinterpret_cast<long long*>(m + 8) = d0.priv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage before the async write
      mbar_expect_tx(bar_u32 + s * 8, bytes);
      if (bytes) bulk_g2s(ring_u32 + s * STAGE_BYTES, store + d0.mat + (long long)c0 * (colbytes >> 3), bytes, bar_u32 + s * 8, pol);
    }
    // source values of this stage (index requested one step ago); index of the next stage
    xs[s] = value_at(e_cur, lane < cnt);
    p_c = c0 + cnt;
    if (d0.ncols == 0) p_c = 1;                 // an op without columns is one empty stage
    if (p_c < d0.ncols) {
      e_cur = (lane < min(cps, d0.ncols - p_c)) ? __ldg(cidx + d0.col + p_c + lane) : 0;
    } else if (id1 < nops) {
      e_cur = (lane < min(stage_cols(d1.nrows), d1.ncols)) ? __ldg(cidx + d1.col + lane) : 0;
    } else {
      e_cur = 0;
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES; ++s) produce(s);
  __syncwarp();

  // ---- consumer ------------------------------------------------------------------------------------------------
  double acc0 = 0.0, acc1 = 0.0;
  int ri0 = 0, ri1 = 0;
  uint32_t parity = 0;
  bool done = false;
  while (!done) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      if (done) break;
      const int* m = meta + s * META_INTS;
      const int id = m[0];
      if (id < 0) { done = true; break; }
      const int c0 = m[1], cnt = m[2], nrows = m[3], ncols = m[4];
      const long long row = *reinterpret_cast<const long long*>(m + 6), pv = *reinterpret_cast<const long long*>(m + 8);
      const int half = (nrows + 1) >> 1, colbytes = half * 16;
      const int G = half <= 8 ? 4 : (half <= 16 ? 2 : 1);
      const int LPG = 32 / G;
      const int grp = lane / LPG, l = lane - grp * LPG;
      const bool active = l < half;
      if (c0 == 0) {
        acc0 = 0.0;
        acc1 = 0.0;
        if (RED && row >= 0 && grp == 0 && active) {      // destination dofs of this lane's two rows, requested early
          ri0 = __ldg(cidx + row + 2 * l);
          ri1 = (2 * l + 1 < nrows) ? __ldg(cidx + row + 2 * l + 1) : 0;
        }
      }
      mbar_wait(bar_u32 + s * 8, parity);
      const unsigned char* sp = ring + s * STAGE_BYTES + (active ? l : 0) * 16;
      // full stages: 32 columns of <= 16 rows; 32 (8 KB stages) / 16 (4 KB) columns of <= 32 rows; 16 / 8 of <= 64 rows
      constexpr int FULL2 = STAGE_BYTES >= 8192 ? 32 : 16, FULL1 = STAGE_BYTES >= 8192 ? 16 : 8;
      if (G == 4) consume_stage<4, 32>(sp, colbytes, cnt, xs[s], active, grp, acc0, acc1);
      else if (G == 2) consume_stage<2, FULL2>(sp, colbytes, cnt, xs[s], active, grp, acc0, acc1);
      else consume_stage<1, FULL1>(sp, colbytes, cnt, xs[s], active, grp, acc0, acc1);
      if (c0 + cnt >= ncols) {                  // last stage of the op: sum the column groups (fixed order), write
        double r0 = acc0, r1 = acc1;
        for (int off = 16; off >= LPG; off >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, off);
          r1 += __shfl_xor_sync(0xffffffffu, r1, off);
        }
        if (grp == 0 && active) {
          const int r = 2 * l;
          if (pv >= 0) {
            double* __restrict__ priv = dstB + pv;
            if (ACCUM) {
              atomicAdd(priv + r, r0);
              if (r + 1 < nrows) atomicAdd(priv + r + 1, r1);
            } else {
              priv[r] = r0;
              if (r + 1 < nrows) priv[r + 1] = r1;
            }
          }
          if (RED && row >= 0) {
            // one reduction (RED.ADD.F64: no return value, no dependent load) per destination.  Launches without
            // ATOMIC add to every destination at most once (patches of a colour share no dof, the tiles of a patch own
            // disjoint rows, distinct shared blocks are disjoint), so the result is that of y[i] += r, bit for bit.
            atomicAdd(y + ri0, r0);
            if (r + 1 < nrows) atomicAdd(y + ri1, r1);
          } else if (row >= 0) {
            const int32_t* __restrict__ ri = cidx + row;
            if (ATOMIC) {
              atomicAdd(y + ri[r], r0);
              if (r + 1 < nrows) atomicAdd(y + ri[r + 1], r1);
            } else {
              y[ri[r]] += r0;
              if (r + 1 < nrows) y[ri[r + 1]] += r1;
            }
          }
        }
      }
      __syncwarp();                             // every lane has read stage s and its meta
      produce(s);
      __syncwarp();
    }
    parity ^= 1;
  }
}
