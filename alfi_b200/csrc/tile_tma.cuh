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
// descriptor and the op after next's number are requested one op ahead, a stage's slice of the source index list
// two stages before the stage is filled and its gathered source values when it is filled (two stages before they
// are used) — nothing in the steady state waits on a dependent global load (first version: index one stage ahead;
// ncu showed 38 % of the samples in long-scoreboard stalls on exactly that chain).
//
// A stage holds a run of whole columns of one op: 32 columns if a column is <= 256 bytes (<= 32 rows), else 16
// (<= 64 rows = 512 bytes), so a stage never straddles one of the 32-entry chunks the source values are held in.
// Same arithmetic per (row, column group) as v2 up to the order in which the columns of a row are added:
// v2 adds even/odd ring slots in two chains, v3 adds stage by stage; results agree to rounding, and are bitwise
// reproducible run to run in deterministic mode (static order inside an op; ops write disjoint rows per launch).
#pragma once

namespace tma {

constexpr int WARPS = 8;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 8192;
constexpr int META_INTS = 12;   // per stage: id, c0, cnt, nrows, ncols, flags, row (2), priv (2), pad (2)
constexpr size_t smem_bytes() {
  return (size_t)WARPS * STAGES * STAGE_BYTES + (size_t)WARPS * STAGES * 8 + (size_t)WARPS * STAGES * META_INTS * 4;
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

// columns cnt of a stage (in shared memory, column j at j * colbytes) times the source values held by the lanes;
// even / odd trips go to separate accumulator pairs (two independent FMA chains per row)
template <int G>
__device__ __forceinline__ void consume_stage(const unsigned char* __restrict__ sp, int colbytes, int cnt, double xs,
                                              bool active, int grp, double (&acc)[4]) {
  if (cnt == 32 || (cnt == 16 && G == 1)) {
    // full stages: fixed trip count (a multiple of 8)
    const int trips = cnt / G;
#pragma unroll 8
    for (int jj = 0; jj < trips; jj += 2) {
      const int j0 = jj * G + grp, j1 = j0 + G;
      const double x0 = __shfl_sync(0xffffffffu, xs, j0);
      const double x1 = __shfl_sync(0xffffffffu, xs, j1);
      if (active) {
        const double2 a = *reinterpret_cast<const double2*>(sp + j0 * colbytes);
        const double2 b = *reinterpret_cast<const double2*>(sp + j1 * colbytes);
        acc[0] = fma(a.x, x0, acc[0]);
        acc[1] = fma(a.y, x0, acc[1]);
        acc[2] = fma(b.x, x1, acc[2]);
        acc[3] = fma(b.y, x1, acc[3]);
      }
    }
  } else {
    for (int jj = 0; jj * G < cnt; ++jj) {
      const int j = jj * G + grp;
      const double xc = __shfl_sync(0xffffffffu, xs, j & 31);
      if (active && j < cnt) {
        const double2 a = *reinterpret_cast<const double2*>(sp + j * colbytes);
        acc[0] = fma(a.x, xc, acc[0]);
        acc[1] = fma(a.y, xc, acc[1]);
      }
    }
  }
}

}  // namespace tma

template <bool ATOMIC, bool ACCUM, int MODE>
__global__ void __launch_bounds__(tma::WARPS * 32, 1)
    tile_ops_kernel_tma(const TileOp* __restrict__ ops, int nops, const int32_t* __restrict__ cidx,
                        const double* __restrict__ store, const double* __restrict__ srcA,
                        const double* __restrict__ srcB, PeerOut yout, double* __restrict__ dstB, const FusedSrc fs,
                        unsigned* __restrict__ counter) {
  using namespace tma;
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

  // ---- the front: walks the warp's chunk sequence LOOK chunks ahead of the copies ----------------------------
  // A chunk = a run of whole columns of one op that fills one stage.  The front takes ops (the first three per warp
  // statically, then from the counter), keeps the next op's descriptor and the number of the op after next in
  // flight, and requests every chunk's slice of the source index list when it generates the chunk; the copy, the
  // gathered source values (which need the indices) and the stage's meta data follow LOOK produce steps later, so the
  // chain op number -> descriptor -> index -> value never waits in the steady state.
  struct Chunk {
    const double* src;     // first column of the chunk in the store
    long long row, priv;
    int id, c0, cnt, nrows, ncols;   // id < 0: no more work
    int e;                 // this lane's source index (valid for lane < cnt)
  };
  int id0 = gw, id1 = gw + W, idp = gw + 2 * W;
  OpRegs d0, d1;
  d0.mat = d0.col = d0.row = d0.priv = 0; d0.nrows = d0.ncols = d0.flags = 0;
  d1 = d0;
  if (id0 < nops) d0 = load_op(ops, id0);
  if (id1 < nops) d1 = load_op(ops, id1);
  int f_c = 0;                                 // next column of the front's op
  auto front = [&]() -> Chunk {
    Chunk ch;
    // op switch (exactly one counter increment per processed op: atomicInc wraps to 0 after the nops-th, so the
    // counter resets itself for the next launch)
    if (id0 < nops && f_c >= d0.ncols && !(f_c == 0 && d0.ncols == 0)) {
      id0 = id1;
      d0 = d1;
      id1 = __shfl_sync(0xffffffffu, idp, 0);
      if (id1 < nops) d1 = load_op(ops, id1);
      if (lane == 0) idp = 3 * W + (int)atomicInc(counter, (unsigned)(nops - 1));
      f_c = 0;
    }
    if (id0 >= nops) {
      ch.src = nullptr; ch.row = ch.priv = -1; ch.id = -1; ch.c0 = ch.cnt = ch.nrows = ch.ncols = 0; ch.e = 0;
      return ch;
    }
    const int half = (d0.nrows + 1) >> 1;
    const int cps = d0.nrows <= 32 ? 32 : 16;
    ch.id = id0; ch.c0 = f_c; ch.cnt = min(cps, d0.ncols - f_c); ch.nrows = d0.nrows; ch.ncols = d0.ncols;
    ch.row = d0.row; ch.priv = d0.priv;
    ch.src = store + d0.mat + (long long)f_c * (half * 2);
    ch.e = lane < ch.cnt ? __ldg(cidx + d0.col + f_c + lane) : 0;
    f_c += ch.cnt;
    if (d0.ncols == 0) f_c = 1;                 // an op without columns is one empty chunk
    return ch;
  };
  constexpr int LOOK = 2;
  Chunk q[LOOK];
#pragma unroll
  for (int i = 0; i < LOOK; ++i) q[i] = front();
  double xs[STAGES];

  auto produce = [&](int s) {
    const Chunk ch = q[0];
#pragma unroll
    for (int i = 0; i + 1 < LOOK; ++i) q[i] = q[i + 1];
    q[LOOK - 1] = front();
    int* m = meta + s * META_INTS;
    if (ch.id < 0) {                            // no more work: sentinel
      if (lane == 0) m[0] = -1;
      xs[s] = 0.0;
      return;
    }
    const int colbytes = ((ch.nrows + 1) >> 1) * 16;
    const uint32_t bytes = (uint32_t)(ch.cnt * colbytes);
    if (lane == 0) {
      m[0] = ch.id; m[1] = ch.c0; m[2] = ch.cnt; m[3] = ch.nrows; m[4] = ch.ncols;
      *reinterpret_cast<long long*>(m + 6) = ch.row;
      *reinterpret_cast<long long*>(m + 8) = ch.priv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage before the async write
      mbar_expect_tx(bar_u32 + s * 8, bytes);
      if (bytes) bulk_g2s(ring_u32 + s * STAGE_BYTES, ch.src, bytes, bar_u32 + s * 8, pol);
    }
    xs[s] = value_at(ch.e, lane < ch.cnt);      // consumed STAGES - 1 stages later
  };

#pragma unroll
  for (int s = 0; s < STAGES; ++s) produce(s);
  __syncwarp();

  // ---- consumer ------------------------------------------------------------------------------------------------
  double acc[4] = {0.0, 0.0, 0.0, 0.0};        // rows (2l, 2l+1) x even / odd column trips: independent FMA chains
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
      if (c0 == 0) { acc[0] = acc[1] = acc[2] = acc[3] = 0.0; }
      mbar_wait(bar_u32 + s * 8, parity);
      const unsigned char* sp = ring + s * STAGE_BYTES + (active ? l : 0) * 16;
      if (G == 4) consume_stage<4>(sp, colbytes, cnt, xs[s], active, grp, acc);
      else if (G == 2) consume_stage<2>(sp, colbytes, cnt, xs[s], active, grp, acc);
      else consume_stage<1>(sp, colbytes, cnt, xs[s], active, grp, acc);
      if (c0 + cnt >= ncols) {                  // last stage of the op: sum the column groups (fixed order), write
        double r0 = acc[0] + acc[2], r1 = acc[1] + acc[3];
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
          if (row >= 0) {
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
