// CPU-only checker of the condensed patch sets (test infrastructure, not product code).
//
// Compiles alfi_b200/csrc/condense_host.h — the very code libalfib.so uses to turn a block
// structure into storage layout, index lists and tile-op lists — together with a plain host
// restatement of what the CUDA kernels of condense.cu / patch_factor.cu do with them:
//   ch_factor : X_SS from a pivoted Gauss-Jordan inverse of the whole patch; per block the gather
//               through the sorted key tables, Gauss-Jordan inverse of A_kk, V = A_Nk D, [D | -W]
//   ch_apply  : K1 (V ops) -> K2 (separator rhs) -> K3 (X_SS ops) [-> K3b (z sums, shared form)] -> K4 ([D | -W] ops)
// tests/test_condense_host.py drives it through ctypes and compares with dense patch solves.
#include <cstring>

#include "condense_host.h"

namespace {

struct Shim {
  CondensedHost cd;
  int npatch, ncolour, bs, ndofs;
  std::vector<int64_t> off;
  std::vector<int32_t> dofs, order, colour, rowptr, colidx;
  std::vector<double> store, g1, rs, us, z;
  std::string err;
};

// in-place Gauss-Jordan inverse with partial (row) pivoting, first-max rule; column-major, ld
bool gj_inverse(double* M, int n, int ld) {
  std::vector<int> piv(n);
  std::vector<double> prow(n), fcol(n);
  for (int j = 0; j < n; ++j) {
    int bi = j;
    double best = -1.0;
    for (int r = j; r < n; ++r)
      if (std::fabs(M[r + (size_t)j * ld]) > best) { best = std::fabs(M[r + (size_t)j * ld]); bi = r; }
    if (!(best > 0.0)) return false;
    piv[j] = bi;
    if (bi != j)
      for (int c = 0; c < n; ++c) std::swap(M[j + (size_t)c * ld], M[bi + (size_t)c * ld]);
    const double d = 1.0 / M[j + (size_t)j * ld];
    for (int c = 0; c < n; ++c) prow[c] = (c == j) ? 0.0 : M[j + (size_t)c * ld] * d;
    for (int r = 0; r < n; ++r) fcol[r] = M[r + (size_t)j * ld];
    for (int c = 0; c < n; ++c)
      for (int r = 0; r < n; ++r) {
        double v;
        if (r == j) v = (c == j) ? d : prow[c];
        else if (c == j) v = -fcol[r] * d;
        else v = M[r + (size_t)c * ld] - fcol[r] * prow[c];
        M[r + (size_t)c * ld] = v;
      }
  }
  for (int k = n - 1; k >= 0; --k)           // undo the row pivoting: column swaps in reverse
    if (piv[k] != k)
      for (int r = 0; r < n; ++r) std::swap(M[r + (size_t)k * ld], M[r + (size_t)piv[k] * ld]);
  return true;
}

int bsearch_i32(const int32_t* a, int n, int key) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] == key) return mid;
    if (a[mid] < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

void run_op(const Shim& s, const TileOp& op, const double* srcA, const double* srcB, double* y, double* dstB) {
  const int rt = ch_roundup2(op.nrows);
  const double* T = s.store.data() + op.mat;
  const int32_t* ci = s.cd.cidx.data() + op.col;
  std::vector<double> acc(op.nrows, 0.0);
  for (int c = 0; c < op.ncols; ++c) {
    const int e = ci[c];
    const double xc = e >= 0 ? srcA[e] : srcB[~e];
    for (int r = 0; r < op.nrows; ++r) acc[r] += T[(size_t)c * rt + r] * xc;
  }
  for (int r = 0; r < op.nrows; ++r) {
    if (op.priv >= 0) {
      if (s.cd.any_accum && (&op >= s.cd.opsS.data() && &op < s.cd.opsS.data() + s.cd.opsS.size()))
        dstB[op.priv + r] += acc[r];          // X_SS lists with column chunks: us is zeroed and accumulated
      else
        dstB[op.priv + r] = acc[r];
    }
    if (op.row >= 0) y[s.cd.cidx[op.row + r]] += acc[r];
  }
}

}  // namespace

extern "C" {

void* ch_create(int n_nodes, int bs, const int32_t* rowptr, const int32_t* colidx, int npatch, const int64_t* off,
                const int32_t* dofs, int norder, const int32_t* order, const int32_t* colour, int ncolour,
                const int32_t* blocks, int allow_shared, int split_wide, char* err, int errlen) {
  Shim* s = new Shim();
  s->npatch = npatch;
  s->ncolour = ncolour;
  s->bs = bs;
  s->ndofs = n_nodes * bs;
  s->off.assign(off, off + npatch + 1);
  s->dofs.assign(dofs, dofs + off[npatch]);
  s->order.assign(order, order + norder);
  s->colour.assign(colour, colour + npatch);
  s->rowptr.assign(rowptr, rowptr + n_nodes + 1);
  s->colidx.assign(colidx, colidx + rowptr[n_nodes]);
  PatchView pv{npatch, ncolour, bs, s->ndofs, s->off.data(), s->dofs.data(), &s->order, s->colour.data(),
               s->rowptr.data(), s->colidx.data()};
  try {
    build_condensed_host(pv, blocks, s->cd, allow_shared != 0, split_wide != 0);
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
    delete s;
    return nullptr;
  }
  s->store.assign((size_t)std::max<int64_t>(s->cd.store_elems, 1), 0.0);
  s->g1.assign((size_t)std::max<int64_t>(s->cd.g1_total, 1), 0.0);
  s->rs.assign((size_t)std::max<int64_t>(s->cd.nsep_total, 1), 0.0);
  s->us.assign((size_t)std::max<int64_t>(s->cd.nsep_total, 1), 0.0);
  s->z.assign((size_t)std::max<int64_t>(s->cd.g1_total, 1), 0.0);
  return s;
}

void ch_destroy(void* h) { delete static_cast<Shim*>(h); }

// stats[0..10] = store_elems, index_bytes, nblocks, nsep_total, maxb, maxm, maxsep, #ops, shared, ndist, any_accum
void ch_stats(void* h, int64_t* stats) {
  const CondensedHost& cd = static_cast<Shim*>(h)->cd;
  stats[0] = cd.store_elems;
  stats[1] = cd.index_bytes;
  stats[2] = cd.nblocks;
  stats[3] = cd.nsep_total;
  stats[4] = cd.maxb;
  stats[5] = cd.maxm;
  stats[6] = cd.maxsep;
  stats[7] = (int64_t)(cd.opsV.size() + cd.opsS.size() + cd.opsDW.size());
  stats[8] = cd.shared ? 1 : 0;
  stats[9] = cd.ndist;
  stats[10] = cd.any_accum ? 1 : 0;
}

// vals: nnzb x bs x bs row-major blocks.  Returns 0, or 1 + index of a singular patch / block.
int ch_factor(void* h, const double* vals) {
  Shim& s = *static_cast<Shim*>(h);
  const CondensedHost& cd = s.cd;
  const int bs = s.bs, b2 = bs * bs;
  std::vector<int32_t> pos(s.ndofs, -1);
  for (int p = 0; p < s.npatch; ++p) {
    const int64_t o = s.off[p];
    const int n = (int)(s.off[p + 1] - o);
    if (n == 0) continue;
    const int32_t* I = s.dofs.data() + o;
    for (int l = 0; l < n; ++l) pos[I[l]] = l;
    std::vector<double> W((size_t)n * n, 0.0);          // column-major
    for (int l = 0; l < n; ++l) {
      const int node = I[l] / bs, comp = I[l] % bs;
      for (int k = s.rowptr[node]; k < s.rowptr[node + 1]; ++k)
        for (int c2 = 0; c2 < bs; ++c2) {
          const int cl = pos[s.colidx[k] * bs + c2];
          if (cl >= 0) W[l + (size_t)cl * n] = vals[(size_t)k * b2 + comp * bs + c2];
        }
    }
    for (int l = 0; l < n; ++l) pos[I[l]] = -1;
    if (!gj_inverse(W.data(), n, n)) return 1 + p;
    const int64_t so = cd.sepoff[p];
    const int ns = (int)(cd.sepoff[p + 1] - so);
    const int32_t* sl = cd.seplocal.data() + so;
    double* out = s.store.data() + cd.ssoff[p];
    for (int row0 = 0; row0 < ns; row0 += ALFIB_TILE_ROWS) {
      const int rows = std::min(ns - row0, ALFIB_TILE_ROWS), rt = ch_roundup2(rows);
      double* tile = out + (size_t)row0 * ns;
      for (int c = 0; c < ns; ++c)
        for (int r = 0; r < rt; ++r) tile[(size_t)c * rt + r] = r < rows ? W[sl[row0 + r] + (size_t)sl[c] * n] : 0.0;
    }
  }
  // what condense_blocks_kernel is launched over: the instances, or the distinct blocks in shared form
  const std::vector<BlockDesc>& fblocks = cd.shared ? cd.sblocks : cd.blocks;
  const std::vector<int32_t>& fdofs = cd.shared ? cd.sdofs : cd.bdofs;
  const std::vector<int32_t>& fkeys = cd.shared ? cd.skeys : cd.bkeys;
  const std::vector<int32_t>& fperm = cd.shared ? cd.sperm : cd.bperm;
  for (int64_t q = 0; q < (int64_t)fblocks.size(); ++q) {
    const BlockDesc& d = fblocks[q];
    const int b = d.b, m = d.m, bm = b + m;
    const int32_t* gd = fdofs.data() + d.dofs;
    const int32_t* keys = fkeys.data() + d.keys;
    const int32_t* perm = fperm.data() + d.keys;
    std::vector<double> Akk((size_t)b * b, 0.0), AkN((size_t)b * std::max(m, 1), 0.0), ANk((size_t)std::max(m, 1) * b, 0.0);
    for (int rp = 0; rp < bm; ++rp) {
      const int node = gd[rp] / bs, comp = gd[rp] % bs;
      for (int k = s.rowptr[node]; k < s.rowptr[node + 1]; ++k)
        for (int c2 = 0; c2 < bs; ++c2) {
          const int hit = bsearch_i32(keys, bm, s.colidx[k] * bs + c2);
          if (hit < 0) continue;
          const int cp = perm[hit];
          const double v = vals[(size_t)k * b2 + comp * bs + c2];
          if (rp < b) {
            if (cp < b) Akk[rp + (size_t)cp * b] = v; else AkN[rp + (size_t)(cp - b) * b] = v;
          } else if (cp < b) {
            ANk[(rp - b) + (size_t)cp * m] = v;
          }
        }
    }
    if (!gj_inverse(Akk.data(), b, b)) return 1 + (int)q;
    const int mr = ch_roundup2(m), br = ch_roundup2(b);
    double* Vt = s.store.data() + d.voff;
    for (int c = 0; c < b; ++c)
      for (int r = 0; r < mr; ++r) {
        double v = 0.0;
        if (r < m)
          for (int k = 0; k < b; ++k) v += ANk[r + (size_t)k * m] * Akk[k + (size_t)c * b];
        Vt[(size_t)c * mr + r] = v;
      }
    double* Dt = s.store.data() + d.dwoff;
    for (int c = 0; c < b; ++c)
      for (int r = 0; r < br; ++r) Dt[(size_t)c * br + r] = r < b ? Akk[r + (size_t)c * b] * d.dscale : 0.0;
    double* Wt = Dt + (size_t)br * b;
    for (int c = 0; c < m; ++c)
      for (int r = 0; r < br; ++r) {
        double v = 0.0;
        if (r < b)
          for (int k = 0; k < b; ++k) v += Akk[r + (size_t)k * b] * AkN[k + (size_t)c * b];
        Wt[(size_t)c * br + r] = -v;
      }
  }
  return 0;
}

// The Schur-complement setup (ALFIB_SCHUR_SETUP; condense_host.h build_schur_lists) as the two kernels run it:
//   block kernel  : per factor block the gather through the key tables, pivoted Gauss-Jordan of A_kk with A_kN
//                   carried through the elimination (Ws = A_kk^-1 A_kN with solve accuracy), V = A_Nk D, [D | -Ws],
//                   C = A_Nk Ws into the scratch buffer;
//   factor kernel : per patch the gather of A_SS through the sorted separator tables, minus every instance's C
//                   entries at nb_pos, pivoted Gauss-Jordan inverse, X_SS tiles.
int ch_factor_schur(void* h, const double* vals) {
  Shim& s = *static_cast<Shim*>(h);
  const CondensedHost& cd = s.cd;
  SchurHost sh;
  build_schur_lists(cd, s.npatch, sh);
  const int bs = s.bs, b2 = bs * bs;
  std::vector<double> cbuf((size_t)std::max<int64_t>(sh.ctotal, 1), 0.0);
  const std::vector<BlockDesc>& fblocks = cd.shared ? cd.sblocks : cd.blocks;
  const std::vector<int32_t>& fdofs = cd.shared ? cd.sdofs : cd.bdofs;
  const std::vector<int32_t>& fkeys = cd.shared ? cd.skeys : cd.bkeys;
  const std::vector<int32_t>& fperm = cd.shared ? cd.sperm : cd.bperm;
  for (int64_t q = 0; q < (int64_t)fblocks.size(); ++q) {
    const BlockDesc& d = fblocks[q];
    const int b = d.b, m = d.m, bm = b + m;
    const int32_t* gd = fdofs.data() + d.dofs;
    const int32_t* keys = fkeys.data() + d.keys;
    const int32_t* perm = fperm.data() + d.keys;
    std::vector<double> Akk((size_t)b * b, 0.0), AkN((size_t)b * std::max(m, 1), 0.0), ANk((size_t)std::max(m, 1) * b, 0.0);
    for (int rp = 0; rp < bm; ++rp) {
      const int node = gd[rp] / bs, comp = gd[rp] % bs;
      for (int k = s.rowptr[node]; k < s.rowptr[node + 1]; ++k)
        for (int c2 = 0; c2 < bs; ++c2) {
          const int hit = bsearch_i32(keys, bm, s.colidx[k] * bs + c2);
          if (hit < 0) continue;
          const int cp = perm[hit];
          const double v = vals[(size_t)k * b2 + comp * bs + c2];
          if (rp < b) {
            if (cp < b) Akk[rp + (size_t)cp * b] = v; else AkN[rp + (size_t)(cp - b) * b] = v;
          } else if (cp < b) {
            ANk[(rp - b) + (size_t)cp * m] = v;
          }
        }
    }
    // in-place pivoted Gauss-Jordan of Akk, the same row operations applied to AkN
    {
      std::vector<int> piv(b);
      std::vector<double> prow(b), fcol(b), prowN(std::max(m, 1));
      for (int j = 0; j < b; ++j) {
        int bi = j;
        double best = -1.0;
        for (int r = j; r < b; ++r)
          if (std::fabs(Akk[r + (size_t)j * b]) > best) { best = std::fabs(Akk[r + (size_t)j * b]); bi = r; }
        if (!(best > 0.0)) return 1 + (int)q;
        piv[j] = bi;
        if (bi != j) {
          for (int c = 0; c < b; ++c) std::swap(Akk[j + (size_t)c * b], Akk[bi + (size_t)c * b]);
          for (int c = 0; c < m; ++c) std::swap(AkN[j + (size_t)c * b], AkN[bi + (size_t)c * b]);
        }
        const double dinv = 1.0 / Akk[j + (size_t)j * b];
        for (int c = 0; c < b; ++c) prow[c] = (c == j) ? 0.0 : Akk[j + (size_t)c * b] * dinv;
        for (int r = 0; r < b; ++r) fcol[r] = Akk[r + (size_t)j * b];
        for (int c = 0; c < m; ++c) prowN[c] = AkN[j + (size_t)c * b] * dinv;
        for (int c = 0; c < b; ++c)
          for (int r = 0; r < b; ++r) {
            double v;
            if (r == j) v = (c == j) ? dinv : prow[c];
            else if (c == j) v = -fcol[r] * dinv;
            else v = Akk[r + (size_t)c * b] - fcol[r] * prow[c];
            Akk[r + (size_t)c * b] = v;
          }
        for (int c = 0; c < m; ++c)
          for (int r = 0; r < b; ++r)
            AkN[r + (size_t)c * b] = (r == j) ? prowN[c] : AkN[r + (size_t)c * b] - fcol[r] * prowN[c];
      }
      for (int k = b - 1; k >= 0; --k)
        if (piv[k] != k)
          for (int r = 0; r < b; ++r) std::swap(Akk[r + (size_t)k * b], Akk[r + (size_t)piv[k] * b]);
    }
    const int mr = ch_roundup2(m), br = ch_roundup2(b);
    double* Vt = s.store.data() + d.voff;
    for (int c = 0; c < b; ++c)
      for (int r = 0; r < mr; ++r) {
        double v = 0.0;
        if (r < m)
          for (int k = 0; k < b; ++k) v += ANk[r + (size_t)k * m] * Akk[k + (size_t)c * b];
        Vt[(size_t)c * mr + r] = v;
      }
    double* Dt = s.store.data() + d.dwoff;
    for (int c = 0; c < b; ++c)
      for (int r = 0; r < br; ++r) Dt[(size_t)c * br + r] = r < b ? Akk[r + (size_t)c * b] * d.dscale : 0.0;
    double* Wt = Dt + (size_t)br * b;
    for (int c = 0; c < m; ++c)
      for (int r = 0; r < br; ++r) Wt[(size_t)c * br + r] = r < b ? -AkN[r + (size_t)c * b] : 0.0;
    double* C = cbuf.data() + sh.coff[q];
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < m; ++i) {
        double v = 0.0;
        for (int k = 0; k < b; ++k) v += ANk[i + (size_t)k * m] * AkN[k + (size_t)j * b];
        C[i + (size_t)j * m] = v;
      }
  }
  for (int p = 0; p < s.npatch; ++p) {
    const int64_t so = cd.sepoff[p];
    const int ns = (int)(cd.sepoff[p + 1] - so);
    if (ns == 0) continue;
    const int32_t* I = cd.sepdofs.data() + so;
    const int32_t* sd = sh.sepsorted.data() + so;
    const int32_t* sp = sh.sepperm.data() + so;
    std::vector<double> W((size_t)ns * ns, 0.0);
    for (int r = 0; r < ns; ++r) {
      const int node = I[r] / bs, comp = I[r] % bs;
      for (int k = s.rowptr[node]; k < s.rowptr[node + 1]; ++k)
        for (int c2 = 0; c2 < bs; ++c2) {
          const int hit = bsearch_i32(sd, ns, s.colidx[k] * bs + c2);
          if (hit >= 0) W[r + (size_t)sp[hit] * ns] = vals[(size_t)k * b2 + comp * bs + c2];
        }
    }
    for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q) {
      const int64_t o2 = cd.nb_off[q];
      const int mq = (int)(cd.nb_off[q + 1] - o2);
      const double* C = cbuf.data() + sh.inst_c[q];
      const int ldc = sh.inst_ld[q];
      for (int jj = 0; jj < mq; ++jj)
        for (int ii = 0; ii < mq; ++ii)
          W[cd.nb_pos[o2 + ii] + (size_t)cd.nb_pos[o2 + jj] * ns] -= C[sh.upos[o2 + ii] + (size_t)sh.upos[o2 + jj] * ldc];
    }
    if (!gj_inverse(W.data(), ns, ns)) return 1 + p;
    double* out = s.store.data() + cd.ssoff[p];
    for (int row0 = 0; row0 < ns; row0 += ALFIB_TILE_ROWS) {
      const int rows = std::min(ns - row0, ALFIB_TILE_ROWS), rt = ch_roundup2(rows);
      double* tile = out + (size_t)row0 * ns;
      for (int c = 0; c < ns; ++c)
        for (int r = 0; r < rt; ++r) tile[(size_t)c * rt + r] = r < rows ? W[(row0 + r) + (size_t)c * ns] : 0.0;
    }
  }
  return 0;
}

// the whole store (X_SS tiles, V, [D | -W] tiles) — to compare the two setups
void ch_store(void* h, double* out) {
  Shim& s = *static_cast<Shim*>(h);
  std::memcpy(out, s.store.data(), sizeof(double) * (size_t)s.cd.store_elems);
}

// y += sum_i R_i^T A_i^-1 R_i x through the op lists (y is NOT zeroed, like the device kernel)
void ch_apply(void* h, const double* x, double* y) {
  Shim& s = *static_cast<Shim*>(h);
  const CondensedHost& cd = s.cd;
  for (const TileOp& op : cd.opsV) run_op(s, op, x, nullptr, nullptr, s.g1.data());
  for (int64_t e = 0; e < cd.nsep_total; ++e) {
    double v = x[cd.sepdofs[e]];
    for (int j = cd.cptr[e]; j < cd.cptr[e + 1]; ++j) v -= s.g1[cd.cg1[j]];
    s.rs[e] = v;
  }
  if (cd.any_accum) std::fill(s.us.begin(), s.us.end(), 0.0);
  for (int col = 0; col < s.ncolour; ++col)
    for (int i = cd.s_colour_start[col]; i < cd.s_colour_start[col + 1]; ++i)
      run_op(s, cd.opsS[i], s.rs.data(), nullptr, y, s.us.data());
  if (cd.shared) {
    // K3b: z[k] = sum of the visited instances' us entries; K4: one [visits D | -Wf] op per distinct block
    for (int64_t e = 0; e < cd.g1_total; ++e) {
      double v = 0.0;
      for (int j = cd.zptr[e]; j < cd.zptr[e + 1]; ++j) v += s.us[cd.zsrc[j]];
      s.z[e] = v;
    }
    for (const TileOp& op : cd.opsDW) run_op(s, op, x, s.z.data(), y, nullptr);
    return;
  }
  for (int col = 0; col < s.ncolour; ++col)
    for (int i = cd.dw_colour_start[col]; i < cd.dw_colour_start[col + 1]; ++i)
      run_op(s, cd.opsDW[i], x, s.us.data(), y, nullptr);
}

// 0 if, within every colour, no two ops of a phase write the same entry of y (the deterministic mode
// of launch_condensed_apply uses plain read-modify-write stores inside a colour), and the private
// slots (g1, us) are written by exactly one op each; otherwise a positive code.
int ch_check_disjoint(void* h) {
  Shim& s = *static_cast<Shim*>(h);
  const CondensedHost& cd = s.cd;
  std::vector<int> stamp(s.ndofs, -1);
  int tick = 0;
  auto phase = [&](const std::vector<TileOp>& ops, const std::vector<int>& start) {
    for (int col = 0; col < s.ncolour; ++col, ++tick)
      for (int i = start[col]; i < start[col + 1]; ++i)
        for (int r = 0; r < ops[i].nrows; ++r) {
          const int g = cd.cidx[ops[i].row + r];
          if (stamp[g] == tick) return false;
          stamp[g] = tick;
        }
    return true;
  };
  if (!cd.any_accum && !phase(cd.opsS, cd.s_colour_start)) return 1;   // column chunks share rows (atomics)
  if (!phase(cd.opsDW, cd.dw_colour_start)) return 2;
  std::vector<char> hit((size_t)std::max<int64_t>(cd.g1_total, 1), 0);
  for (const TileOp& op : cd.opsV)
    for (int r = 0; r < op.nrows; ++r) {
      if (hit[op.priv + r]) return 3;
      hit[op.priv + r] = 1;
    }
  for (int64_t i = 0; i < cd.g1_total; ++i)
    if (!hit[i]) return 4;
  return 0;
}

// dense inverse of one patch (row-major n x n) rebuilt from the condensed factors
void ch_inverse(void* h, int patch, double* out) {
  Shim& s = *static_cast<Shim*>(h);
  const int n = (int)(s.off[patch + 1] - s.off[patch]);
  auto fetch = [&](int64_t off, int64_t count) {
    return std::vector<double>(s.store.begin() + off, s.store.begin() + off + count);
  };
  condensed_inverse_host(s.cd, patch, n, fetch, out);
}

}  // extern "C"
