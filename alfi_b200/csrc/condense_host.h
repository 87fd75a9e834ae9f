// Host side of the condensed patch sets (see condense.cu for the method): block structure ->
// storage layout, index lists and tile-op lists.  CUDA-free on purpose, so that the same code is
// compiled into libalfib.so (condense.cu) and into the CPU-only checker tests/condense_host_shim.cpp,
// which executes the op lists on the host against dense patch solves (tests/test_condense_host.py).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef ALFIB_TILE_ROWS
#define ALFIB_TILE_ROWS 64
#endif

// One dense tile operation of the condensed apply: dst[rows] (+)= M * src[cols].
struct TileOp {
  long long mat;    // element offset of the tile in the store: column-major, roundup2(nrows) rows per column
  long long col;    // offset into cidx of the ncols source indices (i >= 0: srcA[i]; i < 0: srcB[~i])
  long long row;    // offset into cidx of the nrows global destination dofs, or -1
  long long priv;   // offset into the private destination buffer (plain store), or -1
  int nrows, ncols; // nrows <= ALFIB_TILE_ROWS
  int flags;        // TILEOP_ACCUM: the private destination is accumulated (atomicAdd) instead of stored
  int pad_;
};
enum { TILEOP_ACCUM = 1 };
#ifndef ALFIB_SPLIT_COLS
#define ALFIB_SPLIT_COLS 256   // X_SS tiles wider than 2x this are cut into column chunks (one op each)
#endif

// Per (patch, block) descriptor of the per-Newton-step block setup (D = A_kk^-1, V = A_Nk D, W = D A_kN)
struct BlockDesc {
  long long dofs;   // offset into bdofs: b block dofs, then m neighbour (separator) dofs, global numbers
  long long keys;   // offset into bkeys / bperm: the same b+m dofs sorted ascending + their positions
  long long voff;   // store offset of the V tile   (roundup2(m) x b)
  long long dwoff;  // store offset of the [D | -W] tile (roundup2(b) x (b+m))
  int b, m;
  double dscale;    // the D part of the tile is stored multiplied by this (1 unless the block is shared)
};

// what the builder needs to know about a patch set and its level (all host memory, borrowed)
struct PatchView {
  int npatch, ncolour, bs, ndofs;
  const int64_t* off;          // npatch+1
  const int32_t* dofs;
  const std::vector<int32_t>* order;
  const int32_t* colour;       // npatch
  const int32_t* rowptr;       // BSR pattern of the level (block rows)
  const int32_t* colidx;
};

struct CondensedHost {
  int maxb = 0, maxm = 0, maxsep = 0;
  int64_t nblocks = 0, nsep_total = 0, g1_total = 0, store_elems = 0;
  std::vector<int64_t> sepoff;          // npatch+1: separator dofs per patch ...
  std::vector<int32_t> seplocal;        // ... as patch-local indices
  std::vector<int32_t> sepdofs;         // ... and as global dofs
  std::vector<int64_t> ssoff;           // npatch+1: store offsets of the X_SS tiles
  std::vector<int64_t> blk_start;       // npatch+1: blocks of patch p are [blk_start[p], blk_start[p+1])
  std::vector<BlockDesc> blocks;
  std::vector<int64_t> bl_off;          // nblocks+1 into bl_local
  std::vector<int32_t> bl_local;        // patch-local indices of the block dofs
  std::vector<int64_t> nb_off;          // nblocks+1 into nb_pos
  std::vector<int32_t> nb_pos;          // positions of the neighbour dofs in the patch's separator list
  std::vector<int64_t> g1off;           // nblocks+1: private slots of the V outputs
  std::vector<int32_t> cidx, bdofs, bkeys, bperm, cptr, cg1;
  std::vector<TileOp> opsV, opsS, opsDW;
  std::vector<int> s_colour_start, dw_colour_start;   // ncolour+1 ranges of opsS / opsDW
  int64_t index_bytes = 0;              // bytes of index data one apply reads (roofline accounting)
  bool any_accum = false;               // some X_SS ops are column chunks: us is zeroed per apply and accumulated

  // ---- shared blocks -----------------------------------------------------------------------------------
  // A block with a given dof set occurs in several patches (a macro cell lies in the macro stars of all
  // its vertices) and D_k = A_kk^-1 is the same matrix in all of them.  With U_k the union of the
  // separator neighbourhoods N_q of its instances q, V_q / W_q are row / column subsets of
  //     Vf_k = A_Uk D_k  (|U_k| x |B_k|),     Wf_k = D_k A_kU  (|B_k| x |U_k|),
  // so the additive sum over the patches can be taken block by block instead of instance by instance:
  //     g[k]    = Vf_k x[B_k]                                          (one op per distinct block)
  //     rs[p,s] = x[S_p[s]] - sum over the instances q of p with s in N_q of g[k(q)][pos of s in U_k]
  //     us[p]   = X_SS rs[p],  y[S_p] += us[p]                         (unchanged)
  //     z[k]    = sum over the visited instances q of k of us[p(q)] scattered to U_k positions
  //     y[B_k] += [visits_k D_k | -Wf_k] [x[B_k]; z[k]]                (one op per distinct block)
  // which stores and streams Vf, Wf and D once per distinct block (3-D SV k=3: 60x45, 45x60, 45x45 per
  // macro cell instead of 4 x (30x45 + 45x30 + 45x45)).  Used when the distinct blocks are pairwise
  // disjoint and every |U_k| <= ALFIB_TILE_ROWS; otherwise the per-instance form above is kept.
  bool shared = false;
  int64_t ndist = 0;
  std::vector<int64_t> inst_dist;       // nblocks: distinct block of each instance
  std::vector<int32_t> inst_opos;       // parallel to bl_local: position of the dof in the owner's block order
  std::vector<int32_t> inst_upos;       // parallel to nb_pos: position of the neighbour in U_k
  std::vector<BlockDesc> sblocks;       // ndist descriptors, dofs = [B_k (owner order); U_k (ascending)]
  std::vector<int32_t> sdofs, skeys, sperm, svisits;
  std::vector<int64_t> uoff;            // ndist+1: slots of g[k] / z[k] in the private buffers
  std::vector<int32_t> zptr, zsrc;      // CSR: z slot -> positions in us, in iteration order
};

inline int ch_roundup2(int n) { return (n + 1) & ~1; }

// Shared-block form (see CondensedHost): rewrites the V / [D | -W] op lists, the K2 lists, the store
// layout behind the X_SS tiles and the private-buffer sizes of a finished per-instance build.  Leaves
// `cd` untouched (shared = false) when the structure does not allow it.
inline void build_shared_blocks(const PatchView& pv, CondensedHost& cd) {
  if (cd.nblocks == 0) return;
  const int npatch = pv.npatch;
  // ---- distinct blocks: instances with the same dof set; the first instance is the owner ---------------
  std::vector<int64_t> inst_dist(cd.nblocks, -1), owner;
  {
    std::map<std::vector<int32_t>, int64_t> seen;
    std::vector<int32_t> key;
    for (int64_t q = 0; q < cd.nblocks; ++q) {
      const BlockDesc& d = cd.blocks[q];
      key.assign(cd.bdofs.begin() + d.dofs, cd.bdofs.begin() + d.dofs + d.b);
      std::sort(key.begin(), key.end());
      auto it = seen.find(key);
      if (it == seen.end()) {
        it = seen.emplace(key, (int64_t)owner.size()).first;
        owner.push_back(q);
      }
      inst_dist[q] = it->second;
    }
  }
  const int64_t nd = (int64_t)owner.size();
  // the ops of different distinct blocks run in one launch with plain stores: they must not overlap
  std::vector<int32_t> opos_of_dof(pv.ndofs, -1);
  for (int64_t k = 0; k < nd; ++k) {
    const BlockDesc& d = cd.blocks[owner[k]];
    for (int i = 0; i < d.b; ++i) {
      int32_t& st = opos_of_dof[cd.bdofs[d.dofs + i]];
      if (st >= 0) return;
      st = i;
    }
  }
  // ---- U_k: union of the instances' separator neighbourhoods, global dofs ascending ---------------------
  std::vector<std::vector<int32_t>> U(nd);
  for (int64_t q = 0; q < cd.nblocks; ++q) {
    const BlockDesc& d = cd.blocks[q];
    std::vector<int32_t>& u = U[inst_dist[q]];
    u.insert(u.end(), cd.bdofs.begin() + d.dofs + d.b, cd.bdofs.begin() + d.dofs + d.b + d.m);
  }
  int64_t utotal = 0;
  for (int64_t k = 0; k < nd; ++k) {
    std::sort(U[k].begin(), U[k].end());
    U[k].erase(std::unique(U[k].begin(), U[k].end()), U[k].end());
    if ((int)U[k].size() > ALFIB_TILE_ROWS) return;
    utotal += (int64_t)U[k].size();
  }
  if (utotal >= INT_MAX / 2) return;

  // ---- from here on the shared form is used ------------------------------------------------------------
  cd.shared = true;
  cd.ndist = nd;
  cd.inst_dist = inst_dist;
  cd.inst_opos.resize(cd.bl_local.size());
  cd.inst_upos.resize(cd.nb_pos.size());
  for (int64_t q = 0; q < cd.nblocks; ++q) {
    const BlockDesc& d = cd.blocks[q];
    const std::vector<int32_t>& u = U[inst_dist[q]];
    for (int i = 0; i < d.b; ++i) cd.inst_opos[cd.bl_off[q] + i] = opos_of_dof[cd.bdofs[d.dofs + i]];
    for (int j = 0; j < d.m; ++j)
      cd.inst_upos[cd.nb_off[q] + j] =
          (int32_t)(std::lower_bound(u.begin(), u.end(), cd.bdofs[d.dofs + d.b + j]) - u.begin());
  }
  cd.svisits.assign(nd, 0);
  for (int32_t p : *pv.order)
    for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q) cd.svisits[inst_dist[q]]++;

  // store: X_SS tiles (unchanged), then Vf and [visits D | -Wf] of every distinct block
  int64_t cur = cd.ssoff[npatch];
  cd.sblocks.resize(nd);
  cd.uoff.assign(nd + 1, 0);
  cd.maxm = 0;
  std::vector<int64_t> sblk_list(nd);
  std::vector<int32_t> idx;
  int64_t list_entries = 0;
  for (int64_t k = 0; k < nd; ++k) {
    const BlockDesc& od = cd.blocks[owner[k]];
    const int b = od.b, m = (int)U[k].size();
    BlockDesc& d = cd.sblocks[k];
    d.b = b;
    d.m = m;
    d.voff = cur;
    cur += (int64_t)ch_roundup2(m) * b;
    d.dwoff = cur;
    cur += (int64_t)ch_roundup2(b) * (b + m);
    d.dscale = (double)std::max(cd.svisits[k], 1);
    d.dofs = (int64_t)cd.sdofs.size();
    d.keys = d.dofs;
    cd.uoff[k + 1] = cd.uoff[k] + m;
    cd.maxm = std::max(cd.maxm, m);
    sblk_list[k] = (int64_t)cd.cidx.size();
    for (int i = 0; i < b; ++i) {
      const int32_t g = cd.bdofs[od.dofs + i];
      cd.cidx.push_back(g);
      cd.sdofs.push_back(g);
    }
    for (int t = 0; t < m; ++t) {
      cd.cidx.push_back(~(int32_t)(cd.uoff[k] + t));
      cd.sdofs.push_back(U[k][t]);
    }
    list_entries += b + m;
    idx.resize(b + m);
    std::iota(idx.begin(), idx.end(), 0);
    const int32_t* gd = cd.sdofs.data() + d.dofs;
    std::sort(idx.begin(), idx.end(), [&](int x, int y) { return gd[x] < gd[y]; });
    for (int i = 0; i < b + m; ++i) {
      cd.skeys.push_back(gd[idx[i]]);
      cd.sperm.push_back(idx[i]);
    }
  }
  cd.g1_total = cd.uoff[nd];
  cd.store_elems = cur;

  // K2: contributions to every separator entry, now slots of g[k]
  {
    std::vector<int32_t> cursor(cd.cptr.begin(), cd.cptr.end() - 1);
    for (int p = 0; p < npatch; ++p)
      for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q)
        for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j)
          cd.cg1[cursor[cd.sepoff[p] + cd.nb_pos[j]]++] = (int32_t)(cd.uoff[inst_dist[q]] + cd.inst_upos[j]);
  }
  // z[k]: for every slot the us entries of the visited instances, in iteration order
  cd.zptr.assign(cd.g1_total + 1, 0);
  for (int32_t p : *pv.order)
    for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q)
      for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j) cd.zptr[cd.uoff[inst_dist[q]] + cd.inst_upos[j] + 1]++;
  for (int64_t e = 0; e < cd.g1_total; ++e) cd.zptr[e + 1] += cd.zptr[e];
  cd.zsrc.assign((size_t)cd.zptr[cd.g1_total], 0);
  {
    std::vector<int32_t> cursor(cd.zptr.begin(), cd.zptr.end() - 1);
    for (int32_t p : *pv.order)
      for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q)
        for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j)
          cd.zsrc[cursor[cd.uoff[inst_dist[q]] + cd.inst_upos[j]]++] = (int32_t)(cd.sepoff[p] + cd.nb_pos[j]);
  }
  // op lists: one V op and one [D | -W] op per visited distinct block; the S ops stay as they are.
  // All [D | -W] ops sit in the range of colour 0 (they are pairwise disjoint: one launch, plain stores).
  cd.opsV.clear();
  cd.opsDW.clear();
  for (int64_t k = 0; k < nd; ++k) {
    const BlockDesc& d = cd.sblocks[k];
    if (cd.svisits[k] == 0) continue;
    if (d.m > 0) cd.opsV.push_back(TileOp{d.voff, sblk_list[k], -1, cd.uoff[k], d.m, d.b, 0, 0});
    cd.opsDW.push_back(TileOp{d.dwoff, sblk_list[k], sblk_list[k], -1, d.b, d.b + d.m, 0, 0});
  }
  cd.dw_colour_start.assign(pv.ncolour + 1, (int)cd.opsDW.size());
  if (pv.ncolour) cd.dw_colour_start[0] = 0;
  cd.index_bytes = (int64_t)sizeof(int32_t) * (2 * cd.nsep_total + 2 * list_entries + (int64_t)cd.sepdofs.size() +
                                              (int64_t)cd.cptr.size() + (int64_t)cd.cg1.size() +
                                              (int64_t)cd.zptr.size() + (int64_t)cd.zsrc.size()) +
                   (int64_t)sizeof(TileOp) * (int64_t)(cd.opsV.size() + cd.opsS.size() + cd.opsDW.size());
}

// block_of_dof[k] for every entry of the patch dof list: < 0 separator, otherwise a block label
// (any non-negative integer, local to the patch).  Blocks must be pairwise decoupled in the BSR
// pattern; this is checked here, so a wrong hint is an error, never a wrong answer.
inline void build_condensed_host(const PatchView& pv, const int32_t* block_of_dof, CondensedHost& cd,
                                 bool allow_shared = true, bool split_wide = false,
                                 int split_cols = ALFIB_SPLIT_COLS) {
  cd = CondensedHost();
  const int npatch = pv.npatch, bs = pv.bs;
  cd.sepoff.assign(npatch + 1, 0);
  cd.ssoff.assign(npatch + 1, 0);
  cd.blk_start.assign(npatch + 1, 0);
  cd.bl_off.assign(1, 0);
  cd.nb_off.assign(1, 0);
  std::vector<int32_t> mark(pv.ndofs, 0);             // 0: not in patch; k+1: block k; -(s+1): separator position s
  std::vector<int32_t> nbstamp;
  std::vector<std::vector<int32_t>> blk_local, blk_nb;

  for (int p = 0; p < npatch; ++p) {
    const int64_t o = pv.off[p];
    const int n = (int)(pv.off[p + 1] - o);
    const int32_t* I = pv.dofs + o;
    const int32_t* B = block_of_dof + o;
    blk_local.clear();
    std::vector<std::pair<int32_t, int32_t>> labels;   // (label, block index), first-appearance order
    int nsep = 0;
    for (int l = 0; l < n; ++l) {
      if (B[l] < 0) {
        mark[I[l]] = -(nsep + 1);
        cd.seplocal.push_back(l);
        cd.sepdofs.push_back(I[l]);
        ++nsep;
      } else {
        int kk = -1;
        for (auto& lb : labels)
          if (lb.first == B[l]) { kk = lb.second; break; }
        if (kk < 0) {
          kk = (int)labels.size();
          labels.push_back({B[l], kk});
          blk_local.emplace_back();
        }
        blk_local[kk].push_back(l);
        mark[I[l]] = kk + 1;
      }
    }
    const int nblk = (int)blk_local.size();
    blk_nb.assign(nblk, {});
    nbstamp.assign((size_t)nblk * std::max(nsep, 1), 0);
    auto add_nb = [&](int kk, int s) {
      int32_t& st = nbstamp[(size_t)kk * nsep + s];
      if (!st) { st = 1; blk_nb[kk].push_back(s); }
    };
    // structural check + neighbour sets (rows of block dofs and rows of separator dofs)
    int prev_node = -1, prev_mark = 0;
    std::string problem;
    for (int l = 0; l < n && problem.empty(); ++l) {
      const int g = I[l], node = g / bs, mk = mark[g];
      if (node == prev_node && mk == prev_mark) continue;      // same node, same role: same scan
      prev_node = node;
      prev_mark = mk;
      for (int k = pv.rowptr[node]; k < pv.rowptr[node + 1] && problem.empty(); ++k) {
        const int cn = pv.colidx[k];
        for (int c2 = 0; c2 < bs; ++c2) {
          const int mk2 = mark[cn * bs + c2];
          if (mk2 == 0) continue;
          if (mk > 0) {
            if (mk2 > 0) {
              if (mk2 != mk) {
                problem = "patch " + std::to_string(p) + ": blocks " + std::to_string(mk - 1) + " and " +
                          std::to_string(mk2 - 1) + " are coupled in the operator";
                break;
              }
            } else {
              add_nb(mk - 1, -mk2 - 1);
            }
          } else if (mk2 > 0) {
            add_nb(mk2 - 1, -mk - 1);
          }
        }
      }
    }
    for (int t = 0; t < n; ++t) mark[I[t]] = 0;
    if (!problem.empty()) throw std::runtime_error(problem);
    cd.sepoff[p + 1] = cd.sepoff[p] + nsep;
    cd.maxsep = std::max(cd.maxsep, nsep);
    cd.blk_start[p + 1] = cd.blk_start[p] + nblk;
    for (int kk = 0; kk < nblk; ++kk) {
      std::sort(blk_nb[kk].begin(), blk_nb[kk].end());
      const int b = (int)blk_local[kk].size(), m = (int)blk_nb[kk].size();
      if (b > ALFIB_TILE_ROWS || m > ALFIB_TILE_ROWS)
        throw std::runtime_error("condensation block (or its separator neighbourhood) has more than 64 dofs");
      cd.maxb = std::max(cd.maxb, b);
      cd.maxm = std::max(cd.maxm, m);
      cd.bl_local.insert(cd.bl_local.end(), blk_local[kk].begin(), blk_local[kk].end());
      cd.nb_pos.insert(cd.nb_pos.end(), blk_nb[kk].begin(), blk_nb[kk].end());
      cd.bl_off.push_back((int64_t)cd.bl_local.size());
      cd.nb_off.push_back((int64_t)cd.nb_pos.size());
    }
  }
  cd.nblocks = cd.blk_start[npatch];
  cd.nsep_total = cd.sepoff[npatch];
  if (cd.nsep_total >= INT_MAX / 2 || (int64_t)cd.nb_pos.size() >= INT_MAX / 2)
    throw std::runtime_error("condensed set too large for int32 indices");

  // ---- store layout: X_SS tiles of every patch, then V and [D | -W] of every block ------------------
  int64_t cur = 0;
  for (int p = 0; p < npatch; ++p) {
    const int ns = (int)(cd.sepoff[p + 1] - cd.sepoff[p]);
    cd.ssoff[p] = cur;
    cur += (int64_t)ns * ch_roundup2(ns);
  }
  cd.ssoff[npatch] = cur;
  cd.blocks.resize(cd.nblocks);
  cd.g1off.assign(cd.nblocks + 1, 0);
  // cidx = per patch [rs positions (ns)] [global separator dofs (ns)], then per block
  //        [global block dofs (b)] [~(us position) of the neighbours (m)]
  std::vector<int64_t> rs_list(npatch), sg_list(npatch), blk_list(cd.nblocks);
  for (int p = 0; p < npatch; ++p) {
    const int64_t so = cd.sepoff[p];
    const int ns = (int)(cd.sepoff[p + 1] - so);
    rs_list[p] = (int64_t)cd.cidx.size();
    for (int s = 0; s < ns; ++s) cd.cidx.push_back((int32_t)(so + s));
    sg_list[p] = (int64_t)cd.cidx.size();
    for (int s = 0; s < ns; ++s) cd.cidx.push_back(cd.sepdofs[so + s]);
  }
  std::vector<int32_t> idx;
  for (int p = 0; p < npatch; ++p) {
    const int64_t o = pv.off[p], so = cd.sepoff[p];
    for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q) {
      const int b = (int)(cd.bl_off[q + 1] - cd.bl_off[q]), m = (int)(cd.nb_off[q + 1] - cd.nb_off[q]);
      BlockDesc& d = cd.blocks[q];
      d.b = b;
      d.m = m;
      d.voff = cur;
      cur += (int64_t)ch_roundup2(m) * b;
      d.dwoff = cur;
      cur += (int64_t)ch_roundup2(b) * (b + m);
      d.dscale = 1.0;
      d.dofs = (int64_t)cd.bdofs.size();
      d.keys = d.dofs;
      blk_list[q] = (int64_t)cd.cidx.size();
      for (int i = 0; i < b; ++i) {
        const int32_t g = pv.dofs[o + cd.bl_local[cd.bl_off[q] + i]];
        cd.cidx.push_back(g);
        cd.bdofs.push_back(g);
      }
      for (int j = 0; j < m; ++j) {
        const int32_t s = cd.nb_pos[cd.nb_off[q] + j];
        cd.cidx.push_back(~(int32_t)(so + s));
        cd.bdofs.push_back(cd.sepdofs[so + s]);
      }
      idx.resize(b + m);
      std::iota(idx.begin(), idx.end(), 0);
      const int32_t* gd = cd.bdofs.data() + d.dofs;
      std::sort(idx.begin(), idx.end(), [&](int x, int y) { return gd[x] < gd[y]; });
      for (int i = 0; i < b + m; ++i) {
        cd.bkeys.push_back(gd[idx[i]]);
        cd.bperm.push_back(idx[i]);
      }
      cd.g1off[q + 1] = cd.g1off[q] + m;
    }
  }
  cd.g1_total = cd.g1off[cd.nblocks];
  cd.store_elems = cur;

  // ---- K2: contributions to every separator entry, in block order -----------------------------------
  cd.cptr.assign(cd.nsep_total + 1, 0);
  cd.cg1.assign(cd.g1_total, 0);
  for (int p = 0; p < npatch; ++p)
    for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q)
      for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j) cd.cptr[cd.sepoff[p] + cd.nb_pos[j] + 1]++;
  for (int64_t e = 0; e < cd.nsep_total; ++e) cd.cptr[e + 1] += cd.cptr[e];
  {
    std::vector<int32_t> cursor(cd.cptr.begin(), cd.cptr.end() - 1);
    for (int p = 0; p < npatch; ++p)
      for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q)
        for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j)
          cd.cg1[cursor[cd.sepoff[p] + cd.nb_pos[j]]++] = (int32_t)(cd.g1off[q] + (j - cd.nb_off[q]));
  }

  // ---- op lists: V ops of every block; S and DW ops colour-major in iteration order ------------------
  for (int64_t q = 0; q < cd.nblocks; ++q) {
    const BlockDesc& d = cd.blocks[q];
    if (d.m == 0) continue;
    cd.opsV.push_back(TileOp{d.voff, blk_list[q], -1, cd.g1off[q], d.m, d.b, 0, 0});
  }
  cd.s_colour_start.assign(pv.ncolour + 1, 0);
  cd.dw_colour_start.assign(pv.ncolour + 1, 0);
  for (int col = 0; col < pv.ncolour; ++col) {
    cd.s_colour_start[col] = (int)cd.opsS.size();
    cd.dw_colour_start[col] = (int)cd.opsDW.size();
    for (int32_t p : *pv.order) {
      if (pv.colour[p] != col) continue;
      const int64_t so = cd.sepoff[p];
      const int ns = (int)(cd.sepoff[p + 1] - so);
      for (int row0 = 0; row0 < ns; row0 += ALFIB_TILE_ROWS) {
        const int rows = std::min(ns - row0, ALFIB_TILE_ROWS);
        const int64_t tile = cd.ssoff[p] + (int64_t)row0 * ns;
        if (split_wide && split_cols > 0 && ns > 2 * split_cols) {
          // wide separator (a coarse level held as one patch, literal 3-D macro stars): column chunks, each op
          // adds its partial product to us (atomicAdd) and to y
          for (int c0 = 0; c0 < ns; c0 += split_cols) {
            const int nc = std::min(ns - c0, split_cols);
            cd.opsS.push_back(TileOp{tile + (int64_t)c0 * ch_roundup2(rows), rs_list[p] + c0, sg_list[p] + row0, so + row0,
                                     rows, nc, TILEOP_ACCUM, 0});
          }
          cd.any_accum = true;
        } else {
          cd.opsS.push_back(TileOp{tile, rs_list[p], sg_list[p] + row0, so + row0, rows, ns, 0, 0});
        }
      }
      for (int64_t q = cd.blk_start[p]; q < cd.blk_start[p + 1]; ++q) {
        const BlockDesc& d = cd.blocks[q];
        cd.opsDW.push_back(TileOp{d.dwoff, blk_list[q], blk_list[q], -1, d.b, d.b + d.m, 0, 0});
      }
    }
  }
  if (pv.ncolour) {
    cd.s_colour_start[pv.ncolour] = (int)cd.opsS.size();
    cd.dw_colour_start[pv.ncolour] = (int)cd.opsDW.size();
  }
  cd.index_bytes = (int64_t)sizeof(int32_t) * (int64_t)(cd.cidx.size() + cd.sepdofs.size() + cd.cptr.size() + cd.cg1.size()) +
                   (int64_t)sizeof(TileOp) * (int64_t)(cd.opsV.size() + cd.opsS.size() + cd.opsDW.size());
  if (allow_shared) build_shared_blocks(pv, cd);
}

// ---- Schur-complement setup (ALFIB_SCHUR_SETUP) --------------------------------------------------------
// X_SS = (A^-1)[S,S] is the inverse of the Schur complement  A_SS - sum_k A_Sk A_kk^-1 A_kS.  Formed with
// *solves* (W_k = A_kk^-1 A_kN carried through the pivoted elimination of A_kk, not D_k * A_kN with the explicit
// inverse) it is as accurate as cutting X_SS out of the pivoted inverse of the whole patch (measured in numpy on
// the 1275-dof patches at gamma = 1e4, Re = 5000, cond 3e8: 3.5e-9 against 2.4e-9 relative error; with explicit
// D_k: 2.6e-3) and costs ~200x fewer flops (3-D SV k=3: 24 blocks of 45 + one 195 x 195 inverse instead of one
// 1275 x 1275 inverse).  Lists for the two kernels: the block kernel stores C = A_Nk W_k (m x m per factor block);
// the factor kernel, run on the separators as if they were the patches, gathers A_SS through the sorted tables
// and subtracts each instance's C entries at the positions nb_pos.
struct SchurHost {
  std::vector<int32_t> sepsorted, sepperm;   // per patch, parallel to sepdofs: separator dofs ascending / their positions
  std::vector<int64_t> coff;                 // per factor block (sblocks in shared form, else blocks): offset of its C
  int64_t ctotal = 0;
  std::vector<int64_t> inst_c;               // per instance: offset of the C it reads ...
  std::vector<int32_t> inst_ld;              // ... and its leading dimension (m of that factor block)
  std::vector<int32_t> upos;                 // parallel to nb_pos: row / column of C of that neighbour
};

inline void build_schur_lists(const CondensedHost& cd, int npatch, SchurHost& sh) {
  sh = SchurHost();
  sh.sepsorted.resize(cd.sepdofs.size());
  sh.sepperm.resize(cd.sepdofs.size());
  std::vector<int32_t> idx;
  for (int p = 0; p < npatch; ++p) {
    const int64_t so = cd.sepoff[p];
    const int ns = (int)(cd.sepoff[p + 1] - so);
    idx.resize(ns);
    std::iota(idx.begin(), idx.end(), 0);
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return cd.sepdofs[so + a] < cd.sepdofs[so + b]; });
    for (int i = 0; i < ns; ++i) {
      sh.sepsorted[so + i] = cd.sepdofs[so + idx[i]];
      sh.sepperm[so + i] = idx[i];
    }
  }
  const std::vector<BlockDesc>& fb = cd.shared ? cd.sblocks : cd.blocks;
  sh.coff.assign(fb.size() + 1, 0);
  for (size_t k = 0; k < fb.size(); ++k) sh.coff[k + 1] = sh.coff[k] + (int64_t)fb[k].m * fb[k].m;
  sh.ctotal = sh.coff[fb.size()];
  sh.inst_c.resize(cd.nblocks);
  sh.inst_ld.resize(cd.nblocks);
  sh.upos.resize(cd.nb_pos.size());
  for (int64_t q = 0; q < cd.nblocks; ++q) {
    const int64_t k = cd.shared ? cd.inst_dist[q] : q;
    sh.inst_c[q] = sh.coff[k];
    sh.inst_ld[q] = fb[k].m;
    for (int64_t j = cd.nb_off[q]; j < cd.nb_off[q + 1]; ++j)
      sh.upos[j] = cd.shared ? cd.inst_upos[j] : (int32_t)(j - cd.nb_off[q]);
  }
}

// Dense inverse (row-major n x n) of one patch rebuilt from its condensed factors.  `fetch(off,
// count)` returns `count` doubles of the store starting at element `off` (device download in the
// library, a plain copy in the host checker).
template <class Fetch>
inline void condensed_inverse_host(const CondensedHost& cd, int patch, int n, Fetch&& fetch, double* out) {
  if (n == 0) return;
  const int64_t so = cd.sepoff[patch];
  const int ns = (int)(cd.sepoff[patch + 1] - so);
  std::vector<double> XS((size_t)ns * ns, 0.0);        // X_SS row-major, from its 64-row tiles
  {
    const std::vector<double> t = fetch(cd.ssoff[patch], (int64_t)ns * ch_roundup2(ns));
    for (int row0 = 0; row0 < ns; row0 += ALFIB_TILE_ROWS) {
      const int rows = std::min(ns - row0, ALFIB_TILE_ROWS), rt = ch_roundup2(rows);
      const double* tile = t.data() + (int64_t)row0 * ns;
      for (int cc = 0; cc < ns; ++cc)
        for (int r = 0; r < rows; ++r) XS[(size_t)(row0 + r) * ns + cc] = tile[(size_t)cc * rt + r];
    }
  }
  std::fill(out, out + (size_t)n * n, 0.0);
  const int32_t* sl = cd.seplocal.data() + so;
  for (int r = 0; r < ns; ++r)
    for (int cc = 0; cc < ns; ++cc) out[(size_t)sl[r] * n + sl[cc]] = XS[(size_t)r * ns + cc];
  struct Blk {
    int b, m;
    std::vector<double> V, D, W;
    const int32_t* loc;
    const int32_t* nb;
  };
  std::vector<Blk> blks;
  for (int64_t q = cd.blk_start[patch]; q < cd.blk_start[patch + 1]; ++q) {
    const BlockDesc& d = cd.blocks[q];
    Blk k;
    k.b = d.b;
    k.m = d.m;
    k.loc = cd.bl_local.data() + cd.bl_off[q];
    k.nb = cd.nb_pos.data() + cd.nb_off[q];
    k.V.assign((size_t)d.m * d.b, 0.0);
    k.D.assign((size_t)d.b * d.b, 0.0);
    k.W.assign((size_t)d.b * d.m, 0.0);
    if (cd.shared) {
      // rows / columns of the distinct block's tiles: block dofs in the owner's order (opos), the
      // neighbours at their positions in U_k (upos); the D part is stored times dscale
      const BlockDesc& sd = cd.sblocks[cd.inst_dist[q]];
      const int32_t* opos = cd.inst_opos.data() + cd.bl_off[q];
      const int32_t* upos = cd.inst_upos.data() + cd.nb_off[q];
      const int mr = ch_roundup2(sd.m), br = ch_roundup2(sd.b);
      const std::vector<double> tv = fetch(sd.voff, (int64_t)mr * sd.b);
      const std::vector<double> td = fetch(sd.dwoff, (int64_t)br * (sd.b + sd.m));
      for (int cc = 0; cc < d.b; ++cc)
        for (int r = 0; r < d.m; ++r) k.V[(size_t)r * d.b + cc] = tv[(size_t)opos[cc] * mr + upos[r]];
      for (int cc = 0; cc < d.b; ++cc)
        for (int r = 0; r < d.b; ++r) k.D[(size_t)r * d.b + cc] = td[(size_t)opos[cc] * br + opos[r]] / sd.dscale;
      for (int cc = 0; cc < d.m; ++cc)
        for (int r = 0; r < d.b; ++r) k.W[(size_t)r * d.m + cc] = -td[(size_t)(sd.b + upos[cc]) * br + opos[r]];
    } else {
      const int mr = ch_roundup2(d.m), br = ch_roundup2(d.b);
      const std::vector<double> tv = fetch(d.voff, (int64_t)mr * d.b);
      const std::vector<double> td = fetch(d.dwoff, (int64_t)br * (d.b + d.m));
      for (int cc = 0; cc < d.b; ++cc)
        for (int r = 0; r < d.m; ++r) k.V[(size_t)r * d.b + cc] = tv[(size_t)cc * mr + r];
      for (int cc = 0; cc < d.b; ++cc)
        for (int r = 0; r < d.b; ++r) k.D[(size_t)r * d.b + cc] = td[(size_t)cc * br + r];
      for (int cc = 0; cc < d.m; ++cc)
        for (int r = 0; r < d.b; ++r) k.W[(size_t)r * d.m + cc] = -td[(size_t)(d.b + cc) * br + r];
    }
    blks.push_back(std::move(k));
  }
  // X[S, B_l] = -X_SS[:, N_l] V_l ;  X[B_k, B_l] = delta_kl D_k + W_k X_SS[N_k, N_l] V_l ;  X[B_k, S] = -W_k X_SS[N_k, :]
  for (const Blk& l : blks) {
    std::vector<double> XV((size_t)std::max(ns, 1) * l.b, 0.0);     // X_SS[:, N_l] V_l
    for (int r = 0; r < ns; ++r)
      for (int j = 0; j < l.m; ++j) {
        const double xs = XS[(size_t)r * ns + l.nb[j]];
        for (int cc = 0; cc < l.b; ++cc) XV[(size_t)r * l.b + cc] += xs * l.V[(size_t)j * l.b + cc];
      }
    for (int r = 0; r < ns; ++r)
      for (int cc = 0; cc < l.b; ++cc) out[(size_t)sl[r] * n + l.loc[cc]] = -XV[(size_t)r * l.b + cc];
    for (const Blk& k : blks)
      for (int r = 0; r < k.b; ++r)
        for (int cc = 0; cc < l.b; ++cc) {
          double v = (&k == &l) ? k.D[(size_t)r * k.b + cc] : 0.0;
          for (int j = 0; j < k.m; ++j) v += k.W[(size_t)r * k.m + j] * XV[(size_t)k.nb[j] * l.b + cc];
          out[(size_t)k.loc[r] * n + l.loc[cc]] = v;
        }
  }
  for (const Blk& k : blks)
    for (int r = 0; r < k.b; ++r)
      for (int cc = 0; cc < ns; ++cc) {
        double v = 0.0;
        for (int j = 0; j < k.m; ++j) v += k.W[(size_t)r * k.m + j] * XS[(size_t)k.nb[j] * ns + cc];
        out[(size_t)k.loc[r] * n + sl[cc]] = -v;
      }
}
