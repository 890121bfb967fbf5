// nsg_fanlist.inl — host side of assembly variant 5 (nsg_assemble_fan.cuh): the lane records of the "fan"
// scheme.  Included by nsg.cu inside namespace nsg.
//
// For every row owner (velocity P2 node: kind 0; pressure vertex: kind 1) the (owner, cell) pairs are put into
// consecutive lanes of one warp.  Vertex owners: cells in fan order around the vertex (each cell followed by the
// cell across its "previous" edge), so a lane's partner for the columns of the shared edge is its successor in
// the fan.  Edge owners: the two cells of the edge.  A chunk (CTA) takes consecutive owners - its image is a
// contiguous piece of the CSR - and keeps vertex owners and edge owners in different warps.
// *supported = false (and no lists) when the mesh is not an oriented manifold triangulation with vertex valence
// <= 32; the caller then serves the mesh with variant 4.
static int build_fanlist(nsg_ctx *c, int kind, const int32_t *cd, WorkList *out, bool *supported) {
  *supported = false;
  // pattern built on the device (nsg_set_pattern_from_cells): there is no host copy of the column indices; the column
  // offsets of the lanes are looked up by a kernel after the records are uploaded, and every entry of the pattern is
  // covered by construction (no zero-fill bookkeeping)
  const bool dev_off = c->pattern_on_device;
  const int64_t T = c->n_cells, nu = c->n_own_u, nown = c->n_own;
  const int64_t ng = kind == 0 ? nu / 2 : c->n_own_p;
  const int nk = kind == 0 ? 6 : 3;
  constexpr int NW = NPC6 / 32;
  auto group_of = [&](int64_t cell, int k) -> int64_t {
    const int32_t d = cd[15 * cell + (kind == 0 ? uidx(k) : 3 * k + 2)];
    if (kind == 0) return d < nu ? d / 2 : -1;
    return (d >= nu && d < nown) ? d - nu : -1;
  };
  std::vector<int64_t> gptr(ng + 1, 0);
  for (int64_t cell = 0; cell < T; ++cell)
    for (int k = 0; k < nk; ++k) {
      const int64_t g = group_of(cell, k);
      if (g >= 0) gptr[g + 1]++;
    }
  for (int64_t g = 0; g < ng; ++g) gptr[g + 1] += gptr[g];
  const int64_t npairs = gptr[ng];
  std::vector<int32_t> pcell(npairs);
  std::vector<uint8_t> pk(npairs);
  {
    std::vector<int64_t> pos(gptr.begin(), gptr.end() - 1);
    for (int64_t cell = 0; cell < T; ++cell)  // ascending cell order per owner
      for (int k = 0; k < nk; ++k) {
        const int64_t g = group_of(cell, k);
        if (g >= 0) {
          pcell[pos[g]] = (int32_t)cell;
          pk[pos[g]++] = (uint8_t)k;
        }
      }
  }
  // ---- chunks: consecutive owners packed into NW warps; a vertex owner's fan never straddles a warp, nor do the
  //      two lanes of an edge owner; vertex owners fill the first warps of the chunk, edge owners the rest
  std::vector<ChunkInfo> chunks;
  std::vector<uint8_t> place(ng);  // warp << 5 | first lane, per owner
  {
    int vfill[NW], nvw = 0, new_ = 0, ecur = 0;
    std::vector<std::pair<int64_t, uint8_t>> eown;  // edge owners of the open chunk: (owner, edge warp index << 5 | lane)
    ChunkInfo ci{};
    auto close = [&](int64_t g_end) {
      ci.g1 = (int32_t)g_end;
      ci.max_slots = nvw;
      for (auto &e : eown) place[e.first] = (uint8_t)((((e.second >> 5) + nvw) << 5) | (e.second & 31));
      chunks.push_back(ci);
      eown.clear();
      nvw = new_ = ecur = 0;
    };
    bool open = false;
    for (int64_t g = 0; g < ng; ++g) {
      const int np = (int)(gptr[g + 1] - gptr[g]);
      if (np > 32) return NSG_OK;  // unsupported: caller falls back
      const bool is_edge = kind == 0 && np > 0 && pk[gptr[g]] >= 3;
      if (is_edge && np > 2) return NSG_OK;
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (!open) {
          ci = ChunkInfo{};
          ci.g0 = (int32_t)g;
          open = true;
        }
        bool ok = g - ci.g0 < 255;
        if (ok && np > 0) {
          if (!is_edge) {
            int w = 0;
            while (w < nvw && vfill[w] + np > 32) ++w;
            if (w == nvw) {
              if (nvw + new_ + 1 > NW)
                ok = false;
              else
                vfill[nvw++] = 0;
            }
            if (ok) {
              place[g] = (uint8_t)((w << 5) | vfill[w]);
              vfill[w] += np;
            }
          } else {
            if (new_ == 0 || ecur + np > 32) {
              if (nvw + new_ + 1 > NW)
                ok = false;
              else
                ++new_, ecur = 0;
            }
            if (ok) {
              eown.push_back({g, (uint8_t)(((new_ - 1) << 5) | ecur)});
              ecur += np;
            }
          }
        }
        if (ok) break;
        close(g);
        open = false;
      }
    }
    if (open) close(ng);
  }
  const int64_t nchunks = (int64_t)chunks.size();
  PairRec no_work;
  std::memset(&no_work, 0, sizeof no_work);
  no_work.cell = -1;
  // 32 bytes per lane, gigabytes at bench size: allocated untouched and initialised chunk by chunk inside the parallel loop
  // (a std::vector would fill it on one thread first)
  std::unique_ptr<PairRec[]> recs_mem(new PairRec[(size_t)std::max<int64_t>(nchunks * NPC6, 1)]);
  PairRec *recs = recs_mem.get();
  int bad = 0, unsupported = 0;
  int64_t max_smem = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_smem) reduction(+ : bad, unsupported)
  for (int64_t b = 0; b < nchunks; ++b) {
    ChunkInfo &ci = chunks[b];
    if (kind == 0) {
      ci.rs = c->h_rowptr[2 * (int64_t)ci.g0];
      ci.cnt = (int32_t)(c->h_rowptr[2 * (int64_t)ci.g1] - ci.rs);
    } else {
      ci.rs = c->h_rowptr[nu + ci.g0], ci.ms = c->h_pm_rowptr[nu + ci.g0];
      ci.cnt = (int32_t)(c->h_rowptr[nu + ci.g1] - ci.rs), ci.mcnt = (int32_t)(c->h_pm_rowptr[nu + ci.g1] - ci.ms);
    }
    ci.rec_base = b * NPC6;
    for (int l = 0; l < NPC6; ++l) recs[(size_t)(b * NPC6 + l)] = no_work;
    const int64_t img = kind == 0 ? (int64_t)ci.cnt : (int64_t)ci.cnt + ci.mcnt;
    std::vector<uint8_t> touched(dev_off ? (size_t)0 : (size_t)img, 0);
    int64_t n_touched = 0;
    int owners_with_cells = 0;
    for (int64_t g = ci.g0; g < ci.g1; ++g) {
      const int s = (int)(gptr[g + 1] - gptr[g]);
      if (s == 0) continue;
      ++owners_with_cells;
      const int64_t p0 = gptr[g];
      const bool is_edge = kind == 0 && pk[p0] >= 3;
      const int warp = place[g] >> 5, lane0 = place[g] & 31;
      // rotated vertex ids of every pair: v0 (owner's vertex / first vertex of the owner's edge), v1 "next", v2 "previous"
      int32_t v0[32], v1[32], v2[32];
      int rot[32];
      for (int i = 0; i < s; ++i) {
        const int k = pk[p0 + i];
        const int r = k >= 3 ? k - 3 : k;
        rot[i] = r;
        const int32_t *cdc = cd + 15 * (int64_t)pcell[p0 + i];
        v0[i] = cdc[3 * r], v1[i] = cdc[3 * ((r + 1) % 3)], v2[i] = cdc[3 * ((r + 2) % 3)];
      }
      int order[32], succ[32], pred[32];  // succ/pred: index of the pair, -1 none
      for (int i = 0; i < s; ++i) succ[i] = pred[i] = -1;
      bool ok = true;
      if (!is_edge) {
        for (int i = 0; i < s && ok; ++i)
          for (int j = 0; j < s; ++j) {
            if (j == i) continue;
            if (v1[j] == v1[i] || v2[j] == v2[i]) ok = false;  // an edge run twice in the same direction / by 3 cells
            if (v1[j] == v2[i]) succ[i] = j, pred[j] = i;
          }
        int n_ord = 0;
        bool seen[32] = {false};
        for (int pass = 0; pass < 2 && ok; ++pass)
          for (int i = 0; i < s; ++i) {
            if (seen[i] || (pass == 0 && pred[i] >= 0)) continue;  // open fans first, from their first cell; then closed ones
            for (int j = i; j >= 0 && !seen[j]; j = succ[j]) seen[j] = true, order[n_ord++] = j;
          }
        if (n_ord != s) ok = false;
      } else {
        order[0] = 0;
        if (s == 2) {
          order[1] = 1;
          if (v0[0] != v1[1] || v1[0] != v0[1])
            ok = false;
          else
            succ[0] = 1, succ[1] = 0, pred[0] = 1, pred[1] = 0;
        } else if (s > 2)
          ok = false;
      }
      if (!ok) {
        unsupported++;
        continue;
      }
      int posn[32];
      for (int i = 0; i < s; ++i) posn[order[i]] = i;
      const int64_t row = kind == 0 ? 2 * g : nu + g;
      const int64_t rs = c->h_rowptr[row], re = c->h_rowptr[row + 1];
      const int64_t len = re - rs, roff = rs - ci.rs;
      if (len >= 65535 || roff >= 65535) bad++;
      if (kind == 0 && c->h_rowptr[row + 2] - re != len) bad++;
      const int32_t *cb = dev_off ? nullptr : c->h_col.data() + rs, *ce = dev_off ? nullptr : c->h_col.data() + re;
      const int32_t *mb = cb, *me = ce;
      int64_t moff = roff, mrow_off = 0;
      if (kind == 1) {
        if (!dev_off) {
          mb = c->h_pm_col.data() + c->h_pm_rowptr[row];
          me = c->h_pm_col.data() + c->h_pm_rowptr[row + 1];
        }
        mrow_off = c->h_pm_rowptr[row] - ci.ms;
        moff = ci.cnt + mrow_off;
        if (mrow_off >= 65535) bad++;
      }
      for (int i = 0; i < s; ++i) {
        const int lane = lane0 + posn[i];
        PairRec rcd;
        std::memset(&rcd, 0, sizeof rcd);
        rcd.cell = pcell[p0 + i];
        const int32_t *cdc = cd + 15 * (int64_t)rcd.cell;
        const int r = rot[i];
        const bool has_partner = succ[i] >= 0;
        const bool head = posn[i] == 0;
        const bool write_next = pred[i] < 0;
        const int partner = has_partner ? lane0 + posn[succ[i]] : lane;
        // which column groups this lane stores (rotated local index): vertex owners 2,5,4 (+1,3 if write_next) (+0 if head);
        // edge owners 0,2,4,5 (+1 if write_next) (+3 if head); pressure columns likewise
        bool wcol[6], wp[3];
        if (!is_edge) {
          wcol[0] = head, wcol[1] = write_next, wcol[2] = true, wcol[3] = write_next, wcol[4] = true, wcol[5] = true;
          wp[0] = head, wp[1] = write_next, wp[2] = true;
        } else {
          wcol[0] = true, wcol[1] = write_next, wcol[2] = true, wcol[3] = head, wcol[4] = true, wcol[5] = true;
          wp[0] = true, wp[1] = write_next, wp[2] = true;
        }
        for (int l = 0; l < 6 && !dev_off; ++l) {
          const int lc = l < 3 ? (l + r) % 3 : 3 + (l - 3 + r) % 3;  // canonical local node of rotated node l
          const int32_t tgt = cdc[uidx(lc)];
          const int32_t *p = std::lower_bound(cb, ce, tgt);
          if (p == ce || *p != tgt || p + 1 == ce || p[1] != tgt + 1) {
            bad++;
            continue;
          }
          rcd.off[l] = (uint16_t)(p - cb);
          if (wcol[l]) {
            const int64_t e0 = roff + (p - cb);
            const int nrow = kind == 0 ? 2 : 1;
            for (int rr = 0; rr < nrow; ++rr)
              for (int cc = 0; cc < 2; ++cc) {
                uint8_t &f = touched[(size_t)(e0 + rr * len + cc)];
                if (f) bad++;
                f = 1, ++n_touched;
              }
          }
        }
        for (int m = 0; m < 3 && !dev_off; ++m) {
          const int32_t tgt = cdc[3 * ((m + r) % 3) + 2];
          const int32_t *p = std::lower_bound(mb, me, tgt);
          if (p == me || *p != tgt) {
            bad++;
            continue;
          }
          rcd.off[6 + m] = (uint16_t)(p - mb);
          if (wp[m]) {
            const int nrow = kind == 0 ? 2 : 1;
            for (int rr = 0; rr < nrow; ++rr) {
              uint8_t &f = touched[(size_t)(moff + (p - mb) + rr * len)];
              if (f) bad++;
              f = 1, ++n_touched;
            }
          }
        }
        const int pred_lane = pred[i] >= 0 ? lane0 + posn[pred[i]] : lane;
        rcd.k = (int32_t)((uint32_t)pk[p0 + i] | ((uint32_t)partner << 3) | ((uint32_t)has_partner << 8) | ((uint32_t)head << 9) |
                          ((uint32_t)write_next << 10) | ((uint32_t)(s - 1 - posn[i]) << 11) | ((uint32_t)(g - ci.g0) << 16) |
                          ((uint32_t)pred_lane << 24));
        rcd.off[9] = (uint16_t)roff;
        rcd.off[10] = kind == 0 ? (uint16_t)len : (uint16_t)mrow_off;
        recs[(size_t)(ci.rec_base + warp * 32 + lane)] = rcd;
      }
    }
    // an entry of the pattern no cell contributes to (a pattern wider than the mesh implies) must still be written: zero-fill
    // (pressure chunks are always zero-filled: the p-p block of the Jacobian is structurally present and never written)
    if (kind == 0) ci.pad = (dev_off || (n_touched == img && owners_with_cells == ci.g1 - ci.g0)) ? 0 : 1;
    // the first 16 bytes of the header carry what the pipelined kernel needs at the end of an iteration: g0, g1, image
    // entries, zero-fill flag
    ci.n_threads = ci.cnt, ci.max_slots = (int32_t)ci.pad;
    const int64_t smem = kind == 0 ? (int64_t)ci.cnt + 2 * (ci.g1 - ci.g0) + 2 : (int64_t)ci.cnt + ci.mcnt + 2;
    max_smem = std::max(max_smem, smem);
  }
  if (bad) return fail(NSG_ERR_ARG, "cell_dofs do not match the sparsity pattern (or a row has >= 65535 entries)");
  if (unsupported) return NSG_OK;
  out->n_groups = ng;
  out->n_chunks = nchunks;
  out->n_pairs = npairs;
  out->n_recs = nchunks * NPC6;
  out->max_stage = max_smem;
  NSG_TRY(upload(c, &out->chunks, chunks.data(), nchunks));
  NSG_TRY(upload(c, &out->recs, recs, nchunks * NPC6));
  if (dev_off && nchunks > 0) {
    int32_t *err = nullptr, h_err = 0;
    NSG_TRY(dev_alloc(&err, 1));
    NSG_CUDA(cudaMemsetAsync(err, 0, 4, c->stream));
    k_fan_offsets<<<grid_for(nchunks * NPC6, 256, 1 << 30), 256, 0, c->stream>>>(kind, nchunks * NPC6, NPC6, out->recs, out->chunks, c->cell_dofs, nu,
                                                                                c->rowptr, c->col, c->pm_rowptr, c->pm_col, err);
    c->launches++;
    cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    dev_free(err);
    if (h_err) return fail(NSG_ERR_ARG, "cell_dofs do not match the sparsity pattern");
  }
  NSG_CUDA(cudaStreamSynchronize(c->stream));
  *supported = true;
  return NSG_OK;
}
