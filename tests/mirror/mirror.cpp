// mirror.cpp -- TEST INFRASTRUCTURE.  Strings the __host__ __device__ format logic of
// phyngsc_b200/csrc/phy_core.cuh together serially on the CPU so that tokeniser, classification, Huffman,
// header layout, walkers and bit sinks can be checked against the oracle without a GPU.  It is not a
// fallback: nothing in the product links or loads it.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../phyngsc_b200/csrc/phy_core.cuh"

using namespace phy;

static void amax(u32 &a, u32 v) { if (v > a) a = v; }

extern "C" int mirror_compress_window(const u8 *b, u64 readable, i64 rsize, u32 rec_start, i32 overlap, u32 cap,
                                      u8 *out, u32 out_cap, u32 *sec_len, u32 *n_records, u64 *bytes_consumed) {
  // record split, serial (same rule as the plan kernel: phyNGSC.cpp:254-331)
  std::vector<u32> te, se, rstart;
  {
    u64 pos = rec_start;
    rstart.push_back((u32)pos);
    i64 lim = rsize < (i64)readable ? rsize : (i64)readable;
    for (;;) {
      const u8 *q = (const u8 *)memchr(b + pos, '\n', readable - pos);
      if (!q) break;
      u64 t = q - b;
      if (!te.empty() && (i64)t >= lim) break;
      q = (const u8 *)memchr(b + t + 1, '\n', readable - t - 1);
      if (!q) return E_MALFORMED;
      u64 s = q - b;
      u64 nx = 2 * s - t + 3;
      if (nx > readable + 1) return E_MALFORMED;
      te.push_back((u32)t); se.push_back((u32)s); rstart.push_back((u32)nx);
      pos = nx;
      if (te.size() > 1) {
        if ((i64)nx >= rsize - overlap) break;
        if (te.size() > cap) break;
      }
      if (pos >= readable) break;
    }
  }
  u32 R = (u32)te.size();
  if (!R) return E_MALFORMED;
  u8 lut[256];
  fill_char_lut(lut);
  std::vector<u16> kx(R);
  SbAcc *A = (SbAcc *)calloc(1, sizeof(SbAcc));
  // stat1
  TitleCursor c0; c0.init(b, rstart[0], te[0], lut);
  Tok t; u32 nf = 0; u32 off0[MAXF], len0[MAXF];
  while (c0.next(t)) { if (nf < (u32)MAXF) { off0[nf] = t.start - rstart[0]; len0[nf] = t.end - t.start; } ++nf; }
  if (nf == 0 || nf > (u32)MAXF) return E_UNSUPPORTED;
  std::vector<u32> vals((size_t)R * nf);
  for (u32 r = 0; r < R; ++r) {
    u32 L = se[r] - te[r] - 1, qs = se[r] + 3;
    if (L == 0 || b[se[r] + 1] != '+' || b[se[r] + 2] != '\n') return E_MALFORMED;
    SeqStat st;
    seqqual_stat(b, te[r] + 1, L, qs, st, [&](u8 c) { A->dna_occ[c]++; }, [&](u8 q) { A->qpresent[q >> 5] |= 1u << (q & 31); });
    if (st.err) return E_UNSUPPORTED;
    A->dna_occ['A'] += st.acgt[0]; A->dna_occ['C'] += st.acgt[1]; A->dna_occ['G'] += st.acgt[2]; A->dna_occ['T'] += st.acgt[3];
    kx[r] = (u16)(st.kept | (st.xfer << 15));
    amax(A->max_qlen, L); amax(A->max_slen, st.kept); amax(A->inv_min_qlen, ~L);
    if (count_seps(b, rstart[r], te[r]) != nf) return E_FIELDS;
    TitleCursor cur; cur.init(b, rstart[r], te[r], lut);
    for (u32 f = 0; f < nf; ++f) {
      cur.next(t);
      u32 len = t.end - t.start;
      FieldAcc &a = A->f[f];
      amax(a.inv_min_len, ~len); amax(a.max_len, len);
      if (!t.num) a.not_num = 1;
      amax(a.kmax_v, key_of((i32)t.v)); amax(a.kinvmin_v, ~key_of((i32)t.v));
      if (r >= 1) { u32 kd = key_of((i32)(t.v - vals[(size_t)(r - 1) * nf + f])); amax(a.kmax_d, kd); amax(a.kinvmin_d, ~kd); }
      vals[(size_t)r * nf + f] = t.v;
      u32 m = len < len0[f] ? len : len0[f];
      for (u32 p = 0; p < m && p < (u32)MAXLEN0; ++p) if (b[t.start + p] != b[rstart[0] + off0[f] + p]) a.mism[p >> 5] |= 1u << (p & 31);
    }
  }
  const u32 AW = (8u << 20) / 4;
  u32 *arena = (u32 *)calloc(AW, 4);
  SbClass *C = (SbClass *)calloc(1, sizeof(SbClass));
  classify_subblock(b, lut, *A, R, rstart[0], te[0], arena, AW, *C);
  int rc = C->status;
  std::vector<u32> qoff(R + 1), doff(R + 1), blkoff((R + 31) / 32 + 1);
  if (!rc) {
    TableDesc *td = (TableDesc *)(arena + C->tabdesc_off);
    for (u32 i = C->zero_begin; i < C->zero_end; ++i) arena[i] = 0;
    if (!C->plain) for (u32 i = 0; i < C->nsym; ++i) arena[C->dnastat_off + i] = A->dna_occ[C->symbols[i]];
    // qhist
    u32 *gq = arena + C->qstat_off;
    for (u32 r = 0; r < R; ++r) {
      u32 L = se[r] - te[r] - 1, qs = se[r] + 3;
      for (u32 p = 0; p < L; ++p) {
        u8 q = b[qs + p];
        if (kx[r] >> 15) { u32 a = amb_code(b[te[r] + 1 + p]); if (a > 1) q = xfer_qual(a, q); }
        gq[(p + 1) * C->nq + C->qua_code[q]]++; gq[C->qua_code[q]]++;
      }
    }
    // stat2: histograms + block flags
    for (u32 lo = 0; lo < R; lo += 32) {
      u32 hi = lo + 32 < R ? lo + 32 : R, flags = 0;
      for (u32 f = 0; f < nf; ++f) {
        const FieldClass &F = C->f[f];
        if (F.kind == K_CONST) continue;
        bool bit = true;
        i32 bd = 0;
        u32 st_lo = 0, len_lo = 0;
        for (u32 r = lo; r < hi; ++r) {
          TitleCursor cur; cur.init(b, rstart[r], te[r], lut);
          for (u32 k = 0; k <= f; ++k) cur.next(t);
          u32 len = t.end - t.start;
          i32 v = (i32)t.v, pv = r ? (i32)vals[(size_t)(r - 1) * nf + f] : 0;
          if (F.kind == K_STR) {
            if (r == lo) { st_lo = t.start; len_lo = len; }
            else if (len != len_lo || memcmp(b + t.start, b + st_lo, len)) bit = false;
            const u16 *sm = (const u16 *)(arena + F.slotmap_off);
            for (u32 j = 0; j < len; ++j)
              if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) arena[td[sm[j < 128 ? j : 128]].freq_off + b[t.start + j]]++;
          } else if (F.is_delta) {
            i32 dl = wsub(v, pv);
            if (r == lo + 1) bd = dl; else if (r > lo + 1 && dl != bd) bit = false;
            if (F.has_table && r >= 1) arena[td[F.tab].freq_off + (u32)wsub(dl, F.base)]++;
          } else {
            if (v != (i32)vals[(size_t)lo * nf + f]) bit = false;
            if (F.has_table) arena[td[F.tab].freq_off + (u32)wsub(v, F.base)] += r == 0 ? 2 : 1;
          }
        }
        if (F.kind == K_NUM && F.is_delta && bd != F.min_d) bit = false;
        if (bit) flags |= 1u << f;
      }
      arena[C->flagbits_off + lo / 32] = flags;
    }
    // huffman
    HuffScratch *HS = new HuffScratch;
    for (u32 i = 0; i < C->ntab; ++i)
      td[i].tree_len = huff_table(arena + td[i].freq_off, td[i].n, (u64 *)(arena + td[i].cl_off), (u8 *)(arena + td[i].tree_off), *HS, 0u, 1u, NoSync());
    delete HS;
    // lengths
    auto prev_of = [&](u32 r) { return [&, r](u32 f, i32) { return r ? (i32)vals[(size_t)(r - 1) * nf + f] : 0; }; };
    for (u32 r = 0; r < R; ++r) {
      u32 L = se[r] - te[r] - 1; bool xf = kx[r] >> 15;
      CountSink q; q.init();
      QFull qt; qt.cl = (const u64 *)(arena + td[C->tq0].cl_off); qt.nq = C->nq;
      quality_record(b, te[r] + 1, L, se[r] + 3, xf, C->qua_code, qt, q);
      qoff[r] = (u32)q.bits;
      CountSink dn; dn.init();
      dna_record(b, te[r] + 1, L, xf, C->plain != 0, C->sym_code, C->plain ? (const u64 *)0 : (const u64 *)(arena + td[C->tdna].cl_off), dn);
      doff[r] = (u32)dn.bits;
    }
    for (u32 lo = 0; lo < R && C->nnc; lo += 32) {
      u32 hi = lo + 32 < R ? lo + 32 : R; u64 bits = C->nnc;
      for (u32 r = lo; r < hi; ++r) {
        CountSink ts; ts.init();
        title_record(b, lut, rstart[r], te[r], *C, C->f, C->ncf, C->ncskip, arena, arena[C->flagbits_off + lo / 32], r == lo, prev_of(r), ts);
        bits += ts.bits;
      }
      blkoff[lo / 32] = (u32)((bits + 7) / 8);
    }
    // layout
    std::vector<u32> tlen(C->ntab), tdst(C->ntab);
    for (u32 i = 0; i < C->ntab; ++i) tlen[i] = td[i].tree_len;
    if (!layout_headers(b, *C, arena, tlen.data(), tdst.data())) rc = E_UNSUPPORTED;
    for (u32 i = 0; i < C->ntab; ++i) td[i].dst = tdst[i];
    if (!rc) {
      u8 *stage = (u8 *)(arena + C->stage_off);
      for (u32 i = 0; i < C->ntab; ++i) memcpy(stage + td[i].dst, (u8 *)(arena + td[i].tree_off), td[i].tree_len);
      u64 qb = 0, db = 0, tb = 0;
      for (u32 r = 0; r < R; ++r) { u32 v = qoff[r]; qoff[r] = (u32)qb; qb += v; v = doff[r]; doff[r] = (u32)db; db += v; }
      for (u32 k = 0; k < C->nblk && C->nnc; ++k) { u32 v = blkoff[k]; blkoff[k] = (u32)tb; tb += v; }
      finish_layout(*C, (u32)tb, qb, db);
      if (C->payload_len > out_cap) rc = E_CAPACITY;
    }
    // the 16-bit packed copy of the quality tables the GPU keeps in shared memory
    std::vector<u16> pk((size_t)(C->max_qlen + 1) * C->nq);
    bool use_packed = true;
    for (size_t i = 0; i < pk.size(); ++i) use_packed = qpack_entry(((const u64 *)(arena + td[C->tq0].cl_off))[i], pk[i]) && use_packed;
    if (!rc) {
      memset(out, 0, (C->payload_len + 7) & ~3u);
      u32 *outw = (u32 *)out;
      ByteWriter w; w.p = out; w.n = 0;
      w.word(C->R); w.word(C->max_qlen); w.word(C->max_slen); w.byte((u8)C->nsym); w.byte(0); w.byte((u8)C->nq); w.word(C->flags);
      const u32 o_title = C->info_len, o_qual = o_title + C->title_len, o_dna = o_qual + C->qual_len;
      const u8 *stage = (const u8 *)(arena + C->stage_off);
      memcpy(out + o_title, stage, C->thdr_len);
      memcpy(out + o_qual, stage + C->thdr_cap, C->qhdr_len);
      memcpy(out + o_dna, stage + C->thdr_cap + C->qhdr_cap, C->dhdr_len);
      for (u32 r = 0; r < R; ++r) {
        u32 L = se[r] - te[r] - 1; bool xf = kx[r] >> 15;
        OrSink k; k.init(outw, (u64)INFO_FIXED * 8 + (u64)r * C->nb_len); k.put(L, C->nb_len); k.finish();
        OrSink q; q.init(outw, (u64)(o_qual + C->qhdr_len) * 8 + qoff[r]);
        if (use_packed) { QPacked qp; qp.pk = pk.data(); qp.nq = C->nq; quality_record(b, te[r] + 1, L, se[r] + 3, xf, C->qua_code, qp, q); }
        else { QFull qt; qt.cl = (const u64 *)(arena + td[C->tq0].cl_off); qt.nq = C->nq; quality_record(b, te[r] + 1, L, se[r] + 3, xf, C->qua_code, qt, q); }
        q.finish();
        OrSink dn; dn.init(outw, (u64)(o_dna + C->dhdr_len) * 8 + doff[r]);
        dna_record(b, te[r] + 1, L, xf, C->plain != 0, C->sym_code, C->plain ? (const u64 *)0 : (const u64 *)(arena + td[C->tdna].cl_off), dn); dn.finish();
      }
      for (u32 lo = 0; lo < R && C->nnc; lo += 32) {
        u32 hi = lo + 32 < R ? lo + 32 : R, flags = arena[C->flagbits_off + lo / 32];
        OrSink ts; ts.init(outw, (u64)(o_title + C->thdr_len + blkoff[lo / 32]) * 8);
        u32 v = 0;
        for (u32 f = 0; f < nf; ++f) if (C->f[f].kind != K_CONST) v = (v << 1) | ((flags >> f) & 1u);
        ts.put(v, C->nnc);
        for (u32 r = lo; r < hi; ++r) title_record(b, lut, rstart[r], te[r], *C, C->f, C->ncf, C->ncskip, arena, flags, r == lo, prev_of(r), ts);
        ts.finish();
      }
      sec_len[0] = C->info_len; sec_len[1] = C->title_len; sec_len[2] = C->qual_len; sec_len[3] = C->dna_len;
      *n_records = R; *bytes_consumed = rstart[R];
    }
  }
  free(A); free(arena); free(C);
  return rc;
}

extern "C" u32 mirror_huffman(const u32 *freq, u32 n, u64 *cl, u8 *tree) {
  HuffScratch *HS = new HuffScratch;
  u32 k = huff_table(freq, n, cl, tree, *HS, 0u, 1u, NoSync());
  delete HS;
  return k;
}
