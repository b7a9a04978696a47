/*
 * phy_title.cuh -- the title kernels that run on PARSED titles (included by phy_b200.cu after phy_encode.cuh).
 *
 * k_stat1 (phy_kernels.cuh) is the only kernel that tokenises title lines.  It leaves, per 128-record chunk, the mask of
 * the fields in which some record of the chunk differs from record 0 of the subblock, and for those fields one row each of
 * numeric values (Dev::tv) and token positions (Dev::tp).  The two kernels here read those rows instead of the title bytes:
 *
 *   k_stat2      numeric / per-position char histograms and the 32-record block flags (tasks.cpp:64-93, 127-182)
 *   k_enc_title  info length bits and title tokens of the single-walk encoder (phyNGSC.cpp:732-742, tasks.cpp:393-509)
 *
 * Numeric fields never touch the input again; string fields fetch exactly their characters from the input (a token that is
 * record 0's is read from record 0's title line).  Both kernels are warp-autonomous (warp = 8 consecutive 32-record blocks,
 * lane = record) and stage nothing in shared memory, so many more warps fit an SM than with staged title lines.
 */
#pragma once
#include "phy_encode.cuh"

namespace phy {

/* where the rows of one chunk start, and which of its fields have rows */
struct ChunkRows {
  size_t row0;     /* entry index of (chunk, field 0, record 0) in tv / tp */
  size_t cf;       /* entry index of (chunk, field 0) in chunk_first / chunk_last */
  u32 mask;
};
__device__ __forceinline__ ChunkRows chunk_rows(const Dev &d, const SbPlan &P, u32 chunk) {
  ChunkRows c;
  const size_t ch = (size_t)P.chunk_base + chunk;
  c.row0 = ch * d.nfs * CH; c.cf = ch * MAXF; c.mask = d.chunk_mask[ch];
  return c;
}
/* numeric value of field f of record i of the chunk (utils::to_num of its token) */
__device__ __forceinline__ u32 parsed_value(const Dev &d, const ChunkRows &c, u32 f, u32 i) {
  return ((c.mask >> f) & 1u) ? d.tv[c.row0 + (size_t)f * CH + i] : d.chunk_first[c.cf + f]; /* untouched: every record has record 0's value */
}
/* token of field f of record i: offset of its first character in the batch, and its length */
struct TokRef { u32 off, len; bool same0; };
__device__ __forceinline__ TokRef parsed_token(const Dev &d, const ChunkRows &c, const SbClass &C, const FieldClass &F, u32 f, u32 i, u32 ts) {
  TokRef t; t.off = C.ts0 + F.off0; t.len = F.len0; t.same0 = true;
  if ((c.mask >> f) & 1u) {
    const u32 e = d.tp[c.row0 + (size_t)f * CH + i];
    if ((e & 0xFFFFu) != TP_SAME) { t.off = ts + (e & 0xFFFFu); t.len = e >> 16; t.same0 = false; }
  }
  return t;
}

/* ---- stat2 ---------------------------------------------------------------------------------------------------------- */
constexpr u32 S2W = 8;  /* warps per CTA */
constexpr u32 S2B = 8;  /* 32-record blocks per warp */

__global__ void __launch_bounds__(S2W * 32) k_stat2(Dev d) {
  __shared__ TitleTabs T;
  __shared__ u32 chist[CSLOTS * 256];
  const u32 s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const SbClass &C = d.cls[s];
  if (C.status || C.nnc == 0) return; /* every field constant: no histogram, no block flag is ever read */
  if (blockIdx.x * (S2W * S2B) >= C.nblk) return;
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const u32 nnc = C.nnc, R = C.R;
  /* char histograms of the first CSLOTS per-position tables are privatised in shared memory */
  const u32 ncs = min(C.ntab - C.tchr0, (u32)CSLOTS);
  for (u32 i = tid; i < ncs * 256; i += S2W * 32) chist[i] = 0;
  load_title_tabs(C, T);
  __syncthreads();
  const u32 g0 = min((blockIdx.x * S2W + w) * S2B, C.nblk), g1 = min(g0 + S2B, C.nblk);
  for (u32 g = g0; g < g1; ++g) {
    const u32 nrec = min(32u, R - g * 32), rec = g * 32 + lane, i = rec & (CH - 1);
    const bool on = lane < nrec;
    const ChunkRows cr = chunk_rows(d, P, g / (CH / 32));
    const u32 ts = d.rstart[P.first_rec + min(rec, R - 1)];
    u32 flags = 0;
    for (u32 k = 0; k < nnc; ++k) {
      const u32 f = T.ncf[k];
      const FieldClass &F = T.fc[f];
      if (F.kind == K_NUM) {
        const i32 v = on ? (i32)parsed_value(d, cr, f, i) : 0;
        i32 pv = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if (lane == 0 && rec > 0) pv = (i32)(i == 0 ? d.chunk_last[cr.cf - MAXF + f] : parsed_value(d, cr, f, i - 1)); /* last record of the block before */
        const i32 dl = wsub(v, pv);
        const bool hasd = on && rec > 0;
        bool pred;
        if (F.is_delta) {
          /* tasks.cpp:127-147 and :415: delta of the block's 2nd record, all later deltas equal to it, and equal to min_delta */
          i32 bd = __shfl_sync(0xFFFFFFFFu, dl, 1);
          if (nrec < 2) bd = 0;
          pred = !on || lane < 2 || dl == bd;
          pred = __all_sync(0xFFFFFFFFu, pred) && bd == F.min_d;
          if (F.has_table) warp_hist_add(arena + F.freq_off, (u32)wsub(dl, F.base), hasd);
        } else {
          const i32 v_lo = __shfl_sync(0xFFFFFFFFu, v, 0);
          pred = __all_sync(0xFFFFFFFFu, !on || v == v_lo);
          if (F.has_table) {
            warp_hist_add(arena + F.freq_off, (u32)wsub(v, F.base), on);
            if (on && rec == 0) atomicAdd(arena + F.freq_off + (u32)wsub(v, F.base), 1u); /* seed, phyNGSC.cpp:368 */
          }
        }
        if (pred) flags |= 1u << f;
        continue;
      }
      /* string field: block flag = every token of the block equals the block's first (tasks.cpp:64-81); per-position char
       * histogram of the positions that are not constant over the subblock (tasks.cpp:83-93) */
      const TokRef t = parsed_token(d, cr, C, F, f, i, ts);
      const u16 *sm = (const u16 *)(arena + F.slotmap_off);
      if (__all_sync(0xFFFFFFFFu, !on || t.same0)) {
        /* every record of the block carries record 0's token: one addition per counted position for the whole block */
        const u8 *d0 = d.in + C.ts0 + F.off0;
        for (u32 j = lane; j < F.len0; j += 32)
          if ((F.mism[j >> 5] >> (j & 31)) & 1u) {
            const u32 loc = (u32)sm[j < 128 ? j : 128] - C.tchr0, ch = d0[j];
            if (loc < ncs) atomicAdd(&chist[loc * 256 + ch], nrec); else atomicAdd(arena + C.chr_freq_off + loc * 256 + ch, nrec);
          }
        flags |= 1u << f;
        continue;
      }
      const u32 len_lo = __shfl_sync(0xFFFFFFFFu, t.len, 0), off_lo = __shfl_sync(0xFFFFFFFFu, t.off, 0);
      bool pred = !on || t.len == len_lo;
      const u32 maxlen = __reduce_max_sync(0xFFFFFFFFu, on ? t.len : 0u);
      for (u32 j = 0; j < maxlen; ++j) {
        const bool in_tok = on && j < t.len;
        const u32 ch = in_tok ? d.in[t.off + j] : 0u;
        if (in_tok && pred && ch != d.in[off_lo + j]) pred = false;
        const bool need = in_tok && (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u));
        if (!__any_sync(0xFFFFFFFFu, need)) continue;
        /* lanes of the warp that are at the same character of the same table add once, together */
        const u32 loc = (u32)sm[j < 128 ? j : 128] - C.tchr0;
        const u32 grp = __match_any_sync(0xFFFFFFFFu, need ? ch : 0xFFFFFFFFu);
        if (need && (u32)(__ffs(grp) - 1) == lane) {
          if (loc < ncs) atomicAdd(&chist[loc * 256 + ch], (u32)__popc(grp));
          else atomicAdd(arena + C.chr_freq_off + loc * 256 + ch, (u32)__popc(grp));
        }
      }
      if (__all_sync(0xFFFFFFFFu, pred)) flags |= 1u << f;
    }
    if (lane == 0) arena[C.flagbits_off + g] = flags;
  }
  __syncthreads();
  for (u32 i = tid; i < ncs * 256; i += S2W * 32) {
    const u32 v = chist[i];
    if (v) atomicAdd(arena + C.chr_freq_off + i, v);
  }
}

/* ---- title + info of the single-walk encoder ------------------------------------------------------------------------- */
/* Title tokens of one record from the parsed rows (the walk of title_record, phy_core.cuh, without a tokeniser):
 * lane = record of the block, all 32 lanes walk together (the previous record's numeric value is a warp shuffle). */
template <class Sink>
__device__ __forceinline__ void title_record_parsed(const Dev &d, const ChunkRows &cr, const SbClass &C, const TitleTabs &T, const u32 *arena, u32 i, u32 ts,
                                                    u32 flags, bool first, Sink &s) {
  for (u32 k = 0; k < C.nnc; ++k) {
    const u32 f = T.ncf[k];
    const FieldClass &F = T.fc[f];
    const bool flag = (flags >> f) & 1u;
    if (F.kind == K_NUM) {
      const i32 v = (i32)parsed_value(d, cr, f, i);
      const i32 pv = __shfl_up_sync(0xFFFFFFFFu, v, 1);
      if (first) s.put((u32)wsub(v, F.min_v), F.bits_val);
      else if (!flag) {
        const u32 x = F.is_delta ? (u32)wsub(wsub(v, pv), F.min_d) : (u32)wsub(v, F.min_v);
        if (F.has_table) { /* x < diff for real records; lanes that only shadow a record may see anything */
          const u64 e = ((const u64 *)(arena + F.cl_off))[x < F.diff ? x : 0u];
          s.put((u32)e, (u32)(e >> 32));
        } else s.put(x, F.bits_num);
      }
      continue;
    }
    if (!first && flag) continue;
    const TokRef t = parsed_token(d, cr, C, F, f, i, ts);
    if (!F.is_len_const) s.put(t.len - F.min_len, F.bits_len);
    const u16 *sm = (const u16 *)(arena + F.slotmap_off);
    const u8 *a = d.in + t.off;
    for (u32 j = 0; j < t.len; ++j)
      if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) {
        const u32 tid = sm[j < 128 ? j : 128];
        const u64 e = ((const u64 *)(arena + C.chr_cl_off + (tid - C.tchr0) * 512u))[a[j]];
        s.put((u32)e, (u32)(e >> 32));
      }
  }
}

/* dynamic shared memory per warp: [32 * LPW_T staging words][CCW words][32 words for the info bits] */
__host__ __device__ __forceinline__ u32 enc_title_warp_bytes() { return (32u * LPW_T + CCW + 32u) * 4u; }

__global__ void __launch_bounds__(ENC_WARPS * 32) k_enc_title(Dev d) {
  extern __shared__ uint4 dyn_smem[];
  __shared__ TitleTabs TT;
  const u32 s = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  SbClass &C = d.cls[s];
  if (C.status || !C.fast) return;
  const u32 task = blockIdx.x * ENC_WARPS + w;
  if (blockIdx.x * ENC_WARPS >= C.ntask) return;
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const u32 R = C.R, nnc = C.nnc, nb_len = C.nb_len;
  load_title_tabs(C, TT);
  __syncthreads();
  if (task >= C.ntask) return;
  u32 *lp = (u32 *)((u8 *)dyn_smem + (size_t)w * enc_title_warp_bytes()), *cc = lp + 32 * LPW_T, *ci = cc + CCW;
  const u32 lp_a = (u32)__cvta_generic_to_shared(lp) + 4 * lane;
  u32 *tmp = d.tmp + C.tmp_base;
  WarpStream T;
  T.init(cc, tmp + C.info_words + (size_t)task * (C.strd_q + C.strd_d + C.strd_t) + C.strd_q + C.strd_d, C.strd_t);
  const u32 g0 = task * TASK_BLOCKS, g1 = min(g0 + TASK_BLOCKS, C.nblk);
  const u32 *flag_p = arena + C.flagbits_off;
  /* record of this lane in block g (idle lanes shadow the block's last record so that the warp stays converged) */
  u32 n_rs, n_te, n_se, n_fl;
  {
    const u32 r = P.first_rec + min(g0 * 32 + lane, R - 1);
    n_rs = d.rstart[r]; n_te = d.te[r]; n_se = d.se[r]; n_fl = nnc ? flag_p[g0] : 0u;
  }
  for (u32 g = g0; g < g1; ++g) {
    const u32 nrec = min(32u, R - g * 32);
    const bool active = lane < nrec;
    const u32 rs = n_rs, te = n_te, se = n_se, flags = n_fl;
    if (g + 1 < g1) {
      const u32 r = P.first_rec + min((g + 1) * 32 + lane, R - 1);
      n_rs = d.rstart[r]; n_te = d.te[r]; n_se = d.se[r]; n_fl = nnc ? flag_p[g + 1] : 0u;
    }
    { /* info stream: the read length of every record in nb_len bits (phyNGSC.cpp:732-742; always present) */
      ci[lane] = 0;
      __syncwarp();
      if (active && nb_len) {
        const u32 L = se - te - 1, pos = lane * nb_len, sh = pos & 31, v = L << (32 - nb_len);
        cc_or(ci, pos >> 5, v >> sh);
        if (sh + nb_len > 32) cc_or(ci, (pos >> 5) + 1, v << (32 - sh));
      }
      __syncwarp();
      if (lane < (nrec * nb_len + 31) / 32) tmp[g * nb_len + lane] = ci[lane];
    }
    if (nnc) {
      const ChunkRows cr = chunk_rows(d, P, g / (CH / 32));
      const u32 i = min(g * 32 + lane, R - 1) & (CH - 1);
      LaneSink sk; sk.init(SmemStore{lp_a}, LPW_T);
      if (lane == 0) {
        u32 v = 0;
        for (u32 k = 0; k < nnc; ++k) v = (v << 1) | ((flags >> TT.ncf[k]) & 1u);
        sk.put(v, nnc);
      }
      title_record_parsed(d, cr, C, TT, arena, i, rs, flags, lane == 0, sk);
      u32 nbits = sk.finish();
      if (sk.over) T.over = true;
      if (!active) nbits = 0;
      __syncwarp();
      T.append(lp + lane, nbits);
      T.pad_to_byte(); /* FlushPartialWordBuffer per 32-record block (tasks.cpp:508) */
    }
  }
  const u32 tbits = T.finish();
  if (__any_sync(0xFFFFFFFFu, T.over)) { if (lane == 0) atomicMin(&C.status, (i32)E_CAPACITY); return; }
  if (lane == 0) arena[C.task_off + 2 * C.ntask + task] = tbits >> 3;
}

}  // namespace phy
