/*
 * phy_encode.cuh -- the single-walk encoder kernels (included by phy_b200.cu after phy_kernels.cuh).
 *
 *   k_slots      per subblock: bounds from the tables' longest codes, slot sizes, region of the temporary buffer,
 *                and the decision single-walk (SbClass::fast) or two-walk (k_lengths + k_emit)
 *   k_enc_title  (phy_title.cuh) warp = task of 8 title blocks (32 records each, lane = record): info length bits and
 *                title tokens (phyNGSC.cpp:732-742, tasks.cpp:393-509) from the parsed title rows k_stat1 left
 *   k_enc_qd<G>  warp = task of 256 records, G lanes per record (each a run of read positions): quality and DNA codes
 *                (tasks.cpp:609-619, 544-557); record spans staged by the bulk-copy engine
 *   k_place      moves every task's run from the temporary buffer to its final bit position and writes the headers
 *
 * Every record byte is walked once: a lane appends its codes to lane-private staging words in shared memory
 * (LaneSink), the warp concatenates the lanes' pieces (WarpStream: scan of the bit counts, OR into a small buffer,
 * whole words leave for the task's slot with coalesced stores), and the task's exact length is known afterwards.
 */
#pragma once
#include "phy_fast.cuh"
#include "phy_kernels.cuh"

namespace phy {

constexpr u32 CCW = 192;         /* words of a warp's concatenation buffer                                     */
constexpr u32 ENC_WARPS = 8;     /* warps (tasks) per CTA of the encoder kernels                                */
constexpr u32 QD_ROUNDS_MAX = 64; /* rounds of a task in k_enc_qd: 256 records / (32 / G), G <= 8                */

/* lane-private staging in shared memory: word k of lane l at [k * 32 + l] */
struct SmemStore {
  u32 addr; /* shared-memory byte address of the lane's next word */
  __device__ __forceinline__ void put(u32 w) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(w) : "memory"); }
  __device__ __forceinline__ void next() { addr += 128; }
};
typedef LaneSinkT<SmemStore> LaneSink;

/* ---- warp-level stream ------------------------------------------------------------------------------------------- */
/* One output stream of one task.  append(): all 32 lanes call it with their staged piece; pieces are concatenated in lane
 * order behind what the task has emitted so far.  When the lanes' bits do not fit the concatenation buffer at once the
 * lanes go in as many turns as needed (a turn takes the longest prefix of lanes that fits). */
struct WarpStream {
  StreamState st;
  u32 *cc;          /* CCW words of shared memory, zero except cc[0]'s carry bits */
  u32 *slot;        /* the task's slot in the temporary buffer */
  u32 slot_words;
  bool over;
  __device__ __forceinline__ void init(u32 *cc_, u32 *slot_, u32 slot_words_) {
    st.init(); cc = cc_; slot = slot_; slot_words = slot_words_; over = false;
    for (u32 j = threadIdx.x & 31; j < CCW; j += 32) cc[j] = 0;
    __syncwarp();
  }
  /* `bits` more bits stand behind the carry in cc: whole words leave, the rest becomes the new carry */
  __device__ __forceinline__ void flush(u32 bits) {
    const u32 lane = threadIdx.x & 31;
    const u32 nf = st.full_words(bits);
    const bool room = st.tpos + nf + 1 <= slot_words;
    if (!room) over = true;
    u32 *dst = slot + st.tpos;
#pragma unroll 1
    for (u32 j = lane; j < nf; j += 32) { const u32 v = cc[j]; if (room) dst[j] = v; cc[j] = 0; } /* a few trips: kept rolled */
    const u32 rem = cc[nf]; /* no lane clears this word in the loop above */
    __syncwarp();
    if (lane == 0) { cc[nf] = 0; cc[0] = rem; }
    __syncwarp();
    st.advance(bits);
  }
  __device__ __forceinline__ void append(const u32 *lp, u32 nbits) {
    const u32 lane = threadIdx.x & 31;
    u32 incl = nbits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (u32)o) incl += y; }
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31), excl = incl - nbits;
    if (total <= (CCW - 2) * 32 - st.carry) { /* the common case: everything fits at once */
      if (nbits) lane_concat(cc, st.carry + excl, lp, 32u, nbits);
      __syncwarp();
      flush(total);
      return;
    }
    u32 base = 0;
    while (base < total) {
      const u32 room = (CCW - 2) * 32 - st.carry;
      const bool go = excl >= base && incl - base <= room;
      const u32 end = __reduce_max_sync(0xFFFFFFFFu, go ? incl : base);
      if (end == base) { over = true; break; } /* a single lane's piece exceeds the buffer: the staging bound was wrong */
      if (go && nbits) lane_concat(cc, st.carry + excl - base, lp, 32u, nbits);
      __syncwarp();
      flush(end - base);
      base = end;
    }
  }
  __device__ __forceinline__ void pad_to_byte() { const u32 p = st.pad_to_byte(); if (p) flush(p); }
  /* last partial word; returns the task's bits */
  __device__ __forceinline__ u32 finish() {
    if ((threadIdx.x & 31) == 0 && st.carry && !over && st.tpos < slot_words) slot[st.tpos] = cc[0];
    return st.total;
  }
};

/* ---- slots --------------------------------------------------------------------------------------------------------- */
/* one warp per subblock, after k_huff */
__global__ void __launch_bounds__(32) k_slots(Dev d) {
  const u32 s = blockIdx.x, lane = threadIdx.x;
  SbClass &C = d.cls[s];
  if (C.status) return;
  const u32 *arena = d.arena + (size_t)s * d.arena_words;
  const TableDesc *td = (const TableDesc *)(arena + C.tabdesc_off);
  bool fast = d.fg.g != 0; /* g == 0: the single-walk kernels are switched off */
  /* quality: sum and maximum of the tables' longest codes over the read positions */
  u32 qsum = 0, qmax = 0;
  for (u32 p = 1 + lane; p <= C.max_qlen; p += 32) { const u32 m = td[C.tq0 + p].maxlen; qsum += m; qmax = max(qmax, m); }
  qsum = __reduce_add_sync(0xFFFFFFFFu, qsum); qmax = __reduce_max_sync(0xFFFFFFFFu, qmax);
  const u32 tb = __reduce_add_sync(0xFFFFFFFFu, title_bound_part(C, arena, td, lane, 32u));
  if (lane) return;
  const u32 pk_bytes = (C.max_qlen + 1) * C.nq * 2u;
  if (C.qpk_bad || pk_bytes > d.fg.pk_bytes) fast = false; /* the walkers of the single-walk kernels read the packed tables from shared memory */
  atomicMax(&d.hdr->max_qcode, qmax);
  const u32 g = d.fg.g ? d.fg.g : 1u, seg = seg_len(C.max_qlen, g);
  const u32 dper = C.plain ? 2u : td[C.tdna].maxlen;
  if (seg * qmax > 32u * (d.fg.lpw_q - 1) || seg * dper > 32u * (d.fg.lpw_q - 1)) fast = false;
  if (C.nnc + tb > 32u * (LPW_T - 1)) fast = false;
  if (C.nb_len > 24) fast = false;
  const u64 rq = (u64)TASK_RECORDS * qsum, rd = (u64)TASK_RECORDS * C.max_qlen * dper;
  const u64 rt = (u64)TASK_BLOCKS * ((C.nnc + 32ull * tb + 7) / 8 * 8);
  if (rq > 0x7FFFFFFFull || rd > 0x7FFFFFFFull || rt > 0x7FFFFFFFull) fast = false;
  if (fast) {
    C.strd_q = (u32)(rq / 32) + 3; C.strd_d = (u32)(rd / 32) + 3; C.strd_t = (u32)(rt / 32) + 3;
    C.info_words = (u32)(((u64)C.R * C.nb_len + 31) / 32) + 1;
    const u64 words = C.info_words + (u64)C.ntask * (C.strd_q + C.strd_d + C.strd_t);
    const u64 base = atomicAdd((unsigned long long *)d.tmp_used, (unsigned long long)words);
    C.tmp_base = base;
    if (base + words > d.tmp_cap) fast = false; /* no room left in the temporary buffer: two-walk kernels */
  }
  C.fast = fast ? 1u : 0u;
}

/* ---- quality + DNA -------------------------------------------------------------------------------------------------- */
/* dynamic shared memory: [pk_bytes packed quality tables] then per warp [nbuf stages of qd_stage bytes][32 * lpw_q staging
 * words][CCW words quality][CCW words DNA] */
__device__ __forceinline__ u32 enc_qd_warp_bytes(const Dev &d) { return d.qd_nbuf * d.qd_stage + (32u * d.fg.lpw_q + 2u * CCW) * 4u; }

#ifndef PHY_QD_MINB
#define PHY_QD_MINB 3
#endif
template <int G>
__global__ void __launch_bounds__(ENC_WARPS * 32, PHY_QD_MINB) k_enc_qd(Dev d) {
  constexpr u32 RW = 32 / G; /* records per round of a warp */
  extern __shared__ uint4 dyn_smem[];
  __shared__ __align__(16) u8 codes[512];
  __shared__ __align__(16) u8 xq[256];
  __shared__ __align__(8) u64 bars[ENC_WARPS][2];
  __shared__ u32 r_lo[ENC_WARPS][QD_ROUNDS_MAX + 1];
  const u32 s = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  SbClass &C = d.cls[s];
  if (C.status || !C.fast) return;
  if (blockIdx.x * ENC_WARPS >= C.ntask) return;
  const u32 task = blockIdx.x * ENC_WARPS + w;
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const TableDesc *td = (const TableDesc *)(arena + C.tabdesc_off);
  const u32 R = C.R;
  WalkTabs T;
  load_walk_tabs(d, C, arena, td, codes, xq, (u16 *)dyn_smem, T);
  const u32 rec0 = min(task * TASK_RECORDS, R), rec1 = min(rec0 + TASK_RECORDS, R);
  const u32 nround = (rec1 - rec0 + RW - 1) / RW;
  for (u32 k = lane; k <= nround; k += 32) r_lo[w][k] = d.rstart[P.first_rec + min(rec0 + k * RW, rec1)];
  u8 *wb = (u8 *)dyn_smem + d.fg.pk_bytes + (size_t)w * enc_qd_warp_bytes(d);
  const u32 nbuf = d.qd_nbuf, stage_bytes = d.qd_stage;
  u32 *lp = (u32 *)(wb + nbuf * stage_bytes), *ccq = lp + 32 * d.fg.lpw_q, *ccd = ccq + CCW;
  const u32 stage_a = (u32)__cvta_generic_to_shared(wb), bar_a = (u32)__cvta_generic_to_shared(&bars[w][0]);
  const u32 lp_a = (u32)__cvta_generic_to_shared(lp) + 4 * lane;
  if (lane == 0) { mbar_init(bar_a, 1); mbar_init(bar_a + 8, 1); mbar_fence_init(); }
  __syncthreads();
  if (task >= C.ntask || rec0 >= rec1) return;
  {
    bool fits = true;
    for (u32 k = lane; k < nround; k += 32) fits = fits && r_lo[w][k + 1] - (r_lo[w][k] & ~15u) + 16 <= stage_bytes;
    if (!__all_sync(0xFFFFFFFFu, fits)) { if (lane == 0) atomicMin(&C.status, (i32)E_UNSUPPORTED); return; } /* records far beyond the reference's 500-byte domain */
  }
  u32 *slot = d.tmp + C.tmp_base + C.info_words + (size_t)task * (C.strd_q + C.strd_d + C.strd_t);
  WarpStream Q, D;
  Q.init(ccq, slot, C.strd_q);
  D.init(ccd, slot + C.strd_q, C.strd_d);
  auto request = [&](u32 k) { span_request(d.in, r_lo[w][k], r_lo[w][k + 1], stage_a + (k % nbuf) * stage_bytes, bar_a + 8 * (k % nbuf)); };
  if (lane == 0) for (u32 k = 0; k < nbuf && k < nround; ++k) request(k);
  u32 phases = 0;
  const u32 sub = lane / G, part = lane % G;
  u32 n_te, n_se, n_kx;
  { const u32 r = P.first_rec + min(rec0 + sub, rec1 - 1); n_te = d.te[r]; n_se = d.se[r]; n_kx = d.kx[r]; }
  for (u32 k = 0; k < nround; ++k) {
    const u32 i = rec0 + k * RW + sub;
    const bool active = i < rec1;
    const u32 te = n_te, se = n_se, kx = n_kx;
    if (k + 1 < nround) { const u32 r = P.first_rec + min(rec0 + (k + 1) * RW + sub, rec1 - 1); n_te = d.te[r]; n_se = d.se[r]; n_kx = d.kx[r]; }
    const u32 sb = k % nbuf;
    mbar_wait(bar_a + 8 * sb, (phases >> sb) & 1u); phases ^= 1u << sb;
    const u8 *b = wb + sb * stage_bytes - (r_lo[w][k] & ~15u);
    const u32 L = se - te - 1, seg = seg_len(L, G);
    const u32 a = min(L, part * seg), n = active ? min(L, a + seg) - a : 0u;
    const bool xfer = active && (kx >> 15);
    LaneSink sk; sk.init(SmemStore{lp_a}, d.fg.lpw_q);
    quality_walk(b + se + 3 + a, b + te + 1 + a, n, xfer, T, sk, a);
    const u32 qbits = sk.finish();
    bool over = sk.over;
    __syncwarp();
    Q.append(lp + lane, qbits);
    sk.init(SmemStore{lp_a}, d.fg.lpw_q);
    dna_walk(b + te + 1 + a, n, xfer, T, sk);
    const u32 dbits = sk.finish();
    over = over || sk.over;
    __syncwarp();
    D.append(lp + lane, dbits);
    if (over) Q.over = true;
    if (k + nbuf < nround) { /* every lane has left the stage */
      __syncwarp();
      if (lane == 0) request(k + nbuf);
    }
  }
  const u32 qb = Q.finish(), db = D.finish();
  if (__any_sync(0xFFFFFFFFu, Q.over || D.over)) { if (lane == 0) atomicMin(&C.status, (i32)E_CAPACITY); return; }
  if (lane == 0) { arena[C.task_off + task] = qb; arena[C.task_off + C.ntask + task] = db; }
}

/* ---- placement ------------------------------------------------------------------------------------------------------- */
/* one warp: `nbits` logical bits at src -> bit position dbit of the (zeroed) output, byte order of BitStream */
__device__ __forceinline__ void place_run(u32 *outw, u64 dbit, const u32 *src, u32 nbits) {
  if (!nbits) return;
  const u32 lane = threadIdx.x & 31, sh = (u32)(dbit & 31), nsrc = (nbits + 31) / 32, nd = (sh + nbits + 31) / 32;
  u32 *dst = outw + (dbit >> 5);
  /* words 1 .. nd-2 belong to this run alone: plain stores, four per lane in flight; the first and the last word may be
   * shared with a neighbouring run */
  u32 j = 1 + lane;
  for (; j + 96 < nd - 1; j += 128) {
    const u32 a0 = src[j - 1], a1 = src[j], b0 = src[j + 31], b1 = src[j + 32], c0 = src[j + 63], c1 = src[j + 64], d0 = src[j + 95], d1 = src[j + 96];
    dst[j] = bswap32(sh ? (a0 << (32 - sh)) | (a1 >> sh) : a1);
    dst[j + 32] = bswap32(sh ? (b0 << (32 - sh)) | (b1 >> sh) : b1);
    dst[j + 64] = bswap32(sh ? (c0 << (32 - sh)) | (c1 >> sh) : c1);
    dst[j + 96] = bswap32(sh ? (d0 << (32 - sh)) | (d1 >> sh) : d1);
  }
  for (; j + 1 < nd; j += 32) dst[j] = bswap32(shifted_word(src, nsrc, sh, j));
  if (lane < 2) {
    const u32 e = lane ? nd - 1 : 0u;
    if (lane == 0 || nd > 1) {
      const u32 v = bswap32(shifted_word(src, nsrc, sh, e));
      const bool shared = (e == 0 && sh) || (e == nd - 1 && ((sh + nbits) & 31u));
      if (shared) { if (v) atomicOr(dst + e, v); } else dst[e] = v;
    }
  }
}

__global__ void __launch_bounds__(256) k_place(Dev d) {
  const u32 s = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const SbClass &C = d.cls[s];
  if (C.status || !C.fast) return;
  const u32 *arena = d.arena + (size_t)s * d.arena_words;
  u8 *out = d.out + C.out_off;
  u32 *outw = (u32 *)d.out;
  const u64 obase = C.out_off;
  const u32 o_title = C.info_len, o_qual = o_title + C.title_len, o_dna = o_qual + C.qual_len;
  if (blockIdx.x == 0) { /* fixed part of the info stream (phyNGSC.cpp:719-730) and the three staged headers */
    if (threadIdx.x == 0) {
      u8 fx[INFO_FIXED];
      ByteWriter bw; bw.p = fx; bw.n = 0;
      bw.word(C.R); bw.word(C.max_qlen); bw.word(C.max_slen);
      bw.byte((u8)C.nsym); bw.byte(0); bw.byte((u8)C.nq); bw.word(C.flags);
      for (u32 i = 0; i < INFO_FIXED; ++i) or_byte(out, i, fx[i]);
    }
    const u8 *hs = (const u8 *)(arena + C.stage_off);
    for (u32 i = threadIdx.x; i < C.thdr_len; i += 256) or_byte(out, o_title + i, hs[i]);
    for (u32 i = threadIdx.x; i < C.qhdr_len; i += 256) or_byte(out, o_qual + i, hs[C.thdr_cap + i]);
    for (u32 i = threadIdx.x; i < C.dhdr_len; i += 256) or_byte(out, o_dna + i, hs[C.thdr_cap + C.qhdr_cap + i]);
  }
  const u32 *tmp = d.tmp + C.tmp_base;
  const u32 ntask = C.ntask, strd = C.strd_q + C.strd_d + C.strd_t;
  const u32 *len3 = arena + C.task_off, *base3 = len3 + 3 * ntask;
  const u64 info_bits = (u64)C.R * C.nb_len;
  const u32 ni = (u32)((info_bits + 32ull * PIECE_WORDS - 1) / (32ull * PIECE_WORDS));
  const u32 cq = (C.strd_q + PIECE_WORDS - 1) / PIECE_WORDS, cd = (C.strd_d + PIECE_WORDS - 1) / PIECE_WORDS, ct = (C.strd_t + PIECE_WORDS - 1) / PIECE_WORDS;
  const u32 per_task = cq + cd + ct, npieces = ni + ntask * per_task;
  const u64 bit_q = (obase + o_qual + C.qhdr_len) * 8, bit_d = (obase + o_dna + C.dhdr_len) * 8, bit_t = (obase + o_title + C.thdr_len) * 8;
  for (u32 p = blockIdx.x * 8 + w; p < npieces; p += gridDim.x * 8) {
    if (p < ni) { /* info length bits: one dense run per subblock */
      const u64 b0 = (u64)p * PIECE_WORDS * 32;
      const u32 nb = (u32)min((u64)PIECE_WORDS * 32, info_bits - b0);
      place_run(outw, (obase + INFO_FIXED) * 8 + b0, tmp + (size_t)p * PIECE_WORDS, nb);
      continue;
    }
    const u32 q = p - ni, task = q / per_task, r = q % per_task;
    const u32 kind = r < cq ? 0u : r < cq + cd ? 1u : 2u, piece = kind == 0 ? r : kind == 1 ? r - cq : r - cq - cd;
    u32 bits = len3[kind * ntask + task];
    if (kind == 2) bits *= 8; /* title: bytes */
    const u32 b0 = piece * PIECE_WORDS * 32;
    if (b0 >= bits) continue;
    const u32 nb = min(PIECE_WORDS * 32, bits - b0);
    const u32 *src = tmp + C.info_words + (size_t)task * strd + (kind == 0 ? 0u : kind == 1 ? C.strd_q : C.strd_q + C.strd_d) + piece * PIECE_WORDS;
    const u64 tb = (u64)base3[kind * ntask + task] * (kind == 2 ? 8u : 1u);
    place_run(outw, (kind == 0 ? bit_q : kind == 1 ? bit_d : bit_t) + tb + b0, src, nb);
  }
}

}  // namespace phy
