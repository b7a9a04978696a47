/*
 * phy_seqstat.cuh -- sequence / quality statistics of a subblock in one walk (included by phy_b200.cu).
 *
 *   k_zero_raw    clears the raw per-position quality tables of the batch's subblocks
 *   k_seqstat<G>  per record: validation of the four-line shape, ambiguity transfer (phyNGSC.cpp:549-588), kept DNA
 *                 length, DNA symbol presence, read length extremes, and the per-position quality histogram
 *                 raw[position + 1][byte] (tasks.cpp:260-286, counted after the transfer)
 *
 * Mapping: a warp owns a task of 256 consecutive records and walks it in rounds of 32 / G records; G lanes share a record,
 * each taking a run of read positions (the same split as k_enc_qd), so the bytes a warp keeps staged and the length of a
 * lane's loop shrink with G for long reads.  Record spans arrive through the bulk-copy engine (double-buffered).
 *
 * Histogram: a CTA keeps a private table in shared memory, row = read position, 16-bit counters packed two per word
 * (a CTA counts at most 8 x 256 records).  Lanes of a warp start their runs at staggered positions, so the shared-memory
 * reductions of one instruction go to different rows (49 words apart: different banks) and never to the same counter.
 * Bytes outside 33..127 (transferred ambiguity codes, garbage) go straight to the global table.
 */
#pragma once
#include "phy_kernels.cuh"

namespace phy {

constexpr u32 SQ_WARPS = 8;
constexpr u32 SQ_ROWW = 49;       /* words per row of the private table: 95 counters of 16 bits + padding, odd */
constexpr u32 SQ_ROUNDS_MAX = 64;

__global__ void __launch_bounds__(256) k_zero_raw(Dev d) {
  uint4 *raw = (uint4 *)raw_table(d, blockIdx.y);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (u32 i = blockIdx.x * 256 + threadIdx.x; i < RAW_WORDS / 4; i += gridDim.x * 256) raw[i] = z;
}

__device__ __forceinline__ void sm_red_add(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

/* OR / ADD / AND over the G lanes that share a record (G a power of two, lanes of a record are adjacent) */
template <int G> __device__ __forceinline__ u32 grp_or(u32 v) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
template <int G> __device__ __forceinline__ u32 grp_add(u32 v) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__device__ __forceinline__ u32 lds_u32(u32 addr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
/* red.shared with an immediate byte offset (the four positions of a step lie in consecutive rows) */
template <int OFF> __device__ __forceinline__ void sm_red_add_at(u32 addr, u32 v) { asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory"); }

/* One step of four read positions: counters of the private table.  WIDE: 32-bit counters, 97 words per row, the byte
 * itself indexes the row (row base shifted by -33).  Packed: 16-bit counters two per word, 49 words per row. */
template <bool WIDE> struct SqRow {
  static constexpr u32 ROWB = WIDE ? 97u * 4u : SQ_ROWW * 4u; /* bytes per row */
  template <int I> static __device__ __forceinline__ void count(u32 rowbase, u32 vq) {
    const u32 c = __byte_perm(vq, 0, 0x4440 + I); /* byte I, zero-extended */
    if (WIDE) sm_red_add_at<I * (int)ROWB>(rowbase + c * 4u, 1u);
    else { const u32 k = c - 33u; sm_red_add_at<I * (int)ROWB>(rowbase + (k >> 1) * 4u, 1u << ((k & 1u) * 16u)); }
  }
  static __device__ __forceinline__ void count_byte(u32 rowbase, u32 q) { /* q in 33..127 */
    if (WIDE) sm_red_add(rowbase + q * 4u, 1u);
    else { const u32 k = q - 33u; sm_red_add(rowbase + (k >> 1) * 4u, 1u << ((k & 1u) * 16u)); }
  }
};

/* dynamic shared memory: [private table: sq_rows rows][per warp: sq_nbuf stages of sq_stage bytes] */
template <int G, bool WIDE>
__global__ void __launch_bounds__(SQ_WARPS * 32) k_seqstat(Dev d) {
  constexpr u32 RW = 32 / G;
  typedef SqRow<WIDE> Row;
  extern __shared__ uint4 dyn_smem[];
  __shared__ u32 s_dna[256];
  __shared__ u32 s_maxq, s_maxs, s_invminq;
  __shared__ i32 s_err;
  __shared__ __align__(16) u8 xq[256], dlut[256];
  __shared__ __align__(8) u64 bars[SQ_WARPS][2];
  __shared__ u32 r_lo[SQ_WARPS][SQ_ROUNDS_MAX + 1];
  const u32 s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const SbPlan P = d.plans[s];
  const u32 R = P.n_records, ntask = (R + TASK_RECORDS - 1) / TASK_RECORDS;
  if (P.status || blockIdx.x * SQ_WARPS >= ntask) return;
  const u32 task = blockIdx.x * SQ_WARPS + w;
  u32 *hist = (u32 *)dyn_smem;
  const u32 hist_words = d.sq_rows * (Row::ROWB / 4), hist_a = (u32)__cvta_generic_to_shared(hist);
  for (u32 i = tid; i < hist_words; i += SQ_WARPS * 32) hist[i] = 0;
  for (u32 i = tid; i < 256; i += SQ_WARPS * 32) { s_dna[i] = 0; dlut[i] = (u8)(i == 'A' ? 1 : i == 'C' ? 2 : i == 'T' ? 4 : i == 'G' ? 8 : 0); }
  load_xq(xq);
  if (tid == 0) { s_maxq = 0; s_maxs = 0; s_invminq = 0; s_err = 0; }
  const u32 rec0 = min(task * TASK_RECORDS, R), rec1 = min(rec0 + TASK_RECORDS, R);
  const u32 nround = (rec1 - rec0 + RW - 1) / RW;
  for (u32 k = lane; k <= nround; k += 32) r_lo[w][k] = d.rstart[P.first_rec + min(rec0 + k * RW, rec1)];
  const u32 nbuf = d.sq_nbuf, stage_bytes = d.sq_stage;
  u8 *wb = (u8 *)dyn_smem + ((hist_words * 4 + 15u) & ~15u) + (size_t)w * nbuf * stage_bytes;
  const u32 stage_a = (u32)__cvta_generic_to_shared(wb), bar_a = (u32)__cvta_generic_to_shared(&bars[w][0]);
  if (lane == 0) { mbar_init(bar_a, 1); mbar_init(bar_a + 8, 1); mbar_fence_init(); }
  __syncthreads();
  u32 *raw = raw_table(d, s);
  bool fits = true;
  for (u32 k = lane; k < nround; k += 32) fits = fits && r_lo[w][k + 1] - (r_lo[w][k] & ~15u) + 16 <= stage_bytes;
  fits = __all_sync(0xFFFFFFFFu, fits);
  i32 err = 0;
  if (!fits) err = E_UNSUPPORTED; /* records far beyond the reference's 500-byte domain */
  auto request = [&](u32 k) { span_request(d.in, r_lo[w][k], r_lo[w][k + 1], stage_a + (k % nbuf) * stage_bytes, bar_a + 8 * (k % nbuf)); };
  if (fits && lane == 0) for (u32 k = 0; k < nbuf && k < nround; ++k) request(k);
  u32 phases = 0;
  const u32 sub = lane / G, part = lane % G;
  const u32 rb0 = hist_a - (WIDE ? 33u * 4u : 0u); /* row base of position 0 (WIDE: the byte value indexes the row directly) */
  u32 n_te = 0, n_se = 0, n_nx = 0;
  if (fits && nround) { const u32 r = P.first_rec + min(rec0 + sub, rec1 - 1); n_te = d.te[r]; n_se = d.se[r]; n_nx = d.rstart[r + 1]; }
  u32 w_pres = 0, w_maxq = 0, w_maxs = 0, w_invminq = 0;
  for (u32 k = 0; fits && k < nround; ++k) {
    const u32 i = rec0 + k * RW + sub;
    const bool active = i < rec1;
    const u32 te = n_te, se = n_se, nx = n_nx;
    if (k + 1 < nround) { const u32 r = P.first_rec + min(rec0 + (k + 1) * RW + sub, rec1 - 1); n_te = d.te[r]; n_se = d.se[r]; n_nx = d.rstart[r + 1]; }
    const u32 sb = k % nbuf;
    mbar_wait(bar_a + 8 * sb, (phases >> sb) & 1u); phases ^= 1u << sb;
    const u32 b_a = stage_a + sb * stage_bytes - (r_lo[w][k] & ~15u); /* shared address of batch position 0 */
    const u32 L = se - te - 1;
    bool rec_ok = active;
    if (active && part == 0) { /* the four-line shape (phyNGSC.cpp:466-471 assumes it) */
      if (L == 0 || lds_u8(b_a + se + 1) != '+' || lds_u8(b_a + se + 2) != '\n' || nx != 2 * se - te + 3) { err = E_MALFORMED; rec_ok = false; }
      else if (L > (u32)MAX_READ || L + 1 > RAW_ROWS) { err = E_UNSUPPORTED; rec_ok = false; } /* the raw quality table has RAW_ROWS rows */
      else if (i == 0) { /* colour space, phyNGSC.cpp:473-487: not implemented */
        const u32 c0b = lds_u8(b_a + te + 1), c1b = lds_u8(b_a + te + 2);
        if ((c0b >= '0' && c0b <= '3') || (c1b >= '0' && c1b <= '3')) err = E_COLORSPACE;
      }
    }
    rec_ok = __shfl_sync(0xFFFFFFFFu, rec_ok, lane & ~(G - 1)) != 0;
    const u32 seg = seg_len(L, G);
    const u32 a = rec_ok ? min(L, part * seg) : 0u, n = rec_ok ? min(L, a + seg) - a : 0u;
    const u32 s_a = b_a + te + 1 + a, q_a = s_a + L + 3; /* shared addresses of this lane's first base / quality byte */
    /* One walk over the lane's run, four positions per step, started at a staggered step (the lanes of a warp then add to
     * different rows): the bases are checked against "ACTG"[(c >> 1) & 3] with two byte permutes (and leave a one-hot
     * presence byte), the quality bytes are counted unless the base above them is not A/C/G/T -- those positions are
     * kept in `badpos` and settled once the record's ambiguity transfer is decided. */
    u32 ph = 0;
    u64 badpos = 0;
    const u32 ng = n >> 2;
    if (ng) {
      /* steps g0, g0 + 1, .., ng - 1, 0, .., g0 - 1: every lane makes exactly ng steps (lanes with equal run lengths stay
       * converged); the step counter wraps once, where the rolling words are reloaded */
      const u32 g0 = sub % ng;
      const u32 sal = s_a & 3u, qal = q_a & 3u;
      u32 g = g0;
      u32 sw = (s_a + 4 * g) - sal, qw = (q_a + 4 * g) - qal; /* aligned word addresses */
      u32 s0 = lds_u32(sw), q0 = lds_u32(qw);
      u32 rowbase = rb0 + (a + 4 * g) * Row::ROWB;
#pragma unroll 1
      for (u32 t = 0; t < ng; ++t) {
        sw += 4; qw += 4;
        const u32 s1 = lds_u32(sw), q1 = lds_u32(qw);
        const u32 vb = __funnelshift_r(s0, s1, sal * 8), vq = __funnelshift_r(q0, q1, qal * 8);
        const u32 z = (vb >> 1) & 0x03030303u;
        const u32 sel = __byte_perm(z | (z >> 4), 0, 0x4420);
        const u32 bad = __byte_perm(0x47544341u, 0, sel) ^ vb;
        const u32 in_range = ((vq | 0x80808080u) - 0x21212121u) & ~vq & 0x80808080u; /* bit 7 of a byte: 33 <= byte <= 127 */
        if (bad == 0 && in_range == 0x80808080u) {
          ph |= __byte_perm(0x08040201u, 0, sel);
          Row::template count<0>(rowbase, vq); Row::template count<1>(rowbase, vq);
          Row::template count<2>(rowbase, vq); Row::template count<3>(rowbase, vq);
        } else { /* rare: a base that is not A/C/G/T, or a quality byte outside 33..127 */
#pragma unroll
          for (u32 u = 0; u < 4; ++u) {
            const u32 c = (vb >> (8 * u)) & 0xFFu, q = (vq >> (8 * u)) & 0xFFu, f = dlut[c];
            if (!f) { badpos |= 1ull << (4 * g + u); continue; }
            ph |= f;
            if (q - 33u < 95u) Row::count_byte(rowbase + u * Row::ROWB, q);
            else atomicAdd(raw + (size_t)(a + 4 * g + u + 1) * 256 + q, 1u);
          }
        }
        ++g; rowbase += 4 * Row::ROWB;
        if (g == ng) { g = 0; sw = s_a - sal; qw = q_a - qal; s0 = lds_u32(sw); q0 = lds_u32(qw); rowbase = rb0 + a * Row::ROWB; }
        else { s0 = s1; q0 = q1; }
      }
    }
    for (u32 j = 4 * ng; j < n; ++j) { /* the last positions of a read whose length is not a multiple of four */
      const u32 c = lds_u8(s_a + j), q = lds_u8(q_a + j), f = dlut[c];
      if (!f) { badpos |= 1ull << j; continue; }
      ph |= f;
      if (q - 33u < 95u) Row::count_byte(rb0 + (a + j) * Row::ROWB, q);
      else atomicAdd(raw + (size_t)(a + j + 1) * 256 + q, 1u);
    }
    u32 xfer = 0, namb = 0;
    if (__any_sync(0xFFFFFFFFu, badpos != 0)) { /* ambiguity transfer, decided per record (phyNGSC.cpp:549-588) */
      u32 okf = 1, nul = 0;
      for (u64 m = badpos; m; m &= m - 1) {
        const u32 j = (u32)__ffsll((long long)m) - 1, c = lds_u8(s_a + j), q = lds_u8(q_a + j);
        ++namb; nul |= c == 0 ? 1u : 0u;
        if (xq[c] == 0 || q < 33 || q > 40) okf = 0;
      }
      namb = grp_add<G>(namb);
      okf = grp_add<G>(okf) == G ? 1u : 0u;
      nul = grp_or<G>(nul);
      xfer = (namb && okf) ? 1u : 0u;
      for (u64 m = badpos; m; m &= m - 1) { /* the quality byte under such a base: with the transferred code, or as it is */
        const u32 j = (u32)__ffsll((long long)m) - 1, c = lds_u8(s_a + j);
        const u32 q = lds_u8(q_a + j) + (xfer ? xq[c] : 0u);
        if (!xfer) s_dna[c] = 1; /* the byte stays in the DNA */
        if (q - 33u < 95u) Row::count_byte(rb0 + (a + j) * Row::ROWB, q);
        else atomicAdd(raw + (size_t)(a + j + 1) * 256 + q, 1u);
      }
      if (nul) err = E_UNSUPPORTED;
    }
    ph |= ph >> 16; ph |= ph >> 8;
    w_pres |= ph & 0xFu;
    if (rec_ok) {
      const u32 kept = xfer ? L - namb : L;
      w_maxq = max(w_maxq, L); w_maxs = max(w_maxs, kept); w_invminq = max(w_invminq, ~L);
      if (part == 0) d.kx[P.first_rec + i] = (u16)(kept | (xfer << 15));
    }
    if (k + nbuf < nround) { /* every lane has left the stage */
      __syncwarp();
      if (lane == 0) request(k + nbuf);
    }
  }
  { /* warp -> CTA */
    const u32 pr = __reduce_or_sync(0xFFFFFFFFu, w_pres);
    const u32 mq = __reduce_max_sync(0xFFFFFFFFu, w_maxq), ms = __reduce_max_sync(0xFFFFFFFFu, w_maxs), iq = __reduce_max_sync(0xFFFFFFFFu, w_invminq);
    const i32 e = __reduce_min_sync(0xFFFFFFFFu, err);
    if (lane == 0) {
      if (pr & 1u) s_dna['A'] = 1;
      if (pr & 2u) s_dna['C'] = 1;
      if (pr & 4u) s_dna['T'] = 1;
      if (pr & 8u) s_dna['G'] = 1;
      atomicMax(&s_maxq, mq); atomicMax(&s_maxs, ms); atomicMax(&s_invminq, iq);
      if (e) atomicMin(&s_err, e);
    }
  }
  __syncthreads();
  /* CTA -> subblock */
  SbAcc *A = d.acc + s;
  if (tid == 0) {
    if (s_err) atomicMin(&A->status, s_err);
    atomicMax(&A->max_qlen, s_maxq); atomicMax(&A->max_slen, s_maxs); atomicMax(&A->inv_min_qlen, s_invminq);
  }
  for (u32 i = tid; i < 256; i += SQ_WARPS * 32) if (s_dna[i]) atomicAdd(&A->dna_occ[i], s_dna[i]);
  const u32 rows = min(d.sq_rows, s_maxq);
  if (WIDE) {
    for (u32 i = tid; i < rows * 95; i += SQ_WARPS * 32) {
      const u32 pr = i / 95, c = i % 95, v = hist[pr * 97 + c];
      if (v) atomicAdd(raw + (size_t)(pr + 1) * 256 + 33 + c, v);
    }
  } else {
    for (u32 i = tid; i < rows * (SQ_ROWW - 1); i += SQ_WARPS * 32) {
      const u32 pr = i / (SQ_ROWW - 1), wd = i % (SQ_ROWW - 1), v = hist[pr * SQ_ROWW + wd];
      if (v & 0xFFFFu) atomicAdd(raw + (size_t)(pr + 1) * 256 + 33 + 2 * wd, v & 0xFFFFu);
      if (v >> 16) atomicAdd(raw + (size_t)(pr + 1) * 256 + 33 + 2 * wd + 1, v >> 16);
    }
  }
}

}  // namespace phy
