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

/* dynamic shared memory: [sq_rows * SQ_ROWW words: private table][per warp: sq_nbuf stages of sq_stage bytes] */
template <int G>
__global__ void __launch_bounds__(SQ_WARPS * 32) k_seqstat(Dev d) {
  constexpr u32 RW = 32 / G;
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
  const u32 hist_words = d.sq_rows * SQ_ROWW, hist_a = (u32)__cvta_generic_to_shared(hist);
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
    const u8 *b = wb + sb * stage_bytes - (r_lo[w][k] & ~15u);
    const u32 L = se - te - 1;
    bool rec_ok = active;
    if (active && part == 0) { /* the four-line shape (phyNGSC.cpp:466-471 assumes it) */
      if (L == 0 || b[se + 1] != '+' || b[se + 2] != '\n' || nx != 2 * se - te + 3) { err = E_MALFORMED; rec_ok = false; }
      else if (L > (u32)MAX_READ || L + 1 > RAW_ROWS) { err = E_UNSUPPORTED; rec_ok = false; } /* the raw quality table has RAW_ROWS rows */
      else if (i == 0) { /* colour space, phyNGSC.cpp:473-487: not implemented */
        const u8 c0b = b[te + 1], c1b = b[te + 2];
        if ((c0b >= '0' && c0b <= '3') || (c1b >= '0' && c1b <= '3')) err = E_COLORSPACE;
      }
    }
    rec_ok = __shfl_sync(0xFFFFFFFFu, rec_ok, lane & ~(G - 1)) != 0;
    const u32 seg = seg_len(L, G);
    const u32 a = rec_ok ? min(L, part * seg) : 0u, n = rec_ok ? min(L, a + seg) - a : 0u;
    const u8 *sp = b + te + 1 + a;
    const u32 q_a = stage_a + sb * stage_bytes + (se + 3 + a - (r_lo[w][k] & ~15u)); /* shared address of this lane's first quality byte */
    /* bases of this lane's run, four per step: the 2-bit index (c >> 1) & 3 selects the byte the base must equal
     * ("ACTG"[idx]) and a one-hot presence byte with two byte permutes */
    u32 bad = 0, ph = 0;
    {
      const u32 al = (u32)(size_t)sp & 3u;
      const u32 *wp = (const u32 *)(sp - al);
      u32 w0 = wp[0], j = 0;
      for (; j + 4 <= n; j += 4) {
        const u32 w1 = *++wp;
        const u32 v = __funnelshift_r(w0, w1, al * 8);
        w0 = w1;
        const u32 z = (v >> 1) & 0x03030303u;
        const u32 sel = __byte_perm(z | (z >> 4), 0, 0x4420);
        bad |= __byte_perm(0x47544341u, 0, sel) ^ v;
        ph |= __byte_perm(0x08040201u, 0, sel);
      }
      for (; j < n; ++j) { const u32 f = dlut[sp[j]]; ph |= f; bad |= f ? 0u : 1u; }
    }
    u32 xfer = 0, namb = 0;
    if (__any_sync(0xFFFFFFFFu, bad != 0)) { /* some base of the round is not A/C/G/T: ambiguity transfer per record (phyNGSC.cpp:549-588) */
      const bool rec_bad = grp_or<G>(bad) != 0;
      u32 okf = 1, nul = 0;
      if (rec_bad) {
        const u32 qa0 = q_a;
        ph = 0;
        for (u32 j = 0; j < n; ++j) {
          const u8 c = sp[j];
          const u32 f = dlut[c];
          if (f) ph |= f;
          else { const u32 q = lds_u8(qa0 + j); ++namb; nul |= c == 0 ? 1u : 0u; if (xq[c] == 0 || q < 33 || q > 40) okf = 0; }
        }
      }
      const u32 namb_t = grp_add<G>(namb);
      okf = grp_add<G>(okf) == G ? 1u : 0u;
      nul = grp_or<G>(nul);
      xfer = (namb_t && okf) ? 1u : 0u;
      if (rec_bad && !xfer) for (u32 j = 0; j < n; ++j) { const u8 c = sp[j]; if (!dlut[c]) s_dna[c] = 1; } /* the byte stays in the DNA */
      if (nul) err = E_UNSUPPORTED;
      namb = namb_t;
    }
    ph |= ph >> 16; ph |= ph >> 8;
    w_pres |= ph & 0xFu;
    if (rec_ok) {
      const u32 kept = xfer ? L - namb : L;
      w_maxq = max(w_maxq, L); w_maxs = max(w_maxs, kept); w_invminq = max(w_invminq, ~L);
      if (part == 0) d.kx[P.first_rec + i] = (u16)(kept | (xfer << 15));
    }
    /* quality histogram of this lane's run, started at a staggered position */
    if (n) {
      u32 p = sub % n;
      const u32 xm = xfer ? 0xFFu : 0u;
      const u32 sp_a = q_a - 3 - L; /* the base under quality byte j is L + 3 bytes before it */
#pragma unroll 2
      for (u32 j = 0; j < n; ++j) {
        u32 q = lds_u8(q_a + p);
        if (xm) q += xq[lds_u8(sp_a + p)];
        const u32 c = q - 33u, pos = a + p;
        if (c < 95u) sm_red_add(hist_a + (pos * SQ_ROWW + (c >> 1)) * 4u, 1u << ((c & 1u) * 16u));
        else atomicAdd(raw + (size_t)(pos + 1) * 256 + q, 1u);
        if (++p == n) p = 0;
      }
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
  for (u32 i = tid; i < rows * (SQ_ROWW - 1); i += SQ_WARPS * 32) {
    const u32 pr = i / (SQ_ROWW - 1), wd = i % (SQ_ROWW - 1), v = hist[pr * SQ_ROWW + wd];
    if (v & 0xFFFFu) atomicAdd(raw + (size_t)(pr + 1) * 256 + 33 + 2 * wd, v & 0xFFFFu);
    if (v >> 16) atomicAdd(raw + (size_t)(pr + 1) * 256 + 33 + 2 * wd + 1, v >> 16);
  }
}

}  // namespace phy
