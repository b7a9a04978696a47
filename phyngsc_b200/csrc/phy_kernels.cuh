/*
 * phy_kernels.cuh -- sm_100a kernels of the phyNGSC subblock compressor (included by phy_b200.cu).
 *
 * Data layout in HBM for one batch (a slice of one rank's working region):
 *   in[len + slack]            the FASTQ bytes, batch-relative positions are uint32
 *   te[r], se[r], rstart[r]    record table: newline ending the title line / the sequence line, first byte
 *   kx[r]                      kept DNA length | transfer flag << 15          (phyNGSC.cpp:549-588)
 *   qoff[r], doff[r], toff[r]  bit offset of the record inside its 32-record block of the quality / DNA / title body
 *   plans[s], acc[s], cls[s]   per-subblock window, reduced statistics, coding decisions (phy_core.cuh)
 *   arena[s][arena_words]      per-subblock histograms, Huffman tables, tree blobs, header staging, per-block totals;
 *                              its last RAW_WORDS words are the raw per-position quality table of k_qhist
 *   out[]                      payloads info|title|quality|dna, 16-byte aligned per subblock
 *
 * Stages (every launch covers all subblocks of the batch, or of one subblock group -- see run_batch in phy_b200.cu):
 *   nl_count -> nl_scan -> nl_emit      record splitter            (phyNGSC.cpp:254-331)
 *   plan, spanmax                       window chaining, shared-memory sizing (phyNGSC.cpp:168-250, 744-755)
 *   stat1, xdelta (phy_title.cuh)       the one tokeniser pass: title field reductions, parsed title rows (phyNGSC.cpp:383-423, tasks.cpp:22-223)
 *   seqstat (phy_seqstat.cuh)           validation, ambiguity transfer, DNA alphabet, raw per-position quality histogram
 *                                                                  (phyNGSC.cpp:462-653; tasks.cpp:260-286)
 *   classify, zero_hist                 coding decisions, arena layout, coded quality tables (tasks.cpp:196-257)
 *   stat2 (phy_title.cuh), dnacount     numeric / char histograms, 32-record block descriptors (tasks.cpp:64-93,127-182),
 *                                       exact DNA symbol counts when the DNA is Huffman coded (tasks.cpp:233-236)
 *   huff                                one warp per table         (huffman.cpp:18-118)
 *   slots, enc_title, enc_qd            single-walk encoder (phy_encode.cuh, phy_title.cuh): every task's streams into a temporary buffer
 *   lengths -> layout -> outscan        bit lengths of the subblocks the single-walk kernels could not take, scans, header
 *                                       assembly, payload offsets
 *   zero_out -> place, emit             final placement of the tasks' runs / BitStream emission of the two-walk path
 *                                                                  (tasks.cpp:393-509,544-557,609-619)
 *
 * The kernels that read whole records (seqstat, enc_qd, dnacount, lengths, emit) stream them into shared memory with the
 * bulk-copy engine (cp.async.bulk + mbarrier): span_request / ChunkStage / WarpStage below; k_stat1 stages title lines only
 * (16-byte cp.async per lane).
 */
#pragma once
#include <cuda_runtime.h>
#include "phy_core.cuh"
#include "phy_fast.cuh"

namespace phy {

constexpr int CH = 128;            /* records per work item (4 warps; one warp = one 32-record title block) */
#ifndef PHY_NLT
#define PHY_NLT 256
#endif
constexpr int NLT = PHY_NLT;        /* threads per CTA of the record splitter                               */
constexpr int TILE = NLT * 64;     /* bytes per newline-index tile (NLT threads x 64 bytes)                */
constexpr int SUPER = 64;          /* tiles per supertile: the single-CTA scan runs over supertiles         */
constexpr u32 RAW_ROWS = 512;      /* rows of the raw per-position table kept at the end of a subblock's arena      */
constexpr u32 RAW_WORDS = RAW_ROWS * 256;
constexpr u32 PK_ESC = 0xF000u;   /* packed quality entries at or above this value: code longer than 12 bits, read the 64-bit entry */
constexpr u32 R0_MAX = 1024;       /* longest title line of record 0 kept in shared memory                 */
constexpr u32 TP_SAME = 0x7FFFu;   /* Dev::tp start value (15 bits) of lanes without a record              */
constexpr u32 TP_NUM = 0x8000u;    /* Dev::tp bit 15: the token is numeric (utils::is_num), tv holds its value; else tv / tc hold its first 4 / next 4 characters */

struct BatchHdr {      /* device -> host after the plan kernel and again after outscan */
  u32 NL, NR;          /* newlines found, complete records                                      */
  u32 S;               /* subblocks planned in this batch                                       */
  u32 max_chunks;      /* max over subblocks of ceil(n_records / CH)                            */
  u32 max_span;        /* widest 128-record span (16-byte aligned start, one record of halo)     */
  i32 status;          /* batch-level error (capacity ...)                                      */
  u64 total_out;       /* end of the output used (byte offset in d.out)                          */
  u64 out_begin;       /* where this group's payloads start (= the previous group's total_out)   */
  u64 next_pos;        /* region-relative position where the next window starts                 */
  u32 pad_q;
  u32 max_nf;          /* max over subblocks of the separator count of the first title        */
  u32 max_pk_bytes;    /* max over subblocks of the packed quality code tables ((max_qlen + 1) * n_qualities u16) */
  u32 max_len;         /* longest sequence line of the batch's subblocks */
  u32 max_span64, max_span32; /* widest 64- / 32-record span (k_qhist may stage smaller groups than the 128-record chunk) */
  u32 max_rec, max_tlen;      /* longest record and longest title line (both with their newline) */
  u32 max_qcode, pad_hdr;     /* longest quality code of the batch's subblocks (sizes the next batch's lane-private staging) */
};

struct SbOut {         /* device -> host, one per subblock */
  u64 out_off; u32 out_len; i32 status; u32 sec_len[4];
};

struct Dev {
  const u8 *in; u32 len;      /* batch bytes                                                   */
  u32 start_pos;              /* first record start inside the batch                           */
  u32 *te, *se, *rstart; u32 maxrec;
  u16 *kx; u32 *qoff, *doff, *toff;
  u32 *chunk_first, *chunk_last; /* numeric token values of the first / last record of every k_stat1 task (256 records), row = chunk_base + 2 * task, [row][MAXF] */
  /* Parsed titles (written by k_stat1, the only kernel that tokenises): per 32-record block the mask of fields in which
   * some record of the block differs from record 0 of its subblock, and for those fields one row of 32 entries each in tv
   * (a numeric token's value, utils::to_num; else the token's first four characters), tc (characters 4..7 of a non-numeric
   * token of 5..8 characters) and tp (token start inside the title line, 15 bits | TP_NUM | length << 16; start = TP_SAME
   * for lanes without a record).  Short tokens are thus complete in the rows (a numeric token is the decimal form of its
   * value) and the later kernels never go back to the input for them.  Block b of a subblock is block 4 * chunk_base + b of
   * the batch; row of (block, field): (block * nfs + field) * 32.  v0[subblock][field]: numeric value of record 0's token. */
  u32 *blk_mask, *tv, *tc, *tp, *v0; u32 nfs;
  u32 *tile_cnt, *tile_off; u32 ntiles;
  uint2 *nl_mask;             /* newline bit mask of the batch, 64 input bytes per element */
  PlanState *plan_state; SbPlan *plans; u32 max_sb;
  BatchHdr *hdr;
  SbAcc *acc; SbClass *cls; SbOut *sbout;
  u32 *arena; u32 arena_words;
  u8 *out; u64 out_cap;
  const u64 *prev_total;      /* total_out of the subblock group before this one (nullptr: this is the first) */
  i64 batch_base, region_len; i32 batch_is_final; u32 slack;
  u32 span_bytes;             /* dynamic shared memory available for record spans              */
  u32 max_nf;                 /* title fields (sizes the numeric-value table behind the span)  */
  u32 pk_bytes;               /* shared memory behind the span for the packed quality code tables */
  u32 enc_stage;              /* bytes of one warp's stage buffer in the encoder kernels (a 32-record block) */
  /* single-walk encoder (phy_encode.cuh) */
  u32 *tmp; u64 tmp_cap;      /* temporary buffer of the tasks' runs (words) */
  u64 *tmp_used;              /* words handed out so far in this batch */
  FastGeom fg;
  u32 ts;                     /* bytes of one lane's title slot in k_enc_title (a multiple of 16) */
  u32 qd_nbuf, qd_stage;      /* stage buffers per warp of k_enc_qd and their size */
  u32 sq_rows, sq_nbuf, sq_stage; /* k_seqstat: rows of the CTA's private quality table, stage buffers per warp and their size */
};

/* character classes of the title tokeniser (fill_char_lut), uploaded once per context */
__device__ u8 g_char_lut[256];
__device__ __forceinline__ void load_lut(u8 *lut) {
  for (u32 i = threadIdx.x; i < 64; i += blockDim.x) ((u32 *)lut)[i] = ((const u32 *)g_char_lut)[i];
}
/* ambiguity transfer as one lookup (phyNGSC.cpp:184-206, 575-580): 0 for A/C/G/T and for bytes without an ambiguity
 * code, else 79 + 8 * amb_code(c), so that the transferred quality byte is q + g_xq_lut[c] */
__device__ u8 g_xq_lut[256];
__device__ __forceinline__ void load_xq(u8 *xq) {
  for (u32 i = threadIdx.x; i < 64; i += blockDim.x) ((u32 *)xq)[i] = ((const u32 *)g_xq_lut)[i];
}

/* ---------------------------------------------------------------------------------------------- */
template <int NW>
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *warp_sums /*[NW]*/, u32 &total) {
  u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= (u32)o) x += y; }
  if (lane == 31) warp_sums[w] = x;
  __syncthreads();
  u32 base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) { u32 s = warp_sums[k]; if ((u32)k < w) base += s; tot += s; }
  total = tot;
  __syncthreads();
  return base + x - v;
}
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32 *warp_sums /*[8]*/, u32 &total) { return block_excl_scan<8>(v, warp_sums, total); }

__device__ __forceinline__ u32 nl_nibble(u32 w) {
  const u32 x = w ^ 0x0A0A0A0Au;
  const u32 t = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); /* bit 7 of a byte set iff the byte was '\n' */
  return (t * 0x00204081u) >> 28;
}
/* 64-bit newline mask (bit i = byte p + i is '\n') of this thread's 64 bytes [p, p+64) restricted to [lo, hi) */
__device__ __forceinline__ uint2 nl_mask64(const u8 *in, u32 p, u32 lo, u32 hi) {
  uint2 m = make_uint2(0u, 0u);
  if (p >= hi || p + 64 <= lo) return m;
  if (p >= lo && p + 64 <= hi) {
    const uint4 *q = (const uint4 *)(in + p);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
    m.x = nl_nibble(v0.x) | nl_nibble(v0.y) << 4 | nl_nibble(v0.z) << 8 | nl_nibble(v0.w) << 12 | nl_nibble(v1.x) << 16 | nl_nibble(v1.y) << 20 |
          nl_nibble(v1.z) << 24 | nl_nibble(v1.w) << 28;
    m.y = nl_nibble(v2.x) | nl_nibble(v2.y) << 4 | nl_nibble(v2.z) << 8 | nl_nibble(v2.w) << 12 | nl_nibble(v3.x) << 16 | nl_nibble(v3.y) << 20 |
          nl_nibble(v3.z) << 24 | nl_nibble(v3.w) << 28;
  } else {
    for (u32 i = (p < lo ? lo : p); i < p + 64 && i < hi; ++i)
      if (in[i] == '\n') { if (i - p < 32) m.x |= 1u << (i - p); else m.y |= 1u << (i - p - 32); }
  }
  return m;
}

/* (a) record splitter, pass 1: newline mask of every 64-byte piece (kept for pass 2: one bit per input byte instead
 * of a second read of the input) and the newline count per 16 KiB tile */
__global__ void __launch_bounds__(NLT) k_nl_count(Dev d) {
  __shared__ u32 ws[NLT / 32];
  u32 t = blockIdx.x;
  u32 p = t * TILE + threadIdx.x * 64;
  const uint2 m = nl_mask64(d.in, p, d.start_pos, d.len);
  d.nl_mask[(size_t)t * NLT + threadIdx.x] = m;
  u32 n = __popc(m.x) + __popc(m.y);
  n = __reduce_add_sync(0xFFFFFFFFu, n);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x < 32) {
    n = __reduce_add_sync(0xFFFFFFFFu, threadIdx.x < NLT / 32 ? ws[threadIdx.x] : 0u);
    if (threadIdx.x == 0) { d.tile_cnt[t] = n; if (n) atomicAdd(&d.tile_off[t / SUPER], n); } /* tile_off: supertile totals, zeroed by the host */
  }
}

/* exclusive scan over the supertile totals (single CTA) */
__global__ void __launch_bounds__(1024) k_nl_scan(Dev d) {
  __shared__ u32 ws[32];
  const u32 nsuper = (d.ntiles + SUPER - 1) / SUPER;
  u32 carry = 0;
  for (u32 base = 0; base < nsuper; base += 1024) {
    const u32 i = base + threadIdx.x;
    const u32 v = i < nsuper ? d.tile_off[i] : 0u;
    u32 tot;
    const u32 ex = block_excl_scan<32>(v, ws, tot);
    if (i < nsuper) d.tile_off[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) {
    const u32 total = carry;
    d.hdr->NL = total; d.hdr->NR = total / 4;
    d.hdr->status = 0; d.hdr->S = 0; d.hdr->max_chunks = 0; d.hdr->max_span = 0; d.hdr->max_nf = 0; d.hdr->max_pk_bytes = 0; d.hdr->max_len = 0; d.hdr->max_span64 = 0; d.hdr->max_span32 = 0; d.hdr->total_out = 0; d.hdr->pad_q = 0; d.hdr->max_rec = 0; d.hdr->max_tlen = 0; d.hdr->max_qcode = 0; d.hdr->pad_hdr = 0;
    if (total / 4 + 1 > d.maxrec) d.hdr->status = E_CAPACITY;
    d.rstart[0] = d.start_pos;
  }
}

/* (a) record splitter, pass 2: line l = 4r+k ends at the l-th newline; k=0 title, 1 sequence, 3 quality.
 * Each thread owns 64 bytes.  It first builds the 64-bit mask of its newline bytes with word operations (exact
 * zero-byte test of w ^ 0x0A0A0A0A, the four flag bits gathered by a multiply), all lanes converged; only then are
 * the set bits visited -- as many iterations as the thread has newlines, two on average. */
__device__ __forceinline__ void nl_put(const Dev &d, u32 l, u32 pos) {
  const u32 r = l >> 2, k = l & 3;
  u32 *dst = k == 0 ? d.te + r : k == 1 ? d.se + r : d.rstart + r + 1;
  if (k != 2) *dst = pos + (k == 3 ? 1u : 0u);
}
__global__ void __launch_bounds__(NLT) k_nl_emit(Dev d) {
  __shared__ u32 ws[NLT / 32];
  __shared__ u32 tile_base;
  if (d.hdr->status) return;
  const u32 t = blockIdx.x;
  const u32 p = t * TILE + threadIdx.x * 64;
  const uint2 m = d.nl_mask[(size_t)t * NLT + threadIdx.x]; /* pass 1 left it there */
  u32 mlo = m.x, mhi = m.y;
  const u32 n = __popc(mlo) + __popc(mhi);
  if (threadIdx.x < 32) { /* newlines before this tile: supertile base + the earlier tiles of the supertile */
    const u32 t0 = t & ~(u32)(SUPER - 1);
    u32 a = 0;
    for (u32 u = t0 + threadIdx.x; u < t; u += 32) a += d.tile_cnt[u];
    a = __reduce_add_sync(0xFFFFFFFFu, a);
    if (threadIdx.x == 0) tile_base = d.tile_off[t / SUPER] + a;
  }
  u32 tot;
  u32 l = block_excl_scan<NLT / 32>(n, ws, tot); /* its barriers also publish tile_base */
  l += tile_base;
  while (mlo) { const u32 bit = __ffs(mlo) - 1; mlo &= mlo - 1; nl_put(d, l++, p + bit); }
  while (mhi) { const u32 bit = __ffs(mhi) - 1; mhi &= mhi - 1; nl_put(d, l++, p + 32 + bit); }
}

/* ---- window chaining: one warp walks the rank's windows (phyNGSC.cpp:168-250, 744-755) -------------- */
/* first j in [lo, hi) with a[j] >= target (hi if none); a is ascending.  32 probes per step: three windows
 * of growing stride centred on `guess` (the caller's interpolation), then 32-ary narrowing. */
__device__ __forceinline__ void wlb_probe(const u32 *a, i64 target, u32 base, u32 step, u32 &L, u32 &H) {
  u32 lane = threadIdx.x & 31;
  u64 idx = (u64)base + (u64)lane * step;
  bool ge = idx < (u64)H ? ((i64)a[idx] >= target) : true;
  u32 bal = __ballot_sync(0xFFFFFFFFu, ge);
  if (bal == 0) { u64 nl = (u64)base + 31ull * step + 1; if (nl > L) L = (u32)(nl < H ? nl : H); return; }
  u32 k = __ffs(bal) - 1;
  u64 hit = (u64)base + (u64)k * step;
  u32 nL = L;
  if (k > 0) { u64 x = (u64)base + (u64)(k - 1) * step + 1; if (x > nL) nL = (u32)x; }
  if (hit < (u64)H) H = (u32)hit;
  L = nL < H ? nL : H;
}
__device__ u32 warp_lower_bound(const u32 *a, u32 lo, u32 hi, i64 target, i64 guess) {
  u32 L = lo, H = hi;
  const u32 strides[3] = {1, 8, 64};
  for (int t = 0; t < 3 && L < H; ++t) {
    i64 bs = guess - 16 * (i64)strides[t];
    u32 base = bs < (i64)L ? L : (bs > (i64)H - 1 ? H - 1 : (u32)bs);
    wlb_probe(a, target, base, strides[t], L, H);
  }
  while (L < H) wlb_probe(a, target, L, (H - L + 31) / 32, L, H);
  return L;
}

__global__ void __launch_bounds__(32) k_plan(Dev d) {
  PlanState st = *d.plan_state;
  BatchHdr *H = d.hdr;
  u32 lane = threadIdx.x;
  if (H->status || st.done || st.status) { if (lane == 0) { H->S = 0; H->next_pos = (u64)st.bytes_read; } return; }
  const u32 NR = H->NR, NL = H->NL;
  u32 F = 0, S = 0, chunk_base = 0, max_chunks = 0;
  i64 avg_n = 1;
  float dens = 1.0f / 128.0f;
  /* Software pipeline over the windows: the probe of window k + 1 (its position is predictable: one window's worth of records
   * behind this window's) is loaded into registers while window k is resolved, and the title newline of window k + 1's second
   * record is taken from window k's probe, so that in the common case no memory latency is left on the chain. */
  u32 nte = 0xFFFFFFFFu, nte_idx = 0xFFFFFFFFu; /* te[nte_idx], kept from the previous window's probe */
  u32 npbase = 0xFFFFFFFFu;                     /* base index of the pre-loaded probe (0xFFFFFFFF: none) */
  uint4 nprs = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), npte = nprs;
  while (!st.done && S < d.max_sb) {
    i64 ws = st.bytes_read - d.batch_base; /* batch-relative window start */
    if (!d.batch_is_final && ws + st.rsize + (i64)d.slack > (i64)d.len) break;
    i64 readable = (i64)d.len - ws;
    i64 lim = st.rsize < readable ? st.rsize : readable;
    SbPlan P;
    P.win_off = (u64)st.bytes_read; P.win_len = (u64)st.rsize; P.rec_start = st.rec_start; P.overlap = st.overlap;
    P.first_rec = F; P.n_records = 0; P.warnings = 0; P.status = 0; P.bytes_consumed = 0; P.chunk_base = chunk_base; P.pad = 0;
    if (F >= NR) { st.status = E_MALFORMED; break; }
    i64 target = ws + st.rsize - st.overlap, size_lim = ws + lim;
    u32 last = F, rs_next = 0xFFFFFFFFu; /* rs_next: rstart[last + 1] when the probe already holds it */
    bool capped = false;
    const i64 guess = (i64)F + (i64)((float)(target - (ws + st.rec_start)) * dens); /* records per byte of the previous window; the probe absorbs the rounding */
    /* With no_threads > 1 the reference's stop rule only runs in the last thread's slice of the window, from that
     * slice's second record on (phyNGSC.cpp:261-266, 303, 315): every record whose title ends before the slice is taken,
     * and Fs -- the first record whose title ends inside it -- plays the part record F plays with one thread. */
    u32 Fs = F;
    if (st.threads > 1) {
      const i64 bT = ws + (i64)(st.threads - 1) * st.rsize / (i64)st.threads;
      const u32 te_hi = (u32)min((u64)NR, ((u64)NL + 3) / 4); /* te[] is valid below this index */
      if (F < te_hi && (i64)d.te[F] < bT) {
        const i64 g2 = (i64)F + (i64)((float)(bT - (ws + st.rec_start)) * dens);
        Fs = warp_lower_bound(d.te, F + 1, te_hi, bT, g2);
        if (Fs >= te_hi || (i64)d.te[Fs] >= size_lim) Fs = Fs - 1; /* no title ends in the last slice: the earlier threads' records are all there is */
      }
    }
    /* One round trip in the common case: the title newline of the second record, and a 128-wide probe (four table
     * entries per lane) of the record table around the interpolated position of the last record, are loaded together. */
    u32 pbase;
    uint4 prs = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), pte = prs;
    bool have = npbase != 0xFFFFFFFFu && npbase >= Fs + 2;
    if (have) { /* loaded while the previous window was resolved: usable when it brackets the target */
      pbase = npbase; prs = nprs; pte = npte;
      const i64 lo = (i64)__shfl_sync(0xFFFFFFFFu, prs.x, 0), hi = pbase + 127u > NR ? (i64)0xFFFFFFFFu : (i64)__shfl_sync(0xFFFFFFFFu, prs.w, 31);
      have = lo < target && hi >= target;
    }
    if (!have) {
      i64 gb = guess - 64; pbase = gb < (i64)Fs + 2 ? Fs + 2 : (u32)gb; pbase = (pbase + 3u) & ~3u;
      const u32 pi = pbase + 4 * lane;
      prs = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); pte = prs;
      if (pi <= NR) { prs = *(const uint4 *)(d.rstart + pi); pte = *(const uint4 *)(d.te + pi); } /* the tables have four entries of slack behind the last record */
    }
    const u32 pidx = pbase + 4 * lane;
    if (pidx > NR) prs.x = 0xFFFFFFFFu;
    if (pidx + 1 > NR) prs.y = 0xFFFFFFFFu;
    if (pidx + 2 > NR) prs.z = 0xFFFFFFFFu;
    if (pidx + 3 > NR) prs.w = 0xFFFFFFFFu;
    { /* the next window's probe: about as many records behind this window's last as this window holds */
      const i64 gn = 2 * guess - (i64)F - 64;
      npbase = 0xFFFFFFFFu;
      if (S > 0 && gn > (i64)pbase && gn + 4 * 31 + 3 <= (i64)d.maxrec) {
        npbase = ((u32)gn + 3u) & ~3u;
        nprs = *(const uint4 *)(d.rstart + npbase + 4 * lane); npte = *(const uint4 *)(d.te + npbase + 4 * lane);
      }
      /* and the probe region of the window after it into L2 */
      const u64 nidx = (u64)pidx + 2 * (u64)avg_n;
      if (S > 0 && nidx + 3 <= (u64)d.maxrec) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.rstart + nidx));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.te + nidx));
      }
    }
    const u32 te_f1 = Fs + 1 == nte_idx ? nte : 4ull * (Fs + 1) < NL ? d.te[Fs + 1] : 0xFFFFFFFFu;
    last = Fs;
    if ((i64)te_f1 < size_lim) {
      u32 m;
      const u32 kl = (i64)prs.x >= target ? 0u : (i64)prs.y >= target ? 1u : (i64)prs.z >= target ? 2u : (i64)prs.w >= target ? 3u : 4u;
      const u32 bal = __ballot_sync(0xFFFFFFFFu, kl < 4u);
      const u32 hl = bal ? __ffs(bal) - 1 : 0u;
      const u32 hk = __shfl_sync(0xFFFFFFFFu, kl, hl);
      const u32 cand = pbase + 4 * hl + hk;
      const bool hit = bal != 0 && (cand > pbase || pbase == Fs + 2) && cand <= NR;
      u32 te_last = 0xFFFFFFFFu;
      if (hit) {
        m = cand;
        const u32 rsel = hk == 0 ? prs.x : hk == 1 ? prs.y : hk == 2 ? prs.z : prs.w;
        rs_next = __shfl_sync(0xFFFFFFFFu, rsel, hl);
        if (m > pbase) { /* te[m - 1] sits in the probe as well */
          const u32 q = m - 1 - pbase, ql = q >> 2, qk = q & 3u;
          const u32 tsel = qk == 0 ? pte.x : qk == 1 ? pte.y : qk == 2 ? pte.z : pte.w;
          te_last = 4ull * (m - 1) < NL ? __shfl_sync(0xFFFFFFFFu, tsel, ql) : 0xFFFFFFFFu;
        } else te_last = te_f1; /* m == Fs + 2 */
      }
      else m = warp_lower_bound(d.rstart, Fs + 2, NR + 1, target, guess);
      last = m - 1;
      if (last > F + st.record_cap) { last = F + st.record_cap; capped = true; te_last = 0xFFFFFFFFu; rs_next = 0xFFFFFFFFu; }
      if (te_last == 0xFFFFFFFFu && 4ull * last < NL) te_last = d.te[last];
      while (last > Fs && !(4ull * last < NL && (i64)te_last < size_lim)) {
        --last; capped = false; rs_next = 0xFFFFFFFFu;
        te_last = 4ull * last < NL ? d.te[last] : 0xFFFFFFFFu;
      }
      if (last >= NR) { st.status = E_MALFORMED; break; }
    }
    P.n_records = last - F + 1;
    P.warnings = capped ? 1u : 0u;
    { /* the title newline of the next window's second record usually sits in this window's probe */
      const u32 q = last + 2 - pbase;
      nte_idx = 0xFFFFFFFFu;
      if (last + 2 >= pbase && q < 128u) {
        const u32 qk = q & 3u, tsel = qk == 0 ? pte.x : qk == 1 ? pte.y : qk == 2 ? pte.z : pte.w;
        nte = 4ull * (last + 2) < NL ? __shfl_sync(0xFFFFFFFFu, tsel, q >> 2) : 0xFFFFFFFFu;
        nte_idx = last + 2;
      }
    }
    if (rs_next == 0xFFFFFFFFu) rs_next = d.rstart[last + 1];
    P.bytes_consumed = (u64)((i64)rs_next - ws);
    if (lane == 0) d.plans[S] = P;
    __syncwarp();
    u32 nch = (P.n_records + CH - 1) / CH;
    chunk_base += nch;
    max_chunks = max(max_chunks, nch);
    avg_n = (i64)P.n_records; dens = (float)P.n_records / (float)max((i64)1, (i64)P.bytes_consumed); /* the previous window predicts best: record sizes drift along a file */
    ++S; F = last + 1;
    /* phyNGSC.cpp:745-755 */
    st.bytes_read += (i64)P.bytes_consumed;
    if (st.bytes_read + st.rsize > st.wr_len - 1) { if (st.is_last) st.overlap = 0; st.rsize = st.wr_len - 1 - st.bytes_read; }
    st.rec_start = 0;
    st.done = st.bytes_read >= st.region;
    st.n_subblocks_total++;
  }
  if (lane == 0) {
    H->S = S; H->max_chunks = max_chunks;
    H->next_pos = (u64)st.bytes_read;
    if (st.status) H->status = st.status;
    *d.plan_state = st;
  }
}

/* Exact shared-memory needs of the per-record kernels: the widest 128-record span (plus the record before
 * it) and the largest field count over all subblocks of the batch. */
__global__ void __launch_bounds__(256) k_spanmax(Dev d) {
  const u32 s = blockIdx.y;
  if (s >= d.hdr->S) return; /* launched for the context's capacity: the host learns S only afterwards */
  const SbPlan P = d.plans[s];
  if (P.status) return;
  u32 mx = 0, m64 = 0, m32 = 0, ml = 0;
  for (u32 g = blockIdx.x * 256 + threadIdx.x; g * 32 < P.n_records; g += gridDim.x * 256) { /* 32-record groups */
    const u32 i = g * 32, r0 = P.first_rec + i, n = P.n_records;
    const u32 lo = d.rstart[r0] & ~15u;
    m32 = max(m32, d.rstart[P.first_rec + min(i + 32, n)] - lo);
    if ((g & 1u) == 0) m64 = max(m64, d.rstart[P.first_rec + min(i + 64, n)] - lo);
    if ((g & 3u) == 0) mx = max(mx, d.rstart[P.first_rec + min(i + CH, n)] - lo);
  }
  u32 mr = 0, mt = 0;
  for (u32 i = blockIdx.x * 256 + threadIdx.x; i < P.n_records; i += gridDim.x * 256) {
    const u32 r = P.first_rec + i, rs = d.rstart[r], te = d.te[r];
    ml = max(ml, d.se[r] - te - 1); mr = max(mr, d.rstart[r + 1] - rs); mt = max(mt, te + 1 - rs);
  }
  mx = __reduce_max_sync(0xFFFFFFFFu, mx); m64 = __reduce_max_sync(0xFFFFFFFFu, m64); m32 = __reduce_max_sync(0xFFFFFFFFu, m32);
  ml = __reduce_max_sync(0xFFFFFFFFu, ml); mr = __reduce_max_sync(0xFFFFFFFFu, mr); mt = __reduce_max_sync(0xFFFFFFFFu, mt);
  if ((threadIdx.x & 31) == 0 && mr) { atomicMax(&d.hdr->max_rec, mr); atomicMax(&d.hdr->max_tlen, mt); }
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&d.hdr->max_span, mx);
  if ((threadIdx.x & 31) == 0 && m64) atomicMax(&d.hdr->max_span64, m64);
  if ((threadIdx.x & 31) == 0 && m32) atomicMax(&d.hdr->max_span32, m32);
  if ((threadIdx.x & 31) == 0 && ml) atomicMax(&d.hdr->max_len, ml);
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicMax(&d.hdr->max_nf, count_seps(d.in, d.rstart[P.first_rec], d.te[P.first_rec]));
}

/* Dynamic shared memory of the per-record kernels: [span_bytes: staged records][vals: nf x CH numeric values] */
__device__ __forceinline__ u32 *vals_area(uint4 *dyn, u32 span_bytes) { return (u32 *)((u8 *)dyn + span_bytes); }

/* ---- bulk-copy span pipeline ------------------------------------------------------------------------------------ */
/* A CTA that walks several 128-record chunks streams their bytes into shared memory with the bulk-copy engine
 * (cp.async.bulk, one request per chunk issued by one thread, completion counted on an mbarrier), double-buffered so
 * that the next chunk arrives while the current one is processed.  No thread spends instructions on the copy. */
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  u32 ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 20)) __trap(); /* a lost copy must not hang the GPU */
  } while (!ok);
}
/* one thread: request bytes [lo & ~15, hi) of the batch (rounded up to 16) into `dst` (shared address, 16-byte aligned) */
__device__ __forceinline__ void span_request(const u8 *in, u32 lo, u32 hi, u32 dst, u32 bar) {
  const u32 alo = lo & ~15u, n = (hi - alo + 15u) & ~15u;
  mbar_expect_tx(bar, n);
  bulk_g2s(dst, in + alo, n, bar);
}

/* One stage buffer fed by the bulk-copy engine: the CTA's next chunk is requested as soon as every thread has left
 * the current one (chunks of an encoder CTA are short; more resident CTAs hide the copy better than a second buffer). */
struct ChunkStage {
  u32 buf_a, bar_a, phase;
  const u8 *buf;
  __device__ __forceinline__ void init(void *smem, u64 *bar) { /* followed by a __syncthreads() of the caller */
    buf = (const u8 *)smem; buf_a = (u32)__cvta_generic_to_shared(smem); bar_a = (u32)__cvta_generic_to_shared(bar); phase = 0;
    if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_fence_init(); }
  }
  __device__ __forceinline__ void request(const u8 *in, u32 lo, u32 hi) const { span_request(in, lo, hi, buf_a, bar_a); } /* one thread */
  __device__ __forceinline__ const u8 *wait(u32 lo) { mbar_wait(bar_a, phase); phase ^= 1u; return buf - (lo & ~15u); } /* p[pos] valid for the chunk's positions */
};

/* four bytes at any shared-memory address (may read up to three bytes past them) */
__device__ __forceinline__ u32 ld4u(const u8 *p) {
  const u32 a = (u32)(size_t)p & 3u;
  const u32 *w = (const u32 *)(p - a);
  return __funnelshift_r(w[0], w[1], 8 * a);
}
/* n >= 1 bytes equal?  Both sides in shared memory, any alignment (may read up to three bytes past either side).
 * Each side keeps the previous aligned word, so a step of four bytes is two loads, two funnel shifts and one LOP3; the loop
 * is kept rolled (two steps per trip): the runs are short and the set-up of a deeper unrolling costs more than it saves. */
__device__ __forceinline__ u32 lds_w(u32 addr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ bool eq_bytes(const u8 *x, const u8 *y, u32 n) {
  u32 xa = (u32)__cvta_generic_to_shared(x), ya = (u32)__cvta_generic_to_shared(y);
  const u32 sx = (xa & 3u) * 8u, sy = (ya & 3u) * 8u;
  xa &= ~3u; ya &= ~3u;
  u32 x0 = lds_w(xa), y0 = lds_w(ya), diff = 0;
#pragma unroll 2
  for (; n >= 4; n -= 4) {
    xa += 4; ya += 4;
    const u32 x1 = lds_w(xa), y1 = lds_w(ya);
    diff |= __funnelshift_r(x0, x1, sx) ^ __funnelshift_r(y0, y1, sy);
    x0 = x1; y0 = y1;
  }
  if (n) {
    const u32 x1 = lds_w(xa + 4), y1 = lds_w(ya + 4);
    diff |= (__funnelshift_r(x0, x1, sx) ^ __funnelshift_r(y0, y1, sy)) & (0xFFFFFFFFu >> (8 * (4 - n)));
  }
  return diff == 0;
}
/* bit p set iff x[p] != y[p], p < n <= 32 (both sides in shared memory, any alignment; may read up to three bytes past either side) */
__device__ __forceinline__ u32 neq_mask(const u8 *x, const u8 *y, u32 n) {
  u32 xa = (u32)__cvta_generic_to_shared(x), ya = (u32)__cvta_generic_to_shared(y);
  const u32 sx = (xa & 3u) * 8u, sy = (ya & 3u) * 8u;
  xa &= ~3u; ya &= ~3u;
  u32 x0 = lds_w(xa), y0 = lds_w(ya), mm = 0;
#pragma unroll 1
  for (u32 p = 0; p < n; p += 4) {
    xa += 4; ya += 4;
    const u32 x1 = lds_w(xa), y1 = lds_w(ya);
    const u32 v = __funnelshift_r(x0, x1, sx) ^ __funnelshift_r(y0, y1, sy);
    const u32 t = (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u; /* bit 7 of a byte: the byte is not zero */
    mm |= ((t * 0x00204081u) >> 28) << p;
    x0 = x1; y0 = y1;
  }
  return n >= 32 ? mm : mm & ((1u << n) - 1u);
}

/* ---- title line staging (k_stat1, phy_title.cuh) ----------------------------------------------------------------- */
/* per-thread staging of title lines: 16-byte cp.async pieces from the aligned address below the line's first byte into the
 * thread's slot (a stage holds one slot of `ts` bytes per thread) */
__device__ __forceinline__ void cp_async16(u32 dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ bool stage_title_line(const u8 *in, u32 slot_a, u32 ts, u32 rs, u32 te) {
  const u32 a0 = rs & ~15u, n = (te + 1 - a0 + 15u) >> 4;
  if (n * 16u > ts) return false;
  for (u32 k = 0; k < n; ++k) cp_async16(slot_a + 16 * k, in + a0 + 16 * k);
  return true;
}

__device__ __forceinline__ u32 *raw_table(const Dev &d, u32 s) { return d.arena + (size_t)(s + 1) * d.arena_words - RAW_WORDS; }

/* ---- classify + zero ------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(32) k_classify(Dev d) {
  u32 s = blockIdx.x;
  const SbPlan P = d.plans[s];
  SbClass &C = d.cls[s];
  if (!P.status && !d.acc[s].status) { /* quality alphabet = the bytes counted by k_qhist (row 0 of the raw table holds the totals) */
    u32 *raw = raw_table(d, s);
    const u32 rows = d.acc[s].max_qlen;
    u32 tot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (u32 p = 1; p <= rows; ++p) {
#pragma unroll
      for (u32 k = 0; k < 8; ++k) tot[k] += raw[p * 256 + k * 32 + threadIdx.x];
    }
#pragma unroll
    for (u32 k = 0; k < 8; ++k) {
      raw[k * 32 + threadIdx.x] = tot[k]; /* row 0: quality_stats[0] (tasks.cpp:281) */
      const u32 m = __ballot_sync(0xFFFFFFFFu, tot[k] != 0);
      if (threadIdx.x == 0) d.acc[s].qpresent[k] = m;
    }
    if (threadIdx.x == 0 && (d.acc[s].qpresent[0] & 1u)) d.acc[s].status = E_UNSUPPORTED; /* NUL quality byte */
    __syncwarp();
  }
  if (threadIdx.x != 0) return;
  if (P.status) { C.status = P.status; C.R = P.n_records; C.payload_len = 0; return; }
  u32 ts0 = d.rstart[P.first_rec], te0 = d.te[P.first_rec];
  classify_subblock(d.in, g_char_lut, d.acc[s], P.n_records, ts0, te0, d.arena + (size_t)s * d.arena_words, d.arena_words - RAW_WORDS, C);
  C.payload_len = 0;
  if (!C.status) atomicMax(&d.hdr->max_pk_bytes, (C.max_qlen + 1) * C.nq * 2u);
}

/* Clears the title / DNA histograms and fills the per-position quality frequency tables (tasks.cpp:260-286:
 * quality_stats[p][qua_code[q]]) from the raw table: column q of row p becomes entry qua_code[q] of table p. */
__global__ void __launch_bounds__(256) k_zero_hist(Dev d) {
  u32 s = blockIdx.y;
  const SbClass &C = d.cls[s];
  if (C.status) return;
  u32 *a = d.arena + (size_t)s * d.arena_words;
  const u32 *raw = raw_table(d, s);
  const u32 nq = C.nq, nqs = (C.max_qlen + 1) * nq; /* the quality tables are the first words of the zeroed range */
  for (u32 i = C.zero_begin + blockIdx.x * 256 + threadIdx.x; i < C.zero_end; i += gridDim.x * 256) {
    const u32 k = i - C.qstat_off;
    a[i] = k < nqs ? raw[(k / nq) * 256 + C.quals[k % nq]] : 0u;
  }
}

__device__ __forceinline__ u32 lds_u8(u32 addr) { u16 v; asm volatile("ld.shared.u8 %0, [%1];" : "=h"(v) : "r"(addr)); return v; }

/* ---- shared pieces of the per-record title kernels ------------------------------------------------------------ */

/* previous record's numeric value = the neighbouring lane's (lane = record of the 32-record block).  Lane 0
 * receives garbage, which is never used: the first record of a block is coded raw (tasks.cpp:455-458). */
struct PrevShfl {
  __device__ __forceinline__ i32 operator()(u32, i32 v) const { return __shfl_up_sync(0xFFFFFFFFu, v, 1); }
};

/* ---- stat2: numeric / char histograms and 32-record block descriptors (tasks.cpp:64-93, 127-182) --------------- */
__device__ __forceinline__ void warp_hist_add(u32 *hist, u32 idx, bool on) {
  u32 key = on ? idx : 0xFFFFFFFFu;
  u32 m = __match_any_sync(0xFFFFFFFFu, key);
  if (on && (u32)(__ffs(m) - 1) == (threadIdx.x & 31)) atomicAdd(hist + idx, (u32)__popc(m));
}

constexpr int CSLOTS = 8; /* per-position char tables whose histogram a CTA keeps in shared memory */
#ifndef PHY_S2G
#define PHY_S2G 8
#endif
constexpr int S2G = PHY_S2G;    /* 128-record chunks per k_dnacount CTA */

/* field classes and the list of non-constant fields of a subblock, copied to shared memory once per CTA */
struct TitleTabs { FieldClass fc[MAXF]; u16 ncskip[MAXF]; u8 ncf[MAXF]; };
__device__ __forceinline__ void load_title_tabs(const SbClass &C, TitleTabs &T) {
  for (u32 i = threadIdx.x; i < C.nf * (sizeof(FieldClass) / 4); i += blockDim.x) ((u32 *)T.fc)[i] = ((const u32 *)C.f)[i];
  for (u32 i = threadIdx.x; i < MAXF / 2; i += blockDim.x) ((u32 *)T.ncskip)[i] = ((const u32 *)C.ncskip)[i];
  for (u32 i = threadIdx.x; i < MAXF / 4; i += blockDim.x) ((u32 *)T.ncf)[i] = ((const u32 *)C.ncf)[i];
}

/* Exact DNA symbol counts (sym_stats, tasks.cpp:233-236): only needed -- and only run -- when more than four symbols
 * force Huffman-coded DNA.  A CTA walks S2G consecutive chunks of one subblock, thread = record, spans staged by the
 * bulk-copy engine; four bases per step like k_seqstat, A/C/G/T by popcount of their one-hot bytes.
 * dynamic shared memory: [span_bytes stage buffer] */
__global__ void __launch_bounds__(CH) k_dnacount(Dev d) {
  extern __shared__ uint4 dyn_smem[];
  __shared__ __align__(8) u64 bar;
  __shared__ u32 c_lo[S2G + 1];
  __shared__ u32 dnah[256]; /* by symbol code */
  const u32 s = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const SbClass &C = d.cls[s];
  if (C.status || C.plain) return;
  const u32 c0 = blockIdx.x * S2G, c1 = min(c0 + S2G, C.nchunk);
  if (c0 >= C.nchunk) return;
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const u32 R = C.R;
  if (tid <= c1 - c0) c_lo[tid] = d.rstart[P.first_rec + min((c0 + tid) * CH, R)];
  for (u32 i = tid; i < 256; i += CH) dnah[i] = 0;
  ChunkStage stage; stage.init(dyn_smem, &bar);
  __syncthreads();
  {
    bool fits = true;
    for (u32 k = 0; k < c1 - c0; ++k) fits = fits && c_lo[k + 1] - (c_lo[k] & ~15u) + 16 <= d.span_bytes;
    if (!fits) { if (tid == 0) atomicMin(&d.cls[s].status, (i32)E_UNSUPPORTED); return; }
  }
  if (tid == 0) stage.request(d.in, c_lo[0], c_lo[1]);
  for (u32 c = c0; c < c1; ++c) {
    const u32 nrec = min((u32)CH, R - c * CH);
    const bool active = tid < nrec;
    u32 my_te = 0, my_se = 0, my_kx = 0;
    if (active) { const u32 r = P.first_rec + c * CH + tid; my_te = d.te[r]; my_se = d.se[r]; my_kx = d.kx[r]; }
    const u8 *b = stage.wait(c_lo[c - c0]);
    u32 nA = 0, nC = 0, nT = 0, nG = 0;
    if (active) {
      const u8 *sp = b + my_te + 1;
      const u32 L = my_se - my_te - 1;
      const bool xfer = my_kx >> 15;
      const u32 a = (u32)(size_t)sp & 3u;
      const u32 *wp = (const u32 *)(sp - a);
      u32 w0 = wp[0], j = 0;
      for (; j + 4 <= L; j += 4) {
        const u32 w1 = *++wp;
        const u32 v = __funnelshift_r(w0, w1, a * 8);
        w0 = w1;
        const u32 z = (v >> 1) & 0x03030303u;
        const u32 sel = __byte_perm(z | (z >> 4), 0, 0x4420);
        u32 oh = __byte_perm(0x08040201u, 0, sel);
        const u32 bad = __byte_perm(0x47544341u, 0, sel) ^ v;
        if (bad) { /* some of the four is not A/C/G/T: it counts under its own symbol unless the record's codes were transferred */
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if ((bad >> (8 * q)) & 0xFFu) { oh &= ~(0xFFu << (8 * q)); if (!xfer) atomicAdd(&dnah[C.sym_code[(v >> (8 * q)) & 0xFFu]], 1u); }
        }
        nA += __popc(oh & 0x01010101u); nC += __popc(oh & 0x02020202u); nT += __popc(oh & 0x04040404u); nG += __popc(oh & 0x08080808u);
      }
      for (; j < L; ++j) {
        const u8 ch = sp[j];
        if (ch == 'A') ++nA; else if (ch == 'C') ++nC; else if (ch == 'T') ++nT; else if (ch == 'G') ++nG;
        else if (!xfer) atomicAdd(&dnah[C.sym_code[ch]], 1u);
      }
    }
    nA = __reduce_add_sync(0xFFFFFFFFu, nA); nC = __reduce_add_sync(0xFFFFFFFFu, nC);
    nT = __reduce_add_sync(0xFFFFFFFFu, nT); nG = __reduce_add_sync(0xFFFFFFFFu, nG);
    if (lane == 0) {
      if (nA) atomicAdd(&dnah[C.sym_code['A']], nA);
      if (nC) atomicAdd(&dnah[C.sym_code['C']], nC);
      if (nT) atomicAdd(&dnah[C.sym_code['T']], nT);
      if (nG) atomicAdd(&dnah[C.sym_code['G']], nG);
    }
    __syncthreads(); /* the stage buffer is free again */
    if (tid == 0 && c + 1 < c1) stage.request(d.in, c_lo[c + 1 - c0], c_lo[c + 2 - c0]);
  }
  for (u32 i = tid; i < C.nsym; i += CH) if (dnah[i]) atomicAdd(arena + C.dnastat_off + i, dnah[i]);
}

/* ---- Huffman build: one warp per table ------------------------------------------------------------------------- */
struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

/* Tables of up to 64 symbols (the per-position quality tables, the DNA table: nearly all of them) are built with the small
 * scratch by CTAs of 8 warps, so that an SM's warp slots -- not its shared memory -- bound how many are under way at once (the
 * build is serial on lane 0 inside a table); the few larger ones (numeric tables up to 512 symbols, per-position character
 * tables of 256) by the first HUFF_LARGE CTAs of the same launch, one warp each with the large scratch. */
constexpr u32 HUFF_LARGE = 16, HUFF_WARPS = 8;
constexpr u32 HUFF_SMEM = sizeof(HuffScratch) > HUFF_WARPS * sizeof(HuffScratchSmall) ? sizeof(HuffScratch) : HUFF_WARPS * sizeof(HuffScratchSmall);

/* what follows a table's build: its blob length, its longest code, the packed copy of a quality table */
__device__ __forceinline__ void huff_finish(SbClass &C, u32 *arena, TableDesc *td, u32 t, const TableDesc &D, u32 blob, u32 lane) {
  if (lane == 0) td[t].tree_len = blob;
  __syncwarp();
  { /* longest code: bounds the staging of the single-walk encoder (phy_fast.cuh) */
    const u64 *cl = (const u64 *)(arena + D.cl_off);
    u32 ml = 0;
    for (u32 i = lane; i < D.n && blob; i += 32) ml = max(ml, (u32)(cl[i] >> 32));
    ml = __reduce_max_sync(0xFFFFFFFFu, ml);
    if (lane == 0) td[t].maxlen = ml;
  }
  if (t - C.tq0 <= C.max_qlen) { /* quality table: 16-bit copy (len << 12 | code, PK_ESC for the rare codes beyond 12 bits) for the shared-memory walkers */
    u16 *pk = (u16 *)(arena + C.qpk_off) + (size_t)(t - C.tq0) * D.n;
    const u64 *cl = (const u64 *)(arena + D.cl_off);
    bool esc = false;
    for (u32 i = lane; i < D.n && blob; i += 32) { u16 e; if (!qpack_entry(cl[i], e)) { e = (u16)PK_ESC; esc = true; } pk[i] = e; } /* longer than 12 bits: escape to the 64-bit entry */
    if (blob == 0) atomicOr(&C.qpk_bad, 1u);
    if (esc) atomicOr(&C.qpk_esc, 1u);
  }
}

__global__ void __launch_bounds__(HUFF_WARPS * 32) k_huff(Dev d) {
  extern __shared__ uint4 dyn_smem[];
  const u32 s = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SbClass &C = d.cls[s];
  if (C.status) return;
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  TableDesc *td = (TableDesc *)(arena + C.tabdesc_off);
  if (blockIdx.x < HUFF_LARGE) { /* one warp, large scratch: the tables of more than 64 symbols */
    if (w) return;
    HuffScratch &HS = *(HuffScratch *)dyn_smem;
    for (u32 t = blockIdx.x; t < C.ntab; t += HUFF_LARGE) {
      const TableDesc D = td[t];
      if (D.n <= HuffScratchSmall::CAP) continue;
      const u32 blob = huff_table(arena + D.freq_off, D.n, (u64 *)(arena + D.cl_off), (u8 *)(arena + D.tree_off), HS, lane, 32u, WarpSync());
      huff_finish(C, arena, td, t, D, blob, lane);
    }
    return;
  }
  HuffScratchSmall *HS = (HuffScratchSmall *)dyn_smem;
  for (u32 t = (blockIdx.x - HUFF_LARGE) * HUFF_WARPS + w; t < C.ntab; t += (gridDim.x - HUFF_LARGE) * HUFF_WARPS) {
    const TableDesc D = td[t];
    if (D.n > HuffScratchSmall::CAP) continue;
    const u32 blob = huff_table(arena + D.freq_off, D.n, (u64 *)(arena + D.cl_off), (u8 *)(arena + D.tree_off), HS[w], lane, 32u, WarpSync());
    huff_finish(C, arena, td, t, D, blob, lane);
  }
}

/* ---- per-record walkers of the encoder kernels -------------------------------------------------------------------- */
/* Shared-memory tables of one CTA: symbol maps, the ambiguity-transfer lookup and (when they fit and no code is longer
 * than 12 bits) the packed quality code tables of the subblock. */
struct WalkTabs {
  const u8 *qmap, *smap, *xq;  /* qua_code[256], sym_code[256], g_xq_lut */
  const u16 *pk; u32 nq;       /* packed quality tables (nullptr: read the 64-bit entries from the arena) */
  bool has_esc;                /* some packed entry is the escape */
  const u64 *qcl;              /* 64-bit quality entries, table-major */
  u32 dna_mode;                /* 0 Huffman, 1 two bits per base through smap, 2 two bits per base, alphabet exactly ACGT */
  const u64 *dcl;              /* DNA code table (Huffman mode) */
};

__device__ __forceinline__ void put_pk(CountSink &s, u32 e) { s.bits += e >> 12; }
__device__ __forceinline__ void put_pk2(CountSink &s, u32 e0, u32 e1) { s.bits += (e0 >> 12) + (e1 >> 12); }
template <class Sink> __device__ __forceinline__ void put_pk(Sink &s, u32 e) { s.put(e & 0xFFFu, e >> 12); }
template <class Sink> __device__ __forceinline__ void put_pk2(Sink &s, u32 e0, u32 e1) { /* two codes (<= 12 bits each) appended as one */
  const u32 l1 = e1 >> 12;
  s.put(((e0 & 0xFFFu) << l1) | (e1 & 0xFFFu), (e0 >> 12) + l1);
}

/* Quality codes of one record (tasks.cpp:609-619): position k uses table k+1.  Four symbols per step: their bytes,
 * symbol codes and table entries are loaded together (independent shared-memory loads), then appended pairwise.
 * The ambiguity transfer (phyNGSC.cpp:575-580) costs two more loads per symbol and is only compiled into the loop
 * of warps that hold such a record. */
/* one symbol whose packed entry may be the escape: its code then comes from the 64-bit table in the arena */
template <class Sink> __device__ __forceinline__ void put_pk_esc(Sink &s, u32 e, const u64 *full) {
  if (e >= PK_ESC) { const u64 f = *full; s.put((u32)f, (u32)(f >> 32)); }
  else put_pk(s, e);
}
/* ESC: the subblock has packed entries with the escape (a code longer than 12 bits somewhere); compiled separately so
 * that the common case carries neither the test nor its registers */
/* qp / sp point at read position pos0 (0 unless several lanes share a record); L symbols are coded from there */
template <bool ESC, class Sink>
__device__ __forceinline__ void quality_walk_pk(const u8 *qp, const u8 *sp, u32 L, bool xfer, const WalkTabs &T, Sink &s, u32 pos0 = 0) {
  const bool wx = __any_sync(__activemask(), xfer);
  const u32 xm = xfer ? 0xFFu : 0u;
  const u32 nq = T.nq;
  const u16 *row = T.pk + (size_t)(pos0 + 1) * nq;
  const u64 *frow = T.qcl + (size_t)(pos0 + 1) * nq; /* the same position in the 64-bit tables */
  u32 j = 0;
#define PHY_Q4_BODY                                                                                                          \
  const u32 c0 = T.qmap[q0], c1 = T.qmap[q1], c2 = T.qmap[q2], c3 = T.qmap[q3];                                            \
  const u32 e0 = row[c0], e1 = row[nq + c1], e2 = row[2 * nq + c2], e3 = row[3 * nq + c3];                                 \
  if (ESC && max(max(e0, e1), max(e2, e3)) >= PK_ESC) { /* rare: a code longer than 12 bits among the four */               \
    const u64 *f = frow + (size_t)j * nq;                                                                                   \
    put_pk_esc(s, e0, f + c0); put_pk_esc(s, e1, f + nq + c1); put_pk_esc(s, e2, f + 2 * nq + c2); put_pk_esc(s, e3, f + 3 * nq + c3); \
  } else {                                                                                                                  \
    put_pk2(s, e0, e1); put_pk2(s, e2, e3);                                                                                 \
  }
  if (!wx) {
    for (; j + 4 <= L; j += 4, row += 4 * nq) {
      const u32 q0 = qp[j], q1 = qp[j + 1], q2 = qp[j + 2], q3 = qp[j + 3];
      PHY_Q4_BODY
    }
  } else {
    for (; j + 4 <= L; j += 4, row += 4 * nq) {
      const u32 q0 = qp[j] + (T.xq[sp[j]] & xm), q1 = qp[j + 1] + (T.xq[sp[j + 1]] & xm);
      const u32 q2 = qp[j + 2] + (T.xq[sp[j + 2]] & xm), q3 = qp[j + 3] + (T.xq[sp[j + 3]] & xm);
      PHY_Q4_BODY
    }
  }
#undef PHY_Q4_BODY
  for (; j < L; ++j, row += nq) {
    const u32 c = T.qmap[qp[j] + (T.xq[sp[j]] & xm)], e = row[c];
    if (ESC) put_pk_esc(s, e, frow + (size_t)j * nq + c); else put_pk(s, e);
  }
}
template <class Sink>
__device__ __forceinline__ void quality_walk(const u8 *qp, const u8 *sp, u32 L, bool xfer, const WalkTabs &T, Sink &s, u32 pos0 = 0) {
  if (T.pk) {
    if (T.has_esc) quality_walk_pk<true>(qp, sp, L, xfer, T, s, pos0);
    else quality_walk_pk<false>(qp, sp, L, xfer, T, s, pos0);
    return;
  }
  const u32 xm = xfer ? 0xFFu : 0u;
  const u32 nq = T.nq;
  const u64 *row = T.qcl + (size_t)(pos0 + 1) * nq;
  for (u32 j = 0; j < L; ++j, row += nq) {
    const u64 e = row[T.qmap[qp[j] + (T.xq[sp[j]] & xm)]];
    s.put((u32)e, (u32)(e >> 32));
  }
}
template <class Sink>
__device__ __forceinline__ void dna_walk(const u8 *sp, u32 L, bool xfer, const WalkTabs &T, Sink &s) {
  if (T.dna_mode == 2) {
    /* alphabet exactly A, C, G, T: four bases per step, sixteen per append.  A byte that is not one of the four can only
     * be a transferred ambiguity code (anything else would be in the alphabet): it is left out (phyNGSC.cpp:549-588). */
    const u32 a = (u32)(size_t)sp & 3u;
    const u32 *wp = (const u32 *)(sp - a);
    u32 w0 = wp[0], acc = 0, cnt = 0, j = 0;
    for (; j + 4 <= L; j += 4) {
      const u32 w1 = *++wp;
      const u32 v = __funnelshift_r(w0, w1, a * 8);
      w0 = w1;
      const u32 zz = (v >> 1) & 0x03030303u;
      const u32 bad = __byte_perm(0x47544341u, 0, __byte_perm(zz | (zz >> 4), 0, 0x4420)) ^ v;
      const u32 z = zz ^ ((v >> 2) & 0x01010101u);
      if (bad == 0) {
        acc = (acc << 8) | ((z * 0x40100401u) >> 24);
        cnt += 4;
        if (cnt == 16) { s.put(acc, 32); acc = 0; cnt = 0; }
      } else { /* rare: what is pending goes out first, so that cnt stays a multiple of four in the steps above */
        s.put(acc, 2 * cnt); acc = 0; cnt = 0;
#pragma unroll
        for (u32 t = 0; t < 4; ++t)
          if (((bad >> (8 * t)) & 0xFFu) == 0) s.put((z >> (8 * t)) & 3u, 2);
      }
    }
    for (; j < L; ++j) {
      const u8 c = sp[j];
      if (T.xq[c]) continue;
      acc = (acc << 2) | T.smap[c];
      if (++cnt == 16) { s.put(acc, 32); acc = 0; cnt = 0; }
    }
    s.put(acc, 2 * cnt);
    return;
  }
  if (T.dna_mode) {
    u32 w = 0, cnt = 0;
    for (u32 j = 0; j < L; ++j) {
      const u8 c = sp[j];
      if (xfer && T.xq[c]) continue; /* every non-ACGT base of such a record carries a transferable code */
      w = (w << 2) | T.smap[c];
      if (++cnt == 16) { s.put(w, 32); w = 0; cnt = 0; }
    }
    s.put(w, 2 * cnt);
    return;
  }
  for (u32 j = 0; j < L; ++j) {
    const u8 c = sp[j];
    if (xfer && T.xq[c]) continue;
    const u64 e = T.dcl[T.smap[c]];
    s.put((u32)e, (u32)(e >> 32));
  }
}

/* Fills the CTA's walker tables; `pk_smem` (d.pk_bytes of shared memory) receives the packed quality tables when they fit. */
__device__ __forceinline__ void load_walk_tabs(const Dev &d, const SbClass &C, const u32 *arena, const TableDesc *td, u8 *codes /*[512]*/, u8 *xq /*[256]*/,
                                               u16 *pk_smem, WalkTabs &T) {
  for (u32 i = threadIdx.x; i < 64; i += blockDim.x) { ((u32 *)codes)[i] = ((const u32 *)C.qua_code)[i]; ((u32 *)codes)[64 + i] = ((const u32 *)C.sym_code)[i]; }
  load_xq(xq);
  T.qmap = codes; T.smap = codes + 256; T.xq = xq; T.nq = C.nq;
  T.qcl = (const u64 *)(arena + td[C.tq0].cl_off);
  const u32 pk_bytes = (C.max_qlen + 1) * C.nq * 2u;
  T.pk = nullptr; T.has_esc = C.qpk_esc != 0;
  if (!C.qpk_bad && pk_bytes <= d.pk_bytes) {
    const uint4 *src = (const uint4 *)(arena + C.qpk_off);
    for (u32 i = threadIdx.x; i < (pk_bytes + 15) / 16; i += blockDim.x) ((uint4 *)pk_smem)[i] = src[i];
    T.pk = pk_smem;
  }
  T.dna_mode = !C.plain ? 0u : (C.nsym == 4 && C.symbols[0] == 'A' && C.symbols[1] == 'C' && C.symbols[2] == 'G' && C.symbols[3] == 'T') ? 2u : 1u;
  T.dcl = C.plain ? (const u64 *)nullptr : (const u64 *)(arena + td[C.tdna].cl_off);
}

/* ---- lengths ----------------------------------------------------------------------------------------------------- */
/* Bit length of every record in the three bodies.  Offsets are kept two-level: local to the 128-record chunk
 * (qoff/doff/toff per record, title block offsets per 32-record block) plus one total per chunk that k_layout
 * turns into chunk bases with a short scan. */
/* The encoder kernels (k_lengths, k_emit) are warp-autonomous: a warp owns EGW consecutive 32-record blocks (one title
 * block = one warp = one lane per record), streams each block's bytes into its own shared-memory stage with the
 * bulk-copy engine and never meets the other warps of the CTA after the tables are loaded -- no block barrier in the
 * loop, so a warp with long or ambiguous records does not hold the others up.
 * dynamic shared memory: [pk_bytes packed quality tables][EW stages of enc_stage bytes] */
constexpr int EW = 8;   /* warps per encoder CTA */
#ifndef PHY_EGW
#define PHY_EGW 8
#endif
constexpr int EGW = PHY_EGW;  /* 32-record blocks per warp */

struct WarpStage {
  u32 buf_a, bar_a, phase;
  const u8 *buf;
  __device__ __forceinline__ void init(void *smem, u64 *bar, bool leader) { /* followed by a barrier of the caller */
    buf = (const u8 *)smem; buf_a = (u32)__cvta_generic_to_shared(smem); bar_a = (u32)__cvta_generic_to_shared(bar); phase = 0;
    if (leader) { mbar_init(bar_a, 1); mbar_fence_init(); }
  }
  __device__ __forceinline__ void request(const u8 *in, u32 lo, u32 hi) const { span_request(in, lo, hi, buf_a, bar_a); } /* one lane */
  __device__ __forceinline__ const u8 *wait(u32 lo) { mbar_wait(bar_a, phase); phase ^= 1u; return buf - (lo & ~15u); }
};

constexpr int EP = EW / 2; /* warp pairs (= stages) per CTA of the paired encoder kernels */
__device__ __forceinline__ void pair_sync(u32 pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

/* k_lengths<false>: one warp per 32-record block counts all three streams.  k_lengths<true>: two warps share a block's
 * stage -- the even warp counts the quality bits, the odd warp the DNA bits and the title bits; each scans and stores
 * its own offsets, so the pair only meets at its named barrier before the next block is requested (long records:
 * twice the warps per staged byte). */
template <bool PAIR>
__global__ void __launch_bounds__(EW * 32) k_lengths(Dev d) {
  constexpr u32 NST = PAIR ? EP : EW; /* stages per CTA */
  extern __shared__ uint4 dyn_smem[];
  __shared__ TitleTabs TT;
  __shared__ __align__(16) u8 codes[512];
  __shared__ __align__(16) u8 xq[256];
  __shared__ __align__(16) u8 lut[256];
  __shared__ __align__(8) u64 bars[NST];
  __shared__ u32 g_lo[NST][EGW + 1];
  const u32 s = blockIdx.y, lane = threadIdx.x & 31, w = PAIR ? threadIdx.x >> 6 : threadIdx.x >> 5; /* w: stage */
  const u32 role = PAIR ? (threadIdx.x >> 5) & 1u : 0u;
  const bool do_q = !PAIR || role == 0, do_d = !PAIR || role == 1;
  SbClass &C = d.cls[s];
  if (C.status || C.fast) return; /* fast: the single-walk kernels (phy_encode.cuh) encode this subblock */
  const u32 nblk = C.nblk;
  if (blockIdx.x * (NST * EGW) >= nblk) return;
  const u32 g0 = min((blockIdx.x * NST + w) * EGW, nblk), g1 = min(g0 + EGW, nblk);
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const TableDesc *td = (const TableDesc *)(arena + C.tabdesc_off);
  const u32 R = C.R, nnc = C.nnc, plain = C.plain, flagbits_off = C.flagbits_off, blk3_off = C.blk3_off;
  const bool tit = do_d && nnc;
  WalkTabs T;
  load_walk_tabs(d, C, arena, td, codes, xq, (u16 *)dyn_smem, T);
  load_lut(lut);
  load_title_tabs(C, TT);
  if (role == 0 && lane <= g1 - g0) g_lo[w][lane] = d.rstart[P.first_rec + min((g0 + lane) * 32, R)];
  WarpStage stage; stage.init((u8 *)dyn_smem + d.pk_bytes + w * d.enc_stage, &bars[w], role == 0 && lane == 0);
  __syncthreads();
  if (g0 >= g1) return;
  {
    bool fits = true;
    for (u32 k = 0; k < g1 - g0; ++k) fits = fits && g_lo[w][k + 1] - (g_lo[w][k] & ~15u) + 16 <= d.enc_stage;
    if (!fits) { if (lane == 0) atomicMin(&C.status, (i32)E_UNSUPPORTED); return; } /* records far beyond the reference's 500-byte domain */
  }
  if (role == 0 && lane == 0) stage.request(d.in, g_lo[w][0], g_lo[w][1]);
  /* this lane's record of the first block (idle lanes shadow the block's last record so that the warp stays converged) */
  u32 n_te, n_se, n_rs = 0, n_kx, n_fl = 0;
  {
    const u32 i = min(g0 * 32 + lane, R - 1), r = P.first_rec + i;
    n_te = d.te[r]; n_se = d.se[r]; n_kx = d.kx[r];
    if (tit) { n_rs = d.rstart[r]; n_fl = arena[flagbits_off + g0]; }
  }
  for (u32 g = g0; g < g1; ++g) {
    const u32 nrec = min(32u, R - g * 32);
    const bool active = lane < nrec;
    const u32 r = P.first_rec + g * 32 + (active ? lane : nrec - 1);
    const u32 te = n_te, se = n_se, rs_r = n_rs, kx = n_kx, myflags = n_fl, L = se - te - 1;
    if (g + 1 < g1) { /* next block's record, in flight while this block is walked */
      const u32 i = min((g + 1) * 32 + lane, R - 1), rn = P.first_rec + i;
      n_te = d.te[rn]; n_se = d.se[rn]; n_kx = d.kx[rn];
      if (tit) { n_rs = d.rstart[rn]; n_fl = arena[flagbits_off + g + 1]; }
    }
    const u8 *b = stage.wait(g_lo[w][g - g0]);
    const bool xfer = kx >> 15;
    u32 qbits = 0, dbits = 0, tb = 0;
    if (do_q) {
      CountSink q; q.init();
      quality_walk(b + se + 3, b + te + 1, L, xfer, T, q);
      qbits = active ? (u32)q.bits : 0u;
    }
    if (do_d) {
      if (plain) dbits = 2 * (kx & 0x7FFFu);
      else {
        CountSink dn; dn.init();
        dna_walk(b + te + 1, L, xfer, T, dn);
        dbits = (u32)dn.bits;
      }
      if (!active) dbits = 0;
      if (nnc) {
        __syncwarp();
        CountSink t; t.init();
        title_record(b, lut, rs_r, te, C, TT.fc, TT.ncf, TT.ncskip, arena, myflags, lane == 0, PrevShfl(), t);
        tb = active ? (u32)t.bits : 0u;
      }
    }
    if (g + 1 < g1) { /* every warp that reads the stage has left it */
      if (PAIR) pair_sync(w); else __syncwarp();
      if (role == 0 && lane == 0) stage.request(d.in, g_lo[w][g + 1 - g0], g_lo[w][g + 2 - g0]);
    }
    /* offsets inside the block */
    if (do_q) {
      u32 qx = qbits;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xFFFFFFFFu, qx, o); if (lane >= (u32)o) qx += y; }
      if (active) d.qoff[r] = qx - qbits;
      if (lane == 31) arena[blk3_off + g] = qx;
    }
    if (do_d) {
      u32 dx = dbits, tx = tb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 yd = __shfl_up_sync(0xFFFFFFFFu, dx, o), yt = __shfl_up_sync(0xFFFFFFFFu, tx, o);
        if (lane >= (u32)o) { dx += yd; tx += yt; }
      }
      if (active) {
        d.doff[r] = dx - dbits;
        if (nnc) d.toff[r] = tx - tb; /* bits of the block's earlier records (the flag bits come on top) */
      }
      if (lane == 31) {
        arena[blk3_off + nblk + g] = dx;
        arena[blk3_off + 2 * nblk + g] = nnc ? (nnc + tx + 7) / 8 : 0u;
      }
    }
  }
}

/* ---- layout: headers, scans, section sizes ------------------------------------------------------------------------ */
/* in-place exclusive scan of a[0..n) by one 256-thread CTA; returns the total (64-bit) */
__device__ u64 cta_scan_inplace(u32 *a, u32 n, u32 *ws, i32 *overflow) {
  u64 carry = 0;
  for (u32 base = 0; base < n; base += 256) {
    u32 i = base + threadIdx.x;
    u32 v = i < n ? a[i] : 0u, tot;
    u32 ex = block_excl_scan_256(v, ws, tot);
    u64 o = carry + ex;
    if (i < n && o > 0xFFFFFFFFull) *overflow = 1;
    if (i < n) a[i] = (u32)o;
    carry += tot;
  }
  return carry;
}

constexpr u32 LAYOUT_TABS = 2048; /* table lengths / destinations kept in shared memory while the headers are laid out */

__global__ void __launch_bounds__(256) k_layout(Dev d) {
  __shared__ u32 ws[8];
  __shared__ i32 ovf, okh;
  __shared__ u32 tlen[LAYOUT_TABS], tdst[LAYOUT_TABS];
  const u32 s = blockIdx.x, tid = threadIdx.x;
  SbClass &C = d.cls[s];
  if (C.status) return;
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  TableDesc *td = (TableDesc *)(arena + C.tabdesc_off);
  if (C.ntab > LAYOUT_TABS) { if (tid == 0) C.status = E_CAPACITY; return; }
  for (u32 t = tid; t < C.ntab; t += 256) { tlen[t] = td[t].tree_len; tdst[t] = 0; }
  if (tid == 0) ovf = 0;
  __syncthreads();
  if (tid == 0) okh = layout_headers(d.in, C, arena, tlen, tdst) ? 1 : 0;
  __syncthreads();
  if (!okh) { if (tid == 0) C.status = E_UNSUPPORTED; return; }
  /* copy the tree blobs to their place in the staged headers, one warp per table */
  {
    u8 *stage = (u8 *)(arena + C.stage_off);
    for (u32 t = tid >> 5; t < C.ntab; t += 8) {
      const u8 *src = (const u8 *)(arena + td[t].tree_off);
      u8 *dst = stage + tdst[t];
      for (u32 i = tid & 31; i < tlen[t]; i += 32) dst[i] = src[i];
    }
  }
  u64 qb, db, tb;
  if (C.fast) { /* single-walk encoder: one total per task of 256 records; the scans go behind the totals */
    const u32 nt = C.ntask;
    u32 *len3 = arena + C.task_off, *base3 = len3 + 3 * nt;
    for (u32 i = tid; i < 3 * nt; i += 256) base3[i] = len3[i];
    __syncthreads();
    qb = cta_scan_inplace(base3, nt, ws, &ovf);
    db = cta_scan_inplace(base3 + nt, nt, ws, &ovf);
    tb = cta_scan_inplace(base3 + 2 * nt, nt, ws, &ovf);
  } else {
    qb = cta_scan_inplace(arena + C.blk3_off, C.nblk, ws, &ovf);
    db = cta_scan_inplace(arena + C.blk3_off + C.nblk, C.nblk, ws, &ovf);
    tb = C.nnc ? cta_scan_inplace(arena + C.blk3_off + 2 * C.nblk, C.nblk, ws, &ovf) : 0;
  }
  __syncthreads();
  if (tid == 0) {
    if (ovf || tb > 0x7FFFFFFFull) { C.status = E_CAPACITY; return; }
    finish_layout(C, (u32)tb, qb, db);
  }
}

__global__ void __launch_bounds__(256) k_outscan(Dev d) {
  __shared__ u32 ws[8];
  const u32 S = d.hdr->S;
  const u64 begin = d.prev_total ? *d.prev_total : 0ull; /* payloads of all groups lie back to back */
  u64 carry = begin;
  for (u32 base = 0; base < S; base += 256) {
    const u32 s = base + threadIdx.x;
    u32 len = 0, pad = 0;
    if (s < S) { len = d.cls[s].status ? 0u : d.cls[s].payload_len; pad = (len + 15u) & ~15u; }
    /* payloads are 16-byte aligned: scan in 16-byte units (a batch's output is far below 64 GiB) */
    u32 tot;
    const u32 ex = block_excl_scan_256(pad >> 4, ws, tot);
    if (s < S) {
      SbClass &C = d.cls[s];
      const u64 off = carry + ((u64)ex << 4);
      SbOut o;
      o.status = C.status; o.out_off = off; o.out_len = len;
      o.sec_len[0] = C.info_len; o.sec_len[1] = C.title_len; o.sec_len[2] = C.qual_len; o.sec_len[3] = C.dna_len;
      /* a payload that does not fit takes no space, but the later ones keep the offsets of the scan: they all lie
       * beyond the capacity as well and fail the same way */
      if (!C.status && off + len > d.out_cap) { C.status = E_CAPACITY; o.status = E_CAPACITY; o.out_len = 0; }
      C.out_off = off;
      d.sbout[s] = o;
    }
    carry += (u64)tot << 4;
  }
  if (threadIdx.x == 0) { d.hdr->out_begin = begin < d.out_cap ? begin : d.out_cap; d.hdr->total_out = carry < d.out_cap ? carry : d.out_cap; }
}

__global__ void __launch_bounds__(256) k_zero_out(Dev d) {
  const u64 b16 = d.hdr->out_begin / 16, e16 = (d.hdr->total_out + 15) / 16; /* payloads are 16-byte aligned */
  uint4 z = make_uint4(0, 0, 0, 0);
  for (u64 i = b16 + (u64)blockIdx.x * 256 + threadIdx.x; i < e16; i += (u64)gridDim.x * 256) ((uint4 *)d.out)[i] = z;
}

/* ---- emit ------------------------------------------------------------------------------------------------------------ */
/* `base` is 4-byte aligned */
__device__ __forceinline__ void or_byte(u8 *base, u32 pos, u8 v) {
  if (v) atomicOr((u32 *)(base + (pos & ~3u)), (u32)v << (8 * (pos & 3u)));
}

/* k_emit<false>: one warp per 32-record block writes all four streams (short records: enough warps fit an SM).
 * k_emit<true> pairs two warps on every block: the even warp writes the info length bits and the quality codes, the
 * odd warp the DNA and the title tokens.  The streams are independent, the pair shares one stage (twice the warps per
 * staged byte -- what long records need to keep an SM busy) and meets only at its own named barrier before the next
 * block is requested. */

template <bool PAIR>
__global__ void __launch_bounds__(EW * 32, PAIR ? 5 : 4) k_emit(Dev d) {
  constexpr u32 NST = PAIR ? EP : EW; /* stages per CTA */
  extern __shared__ uint4 dyn_smem[];
  __shared__ TitleTabs TT;
  __shared__ __align__(16) u8 codes[512];
  __shared__ __align__(16) u8 xq[256];
  __shared__ __align__(16) u8 lut[256];
  __shared__ __align__(8) u64 bars[NST];
  __shared__ u32 g_lo[NST][EGW + 1];
  const u32 s = blockIdx.y, lane = threadIdx.x & 31, w = PAIR ? threadIdx.x >> 6 : threadIdx.x >> 5; /* w: stage */
  const u32 role = PAIR ? (threadIdx.x >> 5) & 1u : 0u;
  const bool do_q = !PAIR || role == 0, do_d = !PAIR || role == 1;
  const SbClass &C = d.cls[s];
  if (C.status || C.fast) return;
  const u32 nblk = C.nblk;
  if (blockIdx.x * (NST * EGW) >= nblk) return;
  const u32 g0 = min((blockIdx.x * NST + w) * EGW, nblk), g1 = min(g0 + EGW, nblk);
  const SbPlan P = d.plans[s];
  u32 *arena = d.arena + (size_t)s * d.arena_words;
  const TableDesc *td = (const TableDesc *)(arena + C.tabdesc_off);
  u8 *out = d.out + C.out_off;
  u32 *outw = (u32 *)d.out;
  const u64 obase = C.out_off; /* byte offset of the payload inside d.out (16-byte aligned) */
  const u32 R = C.R, nnc = C.nnc, nb_len = C.nb_len;
  WalkTabs T;
  load_walk_tabs(d, C, arena, td, codes, xq, (u16 *)dyn_smem, T);
  load_lut(lut);
  load_title_tabs(C, TT);
  if (role == 0 && lane <= g1 - g0) g_lo[w][lane] = d.rstart[P.first_rec + min((g0 + lane) * 32, R)];
  WarpStage stage; stage.init((u8 *)dyn_smem + d.pk_bytes + w * d.enc_stage, &bars[w], role == 0 && lane == 0);
  __syncthreads();
  const u32 o_title = C.info_len, o_qual = o_title + C.title_len, o_dna = o_qual + C.qual_len;
  if (blockIdx.x == 0) {
    /* fixed part of the info stream (phyNGSC.cpp:719-730) and the three staged headers.  Bytes are OR-ed
     * into the zeroed payload word-atomically because a header may end inside a word whose other bytes
     * belong to a bit stream written by another thread. */
    if (threadIdx.x == 0) {
      u8 fx[INFO_FIXED];
      ByteWriter bw; bw.p = fx; bw.n = 0;
      bw.word(C.R); bw.word(C.max_qlen); bw.word(C.max_slen);
      bw.byte((u8)C.nsym); bw.byte(0); bw.byte((u8)C.nq); bw.word(C.flags);
      for (u32 i = 0; i < INFO_FIXED; ++i) or_byte(out, i, fx[i]);
    }
    const u8 *hs = (const u8 *)(arena + C.stage_off);
    for (u32 i = threadIdx.x; i < C.thdr_len; i += EW * 32) or_byte(out, o_title + i, hs[i]);
    for (u32 i = threadIdx.x; i < C.qhdr_len; i += EW * 32) or_byte(out, o_qual + i, hs[C.thdr_cap + i]);
    for (u32 i = threadIdx.x; i < C.dhdr_len; i += EW * 32) or_byte(out, o_dna + i, hs[C.thdr_cap + C.qhdr_cap + i]);
  }
  if (g0 >= g1) return;
  if (role == 0 && lane == 0) stage.request(d.in, g_lo[w][0], g_lo[w][1]); /* k_lengths has checked that every block fits its stage */
  /* bit positions of the three bodies inside d.out; per 32-record block: base of its quality bits, DNA bits, title bytes */
  const u64 q_bit0 = (obase + o_qual + C.qhdr_len) * 8, d_bit0 = (obase + o_dna + C.dhdr_len) * 8, t_byte0 = obase + o_title + C.thdr_len;
  const u32 *qbase_p = arena + C.blk3_off, *dbase_p = qbase_p + nblk, *tbase_p = dbase_p + nblk, *flag_p = arena + C.flagbits_off;
  /* this lane's record of the first block (idle lanes shadow the block's last record) */
  const bool tit = do_d && nnc;
  u32 n_te, n_se, n_kx, n_qoff = 0, n_doff = 0, n_qb = 0, n_db = 0, n_rs = 0, n_toff = 0, n_fl = 0, n_tb = 0;
  {
    const u32 i = min(g0 * 32 + lane, R - 1), r = P.first_rec + i;
    n_te = d.te[r]; n_se = d.se[r]; n_kx = d.kx[r];
    if (PAIR) { n_qoff = (role ? d.doff : d.qoff)[r]; n_qb = (role ? dbase_p : qbase_p)[g0]; } /* the role's own stream */
    else { n_qoff = d.qoff[r]; n_qb = qbase_p[g0]; n_doff = d.doff[r]; n_db = dbase_p[g0]; }
    if (tit) { n_rs = d.rstart[r]; n_toff = d.toff[r]; n_fl = flag_p[g0]; n_tb = tbase_p[g0]; }
  }
  for (u32 g = g0; g < g1; ++g) {
    const u32 nrec = min(32u, R - g * 32);
    const bool active = lane < nrec;
    const u32 i_sb = g * 32 + (active ? lane : nrec - 1); /* record index inside the subblock */
    const u32 te = n_te, se = n_se, kx = n_kx, qoff = n_qoff, qb = n_qb, doff = PAIR ? n_qoff : n_doff, db = PAIR ? n_qb : n_db;
    const u32 rs_r = n_rs, my_toff = n_toff, flags = n_fl, tblk = n_tb;
    const u32 L = se - te - 1;
    if (g + 1 < g1) { /* next block's record, in flight while this block is encoded */
      const u32 i = min((g + 1) * 32 + lane, R - 1), rn = P.first_rec + i;
      n_te = d.te[rn]; n_se = d.se[rn]; n_kx = d.kx[rn];
      if (PAIR) { n_qoff = (role ? d.doff : d.qoff)[rn]; n_qb = (role ? dbase_p : qbase_p)[g + 1]; }
      else { n_qoff = d.qoff[rn]; n_qb = qbase_p[g + 1]; n_doff = d.doff[rn]; n_db = dbase_p[g + 1]; }
      if (tit) { n_rs = d.rstart[rn]; n_toff = d.toff[rn]; n_fl = flag_p[g + 1]; n_tb = tbase_p[g + 1]; }
    }
    const u8 *b = stage.wait(g_lo[w][g - g0]);
    const bool xfer = kx >> 15;
    if (do_q) {
      if (active) { /* per-record length bits of the info stream (phyNGSC.cpp:732-742; always present) */
        OrSink k; k.init(outw, (obase + INFO_FIXED) * 8 + (u64)i_sb * nb_len);
        k.put(L, nb_len); k.finish();
      }
      /* idle lanes of the last block walk the block's last record without storing */
      OrSink q; q.init(outw, q_bit0 + qb + qoff, active);
      quality_walk(b + se + 3, b + te + 1, L, xfer, T, q);
      q.finish();
    }
    if (do_d) {
      {
        OrSink dn; dn.init(outw, d_bit0 + db + doff, active);
        dna_walk(b + te + 1, L, xfer, T, dn);
        dn.finish();
      }
      if (nnc) {
        /* title body: blocks of 32 records, byte-aligned (tasks.cpp:393-509); all 32 lanes walk together */
        __syncwarp();
        OrSink t; t.init(outw, (t_byte0 + tblk) * 8 + (lane == 0 ? 0u : nnc + my_toff), active);
        if (lane == 0) {
          u32 v = 0;
          for (u32 k = 0; k < nnc; ++k) v = (v << 1) | ((flags >> TT.ncf[k]) & 1u);
          t.put(v, nnc);
        }
        title_record(b, lut, rs_r, te, C, TT.fc, TT.ncf, TT.ncskip, arena, flags, lane == 0, PrevShfl(), t);
        t.finish();
      }
    }
    if (g + 1 < g1) { /* every warp that reads the stage has left it */
      if (PAIR) pair_sync(w); else __syncwarp();
      if (role == 0 && lane == 0) stage.request(d.in, g_lo[w][g + 1 - g0], g_lo[w][g + 2 - g0]);
    }
  }
}

}  // namespace phy
