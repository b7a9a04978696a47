/*
 * phy_fast.cuh -- bit-level pieces of the single-walk encoder, written once as __host__ __device__ functions
 * (the kernels in phy_encode.cuh call them per lane; tests/mirror/ runs the same functions lane by lane on the CPU).
 *
 * The single-walk encoder replaces "count every record's bits, scan, walk again and write" (k_lengths + k_emit)
 * by: every lane encodes its piece of a stream ONCE into lane-private staging words (LaneSink), the warp concatenates
 * the lanes' pieces in shared memory (lane_concat), appends the result to the task's slot of a temporary buffer in
 * HBM (StreamState) and records the task's exact length; after a scan of the task lengths one more kernel moves
 * every task's run to its final bit position of the payload (shifted_word).  The bits are those of
 * BitStream::PutBits / FlushPartialWordBuffer (bit_stream.h:80-265): MSB first, big-endian 32-bit words; inside the
 * encoder a word is kept "logical" (stream bit i of a word = bit 31 - i) and byte-swapped only by the final store.
 */
#pragma once
#include "phy_core.cuh"

namespace phy {

constexpr u32 TASK_BLOCKS = 8;                  /* 32-record title blocks per encoder task (one warp)         */
constexpr u32 TASK_RECORDS = 32 * TASK_BLOCKS;
constexpr u32 PIECE_WORDS = 512;                /* source words one warp of the placement kernel moves at a time */
constexpr u32 LPW_T = 17;                       /* lane-private staging words of a record's title tokens (odd: lanes in different banks) */

/* funnel shifts with the shift amount clamped at 32 (the device has them as one instruction) */
PHY_HD u32 shl_c(u32 lo, u32 hi, u32 n) { /* high word of (hi:lo) << n, n <= 32 */
#if defined(__CUDA_ARCH__)
  return __funnelshift_lc(lo, hi, n);
#else
  return n == 0 ? hi : n >= 32 ? lo : (hi << n) | (lo >> (32 - n));
#endif
}
PHY_HD u32 shr_c(u32 lo, u32 hi, u32 n) { /* low word of (hi:lo) >> n, n <= 32 */
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, n);
#else
  return n == 0 ? lo : n >= 32 ? hi : (lo >> n) | (hi << (32 - n));
#endif
}

/* ---- lane-private bit sink ---------------------------------------------------------------------------------- */
/* Appends bits MSB-first; completed 32-bit logical words go to the lane's staging slots, `stride` words apart
 * (on the GPU: shared memory, word k of lane l at [k * 32 + l], so a warp's stores never collide).  The pending
 * bits sit right-aligned in a 64-bit accumulator: bits [0, fill). */
/* where the completed words go: plain memory here, shared memory by address on the GPU (SmemStore in phy_encode.cuh) */
struct PtrStore {
  u32 *p; u32 stride;
  PHY_HD void put(u32 w) { *p = w; }
  PHY_HD void next() { p += stride; }
};
template <class Store>
struct LaneSinkT {
  u32 lo, hi, fill, nwords, cap;
  bool over; /* more words than the staging holds: cannot happen when the caller's bound holds (k_slots checks it before a subblock takes this path) */
  Store st;
  PHY_HD void init(const Store &store_, u32 cap_words) { lo = hi = 0; fill = 0; nwords = 0; cap = cap_words; over = false; st = store_; }
  PHY_HD void store(u32 w) { st.put(w); st.next(); ++nwords; }
  PHY_HD void put(u32 v, u32 n) { /* n <= 32, v < 2^n */
    hi = shl_c(lo, hi, n);
    lo = shl_c(0u, lo, n) | v;
    fill += n;
    if (fill >= 32) { fill -= 32; store(shr_c(lo, hi, fill)); }
  }
  /* left-aligns the pending bits into one last zero-padded word; returns the bits written in total */
  PHY_HD u32 finish() {
    const u32 n = 32u * nwords + fill;
    if (fill) { store(lo << (32 - fill)); fill = 0; }
    over = nwords > cap;
    return n;
  }
};

/* ---- warp concatenation ----------------------------------------------------------------------------------------- */
/* One lane: ORs its `nbits` staged bits (words lp[0], lp[stride], ...) into the concatenation buffer `cc` at bit
 * position `pos`.  cc is zero where no lane has written yet; neighbouring lanes share boundary words, hence the OR
 * (atomic on the GPU). */
PHY_HD void cc_or(u32 *cc, u32 idx, u32 v) {
  if (!v) return;
#if defined(__CUDA_ARCH__)
  atomicOr(cc + idx, v);
#else
  cc[idx] |= v;
#endif
}
PHY_HD void lane_concat(u32 *cc, u32 pos, const u32 *lp, u32 stride, u32 nbits) {
  const u32 sh = pos & 31, nw = (nbits + 31) >> 5;
  const u32 own_lo = (pos + 31) >> 5, own_hi = (pos + nbits) >> 5; /* words [own_lo, own_hi) hold bits of this lane only: plain stores */
  u32 idx = pos >> 5, prev = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (u32 k = 0; k < nw; ++k, ++idx) { /* a few trips: kept rolled */
    const u32 w = lp[k * stride];
    const u32 v = shr_c(w, prev, sh); /* (prev:w) >> sh: the tail of the previous source word and the head of this one */
    if (idx >= own_lo && idx < own_hi) cc[idx] = v; else cc_or(cc, idx, v);
    prev = w;
  }
  if (sh) cc_or(cc, idx, prev << (32 - sh));
}

/* Where a task's stream stands: `carry` bits of cc[0] are pending (the words before them are in the temporary buffer
 * already), `tpos` is the next word of the task's slot, `total` the bits appended so far.  Uniform across the warp. */
struct StreamState {
  u32 carry, tpos, total;
  PHY_HD void init() { carry = 0; tpos = 0; total = 0; }
  /* after the lanes have concatenated `bits` more bits behind the carry: how many whole words of cc leave now */
  PHY_HD u32 full_words(u32 bits) const { return (carry + bits) >> 5; }
  PHY_HD void advance(u32 bits) { const u32 nf = (carry + bits) >> 5; tpos += nf; carry = (carry + bits) & 31; total += bits; }
  /* FlushPartialWordBuffer (bit_stream.h:183-200): zero bits up to the next byte boundary */
  PHY_HD u32 pad_to_byte() const { return (8u - (total & 7u)) & 7u; }
};

/* ---- final placement --------------------------------------------------------------------------------------------- */
/* A run of `nbits` logical bits in src[0 .. ceil(nbits/32)) (zero-padded) is placed at bit offset `sh` (0..31) of
 * destination word 0: destination word j of the ceil((sh + nbits)/32) words it touches. */
PHY_HD u32 shifted_word(const u32 *src, u32 nsrc, u32 sh, u32 j) {
  const u32 hi = j ? src[j - 1] : 0u, lo = j < nsrc ? src[j] : 0u;
  return sh ? (hi << (32 - sh)) | (lo >> sh) : lo;
}

/* ---- staging bounds ------------------------------------------------------------------------------------------------ */
/* Longest possible code of every table (k_huff stores it in the directory) bounds what one record can emit; the
 * bounds size the lane-private staging and the tasks' slots in the temporary buffer.  A subblock whose bounds exceed the
 * staging the kernels were launched with is encoded by the two-walk kernels instead (SbClass::fast = 0). */
struct FastGeom {          /* launch geometry of the single-walk kernels (host decides, device checks)          */
  u32 g;                   /* lanes per record in the quality / DNA kernel: 1, 2, 4 or 8                         */
  u32 lpw_q;               /* staging words per lane there                                                       */
  u32 pk_bytes;            /* shared memory reserved for the packed quality tables                               */
};

PHY_HD u32 seg_len(u32 L, u32 g) { return (L + 4 * g - 1) / (4 * g) * 4; } /* positions per lane of a read of L symbols: a multiple of four */

/* bits one record's title tokens can take at most (tasks.cpp:427-506); lanes split the per-position tables */
PHY_HD u32 title_bound_part(const SbClass &C, const u32 *arena, const TableDesc *td, u32 lane, u32 nl) {
  u32 b = 0;
  for (u32 k = 0; k < C.nnc; ++k) {
    const FieldClass &F = C.f[C.ncf[k]];
    if (F.kind == K_NUM) {
      if (lane == 0) {
        const u32 code = F.has_table ? td[F.tab].maxlen : F.bits_num;
        b += code > F.bits_val ? code : F.bits_val;
      }
      continue;
    }
    if (lane == 0 && !F.is_len_const) b += F.bits_len;
    const u16 *sm = (const u16 *)(arena + F.slotmap_off);
    const u32 nt = F.max_len < 128 ? F.max_len : 128;
    for (u32 j = lane; j < nt; j += nl)
      if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) b += td[sm[j]].maxlen;
    if (lane == 0 && F.max_len > 128) b += (F.max_len - 128) * td[sm[128]].maxlen;
  }
  return b;
}

}  // namespace phy
