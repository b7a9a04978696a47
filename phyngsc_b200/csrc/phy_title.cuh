/*
 * phy_title.cuh -- the title kernels (included by phy_b200.cu after phy_encode.cuh).
 *
 *   k_stat1      the only kernel that tokenises title lines: per-field reductions of AnalyzeTitleFields (tasks.cpp:22-223
 *                as closed forms: length and value extremes, numeric-ness, delta extremes, Hamming masks against record 0)
 *                and, per 32-record block, the mask of the fields in which some record differs from record 0 of the
 *                subblock plus, for those fields, one row each of numeric values (Dev::tv) and token positions (Dev::tp)
 *   k_xdelta     numeric deltas across the boundaries of k_stat1's tasks
 *   k_stat2      numeric / per-position char histograms and the 32-record block flags (tasks.cpp:64-93, 127-182)
 *   k_enc_title  info length bits and title tokens of the single-walk encoder (phyNGSC.cpp:732-742, tasks.cpp:393-509)
 *
 * k_stat2 and k_enc_title read the parsed rows instead of the title bytes: numeric fields never touch the input again,
 * string fields fetch exactly their characters from the input (a token that is record 0's is read from record 0's title
 * line).  All three record kernels are warp-autonomous (warp = task of 8 consecutive 32-record blocks, lane = record):
 * no block barrier inside the walk; k_stat1 stages only title lines (one slot per lane), the other two stage nothing.
 */
#pragma once
#include "phy_encode.cuh"

namespace phy {

/* where the rows of one 32-record block start, and which of its fields have rows */
struct BlockRows {
  size_t row0;     /* entry index of (block, field 0, lane 0) in tv / tp */
  const u32 *v0;   /* numeric values of record 0's tokens (fields without a row) */
  u32 mask;
};
__device__ __forceinline__ BlockRows block_rows(const Dev &d, const SbPlan &P, u32 s, u32 g) {
  BlockRows c;
  const size_t blk = (size_t)P.chunk_base * (CH / 32) + g;
  c.row0 = blk * d.nfs * 32; c.v0 = d.v0 + (size_t)s * MAXF; c.mask = d.blk_mask[blk];
  return c;
}
/* One lane's entries of a (block, field) row, loaded ahead of their use (the walkers below fetch field k + 1's entries while
 * they work on field k: the rows are the only global memory they depend on).  Meaningless when the block has no row. */
struct RawRow { u32 e, v, c; }; /* tp, tv and -- string fields -- tc; numeric fields carry the block before's last value in c (lane 0) */
template <bool AHEAD> /* AHEAD: fetched before the entries are looked at, so tc comes along unconditionally; else only when the token needs it */
__device__ __forceinline__ RawRow load_row(const Dev &d, const BlockRows &c, const BlockRows &prev, bool have_prev, u32 f, u32 lane, bool numeric) {
  RawRow r; r.e = r.v = r.c = 0;
  const size_t i = c.row0 + (size_t)f * 32 + lane;
  if ((c.mask >> f) & 1u) {
    r.v = d.tv[i];
    if (!numeric) {
      r.e = d.tp[i];
      if (AHEAD || (!(r.e & TP_NUM) && (r.e >> 16) > 4 && (r.e >> 16) <= 8 && (r.e & 0x7FFFu) != TP_SAME)) r.c = d.tc[i];
    }
  }
  if (numeric && have_prev && lane == 0 && ((prev.mask >> f) & 1u)) r.c = d.tv[prev.row0 + (size_t)f * 32 + 31];
  return r;
}
/* numeric value of field f of this lane's record (utils::to_num of its token) */
__device__ __forceinline__ u32 parsed_value(const BlockRows &c, u32 f, const RawRow &r) { return ((c.mask >> f) & 1u) ? r.v : c.v0[f]; } /* no row: record 0's token */
/* ... and of the last record of the block before (lane 0) */
__device__ __forceinline__ u32 parsed_prev_value(const BlockRows &prev, u32 f, const RawRow &r) { return ((prev.mask >> f) & 1u) ? r.c : prev.v0[f]; }
/* token of field f of this lane's record: its length and its characters -- in a register when the rows hold them
 * (inreg: character j in bits 8j.. of `chars`, tokens of up to 8 characters), else at offset `off` of the batch input */
struct TokRef { u32 off, len; u64 chars; bool inreg, same0; };
__device__ __forceinline__ u64 decimal_chars(u32 v, u32 len) { /* the len-digit decimal form of v, first digit in the low byte */
  u64 c = 0;
  for (u32 j = len; j-- > 0;) { const u32 q = v / 10u; c |= (u64)(v - 10u * q + '0') << (8 * j); v = q; }
  return c;
}
__device__ __forceinline__ TokRef parsed_token(const BlockRows &c, const SbClass &C, const FieldClass &F, u32 f, const RawRow &r, u32 ts) {
  TokRef t; t.off = C.ts0 + F.off0; t.len = F.len0; t.same0 = true; t.inreg = false; t.chars = 0;
  if (((c.mask >> f) & 1u) && (r.e & 0x7FFFu) != TP_SAME) {
    t.off = ts + (r.e & 0x7FFFu); t.len = r.e >> 16; t.same0 = false;
    if (t.len <= 8) {
      t.inreg = true;
      if (r.e & TP_NUM) t.chars = decimal_chars(r.v, t.len);
      else { t.chars = r.v; if (t.len > 4) t.chars |= (u64)r.c << 32; }
    }
  }
  return t;
}
__device__ __forceinline__ u32 token_char(const Dev &d, const TokRef &t, u32 j) { return t.inreg ? (u32)(t.chars >> (8 * j)) & 0xFFu : (u32)d.in[t.off + j]; }

/* ---- stat1 ---------------------------------------------------------------------------------------------------------- */
constexpr u32 S1W = 8; /* warps (tasks) per CTA */

struct Stat1S {
  u32 facc[MAXF][8];   /* FieldAcc scalars of the CTA (atomicMax encodings, see SbAcc) */
  u32 mism[MAXF][MASKW];
  i32 err;
  u32 nf;
  u32 off0[MAXF + 1], len0[MAXF]; /* record 0's tokens; off0[nf] = length of its title line with the newline */
  u32 v0[MAXF];                   /* numeric value / is_num of record 0's tokens */
  u8 num0[MAXF];
  u8 r0[R0_MAX];
};

/* Title field reductions (tasks.cpp:22-223 as closed forms) and the parsed rows.  A warp walks one task of TASK_BLOCKS
 * consecutive 32-record blocks of one subblock, lane = record; only the title lines are staged (two stages of one slot per
 * lane, 16-byte cp.async pieces: the next block's lines arrive while the current block is walked).  Record 0 is tokenised
 * once per CTA and the CTA's accumulators are flushed to the subblock's once.
 * Most tokens repeat record 0's: a token whose bytes AND separator equal record 0's is that token, so a warp whose 32
 * records all pass this comparison neither tokenises the field nor reduces anything -- record 0's own length and value
 * are folded into the accumulators once per CTA instead.  Fields that the warp has not seen differ so far are compared
 * as whole runs of consecutive fields, four bytes per step (`touched` steers only how the comparison is done, not its
 * result).  Numeric deltas (tasks.cpp:149-166 runs over all records): inside a block by warp shuffle, across the blocks of
 * a task through the last value of the block before (lastv, or record 0's value when that block had no row), across tasks
 * by k_xdelta from the first / last values every task leaves in chunk_first / chunk_last.
 * dynamic shared memory: per warp [2 stages of 32 slots of d.ts bytes] */
__global__ void __launch_bounds__(S1W * 32) k_stat1(Dev d) {
  extern __shared__ uint4 dyn_smem[];
  __shared__ Stat1S S;
  __shared__ __align__(16) u8 lut[256];
  __shared__ u32 lastv[S1W][MAXF];
  const u32 s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, NT = blockDim.x;
  const SbPlan P = d.plans[s];
  const u32 R = P.n_records, ntask = (R + TASK_RECORDS - 1) / TASK_RECORDS, nblk = (R + 31) / 32, nw = NT / 32;
  if (P.status || blockIdx.x * nw >= ntask) return;
  for (u32 i = tid; i < (MAXF * 8 + MAXF * MASKW + 2); i += NT) ((u32 *)&S)[i] = 0; /* accumulators, err, nf */
  load_lut(lut);
  const u32 ts0 = d.rstart[P.first_rec], te0 = d.te[P.first_rec];
  /* seed from record 0 of the subblock (phyNGSC.cpp:345-379) */
  const bool r0_ok = te0 - ts0 + 1 <= R0_MAX;
  if (r0_ok) for (u32 i = tid; i <= te0 - ts0; i += NT) S.r0[i] = d.in[ts0 + i];
  __syncthreads();
  if (tid == 0) {
    if (!r0_ok) S.err = E_UNSUPPORTED;
    else {
      TitleCursor c; c.init(S.r0, 0, te0 - ts0, lut);
      Tok t; u32 nf = 0;
      while (c.next(t)) { if (nf < (u32)MAXF) { S.off0[nf] = t.start; S.len0[nf] = t.end - t.start; S.v0[nf] = t.v; S.num0[nf] = t.num ? 1 : 0; } ++nf; }
      S.nf = nf;
      if (nf == 0 || nf > (u32)MAXF || nf > d.max_nf) S.err = E_UNSUPPORTED;
      else S.off0[nf] = te0 - ts0 + 1;
    }
  }
  __syncthreads();
  const u32 nf = S.nf;
  const bool seed_ok = S.err == 0;
  if (seed_ok && tid < nf) { /* record 0's own token lengths and values, folded in once per CTA (see above) */
    const u32 l0 = S.len0[tid], k0 = key_of((i32)S.v0[tid]);
    atomicMax(&S.facc[tid][0], ~l0); atomicMax(&S.facc[tid][1], l0);
    if (!S.num0[tid]) S.facc[tid][2] = 1;
    atomicMax(&S.facc[tid][3], k0); atomicMax(&S.facc[tid][4], ~k0);
    if (blockIdx.x == 0) d.v0[(size_t)s * MAXF + tid] = S.v0[tid];
  }
  const u32 task = blockIdx.x * nw + w;
  if (seed_ok && task < ntask) {
    const u32 tsz = d.ts;
    const u8 *slots = (const u8 *)dyn_smem + (size_t)w * 2u * 32u * tsz;
    const u32 slots_a = (u32)__cvta_generic_to_shared(slots);
    const u32 g0 = task * TASK_BLOCKS, g1 = min(g0 + TASK_BLOCKS, nblk);
    const size_t blk_base = (size_t)P.chunk_base * (CH / 32), task_row = ((size_t)P.chunk_base + 2u * task) * MAXF;
    const u32 nfmask = nf >= 32 ? 0xFFFFFFFFu : (1u << nf) - 1u;
    u32 touched = 0, prev_bm = 0, first_bm = 0, zero_d = 0;
    i32 err = 0;
    bool fits = true, n_fits = true; /* the slots are sized for the longest title line of the batch (BatchHdr::max_tlen): always true */
    u32 n_ts = 0, n_te = 0, nn_ts = 0, nn_te = 0; /* title line of this lane's record in the next block, and in the one after it */
    { const u32 i = g0 * 32 + lane; if (i < R) { const u32 r = P.first_rec + i; n_ts = d.rstart[r]; n_te = d.te[r]; n_fits = stage_title_line(d.in, slots_a + lane * tsz, tsz, n_ts, n_te); } }
    cp_async_commit();
    { const u32 i = (g0 + 1) * 32 + lane; if (g0 + 1 < g1 && i < R) { const u32 r = P.first_rec + i; nn_ts = d.rstart[r]; nn_te = d.te[r]; } }
    for (u32 g = g0; g < g1; ++g) {
      const u32 nrec = min(32u, R - g * 32), buf = (g - g0) & 1u;
      const bool active = lane < nrec;
      const u32 ts = n_ts, te = n_te;
      const bool cur_fits = n_fits;
      fits = fits && cur_fits;
      { /* next block's record: its title line starts to arrive now (its position was loaded one block earlier) */
        if (g + 1 < g1 && (g + 1) * 32 + lane < R) { n_ts = nn_ts; n_te = nn_te; n_fits = stage_title_line(d.in, slots_a + ((buf ^ 1u) * 32 + lane) * tsz, tsz, n_ts, n_te); }
        cp_async_commit();
        const u32 i2 = (g + 2) * 32 + lane;
        if (g + 2 < g1 && i2 < R) { const u32 rn = P.first_rec + i2; nn_ts = d.rstart[rn]; nn_te = d.te[rn]; }
      }
      cp_async_wait<1>();
      const u8 *b = slots + (size_t)(buf * 32 + lane) * tsz - (ts & ~15u); /* b[pos] is valid for the positions of this lane's title line */
      const bool walk = active && cur_fits;
      bool fields_ok = true;
      u32 bm = 0; /* fields tokenised in this block */
      TitleCursor cur; cur.init(b, ts, te, lut);
      for (u32 f = 0; f < nf;) {
        u32 fe = f + 1;
        if (!((touched >> f) & 1u)) { /* a run of fields that have matched record 0 everywhere so far */
          const u32 rest = touched >> f;
          fe = rest ? f + (u32)__ffs(rest) - 1u : nf;
          const u32 rl = S.off0[fe] - S.off0[f];
          const bool ok = walk && fields_ok;
          const bool same = ok && cur.pos + rl - 1 <= te && eq_bytes(b + cur.pos, S.r0 + S.off0[f], rl);
          if (__all_sync(0xFFFFFFFFu, same || !ok)) { if (same) cur.pos += rl; f = fe; continue; }
        }
        for (; f < fe; ++f) {
          const u32 len0 = S.len0[f];
          const u8 *d0 = S.r0 + S.off0[f];
          bool ok = walk && fields_ok;
          const bool same = ok && cur.pos + len0 <= te && eq_bytes(b + cur.pos, d0, len0 + 1);
          if (__all_sync(0xFFFFFFFFu, same || !ok)) { if (same) cur.pos += len0 + 1; continue; }
          touched |= 1u << f; bm |= 1u << f;
          Tok t; t.start = t.end = 0; t.v = 0; t.num = true;
          u32 len = 0;
          if (ok && !cur.next(t)) { fields_ok = false; ok = false; t.start = t.end = 0; t.v = 0; t.num = true; }
          if (ok) {
            len = t.end - t.start;
            const u32 m = len < len0 ? len : len0;
            const u8 *dp = b + t.start;
            if (len0 <= 32) { /* Hamming mask of the field in one register */
              const u32 mm = m ? neq_mask(dp, d0, m) : 0u;
              if (mm & ~S.mism[f][0]) atomicOr(&S.mism[f][0], mm);
            } else {
              u32 diff = 0;
              for (u32 p = 0; p < m; ++p) diff |= (u32)(dp[p] ^ d0[p]);
              if (diff)
                for (u32 p = 0; p < m; ++p)
                  if (dp[p] != d0[p]) { const u32 bit = 1u << (p & 31); if (p < (u32)MAXLEN0 && !(S.mism[f][p >> 5] & bit)) atomicOr(&S.mism[f][p >> 5], bit); }
            }
          }
          { /* the block's row of this field: a numeric token leaves its value, any other token its first characters */
            const size_t row = ((blk_base + g) * d.nfs + f) * 32 + lane;
            u32 rv = t.v;
            if (ok && !t.num) {
              rv = ld4u(b + t.start);
              if (len > 4 && len <= 8) d.tc[row] = ld4u(b + t.start + 4);
            }
            d.tv[row] = rv;
            d.tp[row] = ok ? ((t.start - ts) | (t.num ? TP_NUM : 0u) | (len << 16)) : TP_SAME;
          }
          const u32 inv_min = __reduce_max_sync(0xFFFFFFFFu, ok ? ~len : 0u);
          const u32 mx = __reduce_max_sync(0xFFFFFFFFu, ok ? len : 0u);
          const u32 nn = __ballot_sync(0xFFFFFFFFu, ok && !t.num);
          const u32 kv = key_of((i32)t.v);
          const u32 kmax = __reduce_max_sync(0xFFFFFFFFu, ok ? kv : 0u);
          const u32 kinv = __reduce_max_sync(0xFFFFFFFFu, ok ? ~kv : 0u);
          /* deltas: the record before lane 0 is the last record of the block before (none for the first block of the task) */
          u32 pv = __shfl_up_sync(0xFFFFFFFFu, t.v, 1);
          if (lane == 0 && g > g0) pv = ((prev_bm >> f) & 1u) ? lastv[w][f] : S.v0[f];
          const bool hasd = walk && (lane > 0 || g > g0);
          const u32 kd = key_of((i32)(t.v - pv));
          const u32 dmax = __reduce_max_sync(0xFFFFFFFFu, hasd ? kd : 0u);
          const u32 dinv = __reduce_max_sync(0xFFFFFFFFu, hasd ? ~kd : 0u);
          const u32 v_last = __shfl_sync(0xFFFFFFFFu, t.v, nrec - 1);
          if (lane == 0) {
            atomicMax(&S.facc[f][0], inv_min); atomicMax(&S.facc[f][1], mx);
            if (nn) S.facc[f][2] = 1;
            atomicMax(&S.facc[f][3], kmax); atomicMax(&S.facc[f][4], kinv);
            if (dmax) atomicMax(&S.facc[f][5], dmax);
            if (dinv) atomicMax(&S.facc[f][6], dinv);
            lastv[w][f] = v_last;
            if (g == g0) d.chunk_first[task_row + f] = t.v;
          }
        }
      }
      if (walk && (!fields_ok || cur.pos <= cur.lim)) err = E_FIELDS; /* fewer or more separators than record 0 */
      /* fields without a row in this block: every record carries record 0's token, so every delta inside the block is 0, and
       * so is the delta across the boundary to a block before that had no row either; a block before WITH a row ends in
       * lastv, which gives the boundary delta */
      {
        const u32 none = ~bm & nfmask;
        if (nrec >= 2) zero_d |= none;
        if (g > g0) {
          zero_d |= none & ~prev_bm;
          if (lane == 0)
            for (u32 m = none & prev_bm; m; m &= m - 1) {
              const u32 f = __ffs(m) - 1, kd = key_of((i32)(S.v0[f] - lastv[w][f]));
              atomicMax(&S.facc[f][5], kd); atomicMax(&S.facc[f][6], ~kd);
            }
        } else first_bm = bm;
      }
      if (lane == 0) d.blk_mask[blk_base + g] = bm;
      prev_bm = bm;
    }
    cp_async_wait<0>();
    __syncwarp();
    for (u32 f = lane; f < nf; f += 32) { /* what the task leaves for k_xdelta, and the zero deltas */
      if (!((first_bm >> f) & 1u)) d.chunk_first[task_row + f] = S.v0[f];
      d.chunk_last[task_row + f] = ((prev_bm >> f) & 1u) ? lastv[w][f] : S.v0[f];
      if ((zero_d >> f) & 1u) { atomicMax(&S.facc[f][5], key_of(0)); atomicMax(&S.facc[f][6], ~key_of(0)); }
    }
    if (!fits) err = err < (i32)E_UNSUPPORTED ? err : (i32)E_UNSUPPORTED; /* a title line longer than the slots */
    if (err) atomicMin(&S.err, err);
  }
  __syncthreads();
  /* flush to the subblock accumulators */
  SbAcc *A = d.acc + s;
  if (tid == 0 && S.err) atomicMin(&A->status, S.err);
  if (seed_ok)
    for (u32 i = tid; i < nf * 8; i += NT) {
      const u32 f = i >> 3, k = i & 7, v = S.facc[f][k];
      u32 *dst = &A->f[f].inv_min_len + k;
      if (v) atomicMax(dst, v);
    }
  if (seed_ok)
    for (u32 i = tid; i < nf * MASKW; i += NT) {
      const u32 f = i / MASKW, k = i % MASKW, v = S.mism[f][k];
      if (v) atomicOr(&A->f[f].mism[k], v);
    }
}

/* min / max of the numeric deltas that cross a task boundary of k_stat1 (tasks.cpp:149-166 runs over all records) */
__global__ void __launch_bounds__(128) k_xdelta(Dev d) {
  __shared__ u32 mx[MAXF], mn[MAXF], nf_s;
  const u32 s = blockIdx.x, tid = threadIdx.x;
  const SbPlan P = d.plans[s];
  SbAcc *A = d.acc + s;
  if (P.status || A->status) return;
  if (A->max_qlen + 1 > RAW_ROWS) { if (tid == 0) atomicMin(&A->status, (i32)E_UNSUPPORTED); return; } /* the raw per-position quality table has RAW_ROWS rows */
  const u32 ntask = (P.n_records + TASK_RECORDS - 1) / TASK_RECORDS;
  if (tid < MAXF) { mx[tid] = 0; mn[tid] = 0; }
  if (tid == 0) nf_s = min((u32)MAXF, count_seps(d.in, d.rstart[P.first_rec], d.te[P.first_rec]));
  __syncthreads();
  const u32 nf = nf_s;
  for (u32 i = tid; i < (ntask - 1) * nf; i += 128) {
    const u32 t = 1 + i / nf, f = i % nf;
    const u32 a = d.chunk_first[((size_t)P.chunk_base + 2u * t) * MAXF + f], b = d.chunk_last[((size_t)P.chunk_base + 2u * (t - 1)) * MAXF + f];
    const u32 kd = key_of((i32)(a - b));
    atomicMax(&mx[f], kd); atomicMax(&mn[f], ~kd);
  }
  __syncthreads();
  if (tid < nf) { if (mx[tid]) atomicMax(&A->f[tid].kmax_d, mx[tid]); if (mn[tid]) atomicMax(&A->f[tid].kinvmin_d, mn[tid]); }
}

/* ---- stat2 ---------------------------------------------------------------------------------------------------------- */
constexpr u32 S2W = 8;  /* warps per CTA */
constexpr u32 S2B = 8;  /* 32-record blocks per warp */

__global__ void __launch_bounds__(S2W * 32, 6) k_stat2(Dev d) {
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
  BlockRows cr = block_rows(d, P, s, g0 < g1 ? g0 : 0), prev = cr;
  if (g0 > 0 && g0 < g1) prev = block_rows(d, P, s, g0 - 1);
  for (u32 g = g0; g < g1; ++g) {
    const u32 nrec = min(32u, R - g * 32), rec = g * 32 + lane;
    const bool on = lane < nrec;
    const BlockRows nextb = g + 1 < g1 ? block_rows(d, P, s, g + 1) : cr; /* the next block's mask is on its way */
    const u32 ts = d.rstart[P.first_rec + min(rec, R - 1)];
    u32 flags = 0;
    for (u32 k = 0; k < nnc; ++k) {
      const u32 f = T.ncf[k];
      const FieldClass &F = T.fc[f];
      const RawRow row = load_row<false>(d, cr, prev, g > 0, f, lane, F.kind == K_NUM); /* fetching a field ahead did not pay here (measured) */
      if (F.kind == K_NUM) {
        const i32 v = on ? (i32)parsed_value(cr, f, row) : 0;
        i32 pv = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if (lane == 0 && g > 0) pv = (i32)parsed_prev_value(prev, f, row); /* last record of the block before */
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
      const TokRef t = parsed_token(cr, C, F, f, row, ts);
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
      TokRef t0; /* the block's first token */
      t0.len = __shfl_sync(0xFFFFFFFFu, t.len, 0); t0.off = __shfl_sync(0xFFFFFFFFu, t.off, 0);
      t0.inreg = __shfl_sync(0xFFFFFFFFu, (u32)t.inreg, 0) != 0; t0.same0 = false;
      t0.chars = (u64)__shfl_sync(0xFFFFFFFFu, (u32)t.chars, 0) | (u64)__shfl_sync(0xFFFFFFFFu, (u32)(t.chars >> 32), 0) << 32;
      bool pred = !on || t.len == t0.len;
      const bool cmp_reg = t.inreg && t0.inreg; /* both tokens in registers: one comparison */
      if (on && pred && cmp_reg) pred = ((t.chars ^ t0.chars) & (t.len >= 8 ? ~0ull : (1ull << (8 * t.len)) - 1ull)) == 0;
      const u32 maxlen = __reduce_max_sync(0xFFFFFFFFu, on ? t.len : 0u);
      for (u32 j = 0; j < maxlen; ++j) {
        const bool in_tok = on && j < t.len;
        const bool need = in_tok && (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u));
        const bool cmp = in_tok && pred && !cmp_reg;
        u32 ch = 0;
        if (need || cmp) ch = token_char(d, t, j);
        if (cmp && ch != token_char(d, t0, j)) pred = false;
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
    prev = cr; cr = nextb;
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
__device__ __forceinline__ void title_record_parsed(const Dev &d, const BlockRows &cr, const SbClass &C, const TitleTabs &T, const u32 *arena, u32 i, u32 ts,
                                                    u32 flags, bool first, Sink &s) {
  RawRow nx = load_row<true>(d, cr, cr, false, T.ncf[0], i, T.fc[T.ncf[0]].kind == K_NUM);
  for (u32 k = 0; k < C.nnc; ++k) {
    const u32 f = T.ncf[k];
    const FieldClass &F = T.fc[f];
    const bool flag = (flags >> f) & 1u;
    const RawRow row = nx;
    if (k + 1 < C.nnc) nx = load_row<true>(d, cr, cr, false, T.ncf[k + 1], i, T.fc[T.ncf[k + 1]].kind == K_NUM);
    if (F.kind == K_NUM) {
      const i32 v = (i32)parsed_value(cr, f, row);
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
    const TokRef t = parsed_token(cr, C, F, f, row, ts);
    if (!F.is_len_const) s.put(t.len - F.min_len, F.bits_len);
    const u16 *sm = (const u16 *)(arena + F.slotmap_off);
    for (u32 j = 0; j < t.len; ++j)
      if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) {
        const u32 tid = sm[j < 128 ? j : 128];
        const u64 e = ((const u64 *)(arena + C.chr_cl_off + (tid - C.tchr0) * 512u))[token_char(d, t, j)];
        s.put((u32)e, (u32)(e >> 32));
      }
  }
}

/* dynamic shared memory per warp: [32 * LPW_T staging words][CCW words][32 words for the info bits] */
__host__ __device__ __forceinline__ u32 enc_title_warp_bytes() { return (32u * LPW_T + CCW + 32u) * 4u; }

__global__ void __launch_bounds__(ENC_WARPS * 32, 4) k_enc_title(Dev d) {
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
      const BlockRows cr = block_rows(d, P, s, g);
      const u32 i = min(lane, nrec - 1);
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
