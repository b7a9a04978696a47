/*
 * phy_core.cuh -- the format logic of the phyNGSC subblock codec, written once as
 * __host__ __device__ functions.  The CUDA kernels (phy_kernels.cu) call these per thread / per warp;
 * tests/mirror/ compiles the very same header with g++ and strings the functions together serially so
 * that the bit-level logic can be checked against the oracle on a machine without a GPU.  (That
 * mirror is test infrastructure -- the library has no CPU path.)
 *
 * Reference lines each piece follows are cited at the piece (paths relative to the reference tree).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHY_HD __host__ __device__ __forceinline__
#define PHY_HDN __host__ __device__ __noinline__
#else
#define PHY_HD inline
#define PHY_HDN inline
#endif

namespace phy {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

constexpr int MAXF = 32;       /* title fields per record                                    */
constexpr int MAXLEN0 = 512;   /* longest token of record 0 (Hamming mask bits)              */
constexpr int MASKW = MAXLEN0 / 32;
constexpr int CHARPOS = 129;   /* per-position char tables 0..127 + the shared one (128)     */
constexpr int NUMH = 512;      /* numeric Huffman table only if the range is <= 512          */
constexpr int MAX_READ = 32767;/* kept-count + transfer flag packed in 16 bits               */
constexpr u32 NOTAB = 0xFFFFu;
constexpr u32 CHUNK_RECORDS = 128; /* records per GPU work item; offsets inside the streams are kept per chunk + local */

enum { K_CONST = 0, K_NUM = 1, K_STR = 2 };

enum {
  E_OK = 0, E_MALFORMED = -1, E_FIELDS = -2, E_COLORSPACE = -3, E_UNSUPPORTED = -4, E_CAPACITY = -5, E_CUDA = -6, E_ARG = -7
};

/* ---- small helpers ------------------------------------------------------------------------- */
/* separator set " ._,=:/-#\n" (phyNGSC.cpp:208) as a 96-bit membership mask */
PHY_HD bool is_sep(u8 c) {
  u32 k = c >> 5;
  u32 w = (k == 1) ? 0x2400F009u : (k == 0) ? 0x00000400u : (k == 2) ? 0x80000000u : 0u;
  return (w >> (c & 31)) & 1u;
}

/* trans_amb_codes, phyNGSC.cpp:184-206: ACGT = 1, YRWSKMDVHBNXU.- = 2..16, everything else 0 */
PHY_HD u32 amb_code(u8 c) {
  switch (c) {
    case 'A': case 'C': case 'G': case 'T': return 1;
    case 'Y': return 2; case 'R': return 3; case 'W': return 4; case 'S': return 5; case 'K': return 6;
    case 'M': return 7; case 'D': return 8; case 'V': return 9; case 'H': return 10; case 'B': return 11;
    case 'N': return 12; case 'X': return 13; case 'U': return 14; case '.': return 15; case '-': return 16;
    default: return 0;
  }
}
PHY_HD bool is_acgt(u8 c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
/* quality byte that carries a transferred ambiguity code (phyNGSC.cpp:575-580) */
PHY_HD u8 xfer_qual(u32 a, u8 q) { return (u8)(128u + (a << 3) - 16u + (u32)(q - 33)); }

/* BitStream::BitLength (bit_stream.h:268-277) of an int32 difference widened like the callers do */
PHY_HD u32 bit_length_i32(i32 d) {
  if (d < 0) return 64;
  u32 x = (u32)d;
  for (u32 i = 0; i < 32; ++i) if (x < (1u << i)) return i;
  return 64;
}
PHY_HD u32 bit_length_u32(u32 x) {
  for (u32 i = 0; i < 32; ++i) if (x < (1u << i)) return i;
  return 64;
}
PHY_HD i32 wsub(i32 a, i32 b) { return (i32)((u32)a - (u32)b); }
/* order-preserving map int32 -> uint32 so that signed min/max become unsigned atomicMax on zeroed memory */
PHY_HD u32 key_of(i32 v) { return (u32)v ^ 0x80000000u; }
PHY_HD i32 val_of(u32 k) { return (i32)(k ^ 0x80000000u); }
PHY_HD u32 bswap32(u32 x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24);
#endif
}
PHY_HD u32 align_up(u32 x, u32 a) { return (x + a - 1) / a * a; }

/* bits needed for a leaf id in the serialised tree (huffman.cpp:96-98) */
PHY_HD u32 tree_id_bits(u32 n) {
  u32 b = 0;
  for (u32 t = 2; t <= n; t *= 2) ++b;
  if (n & (n - 1)) ++b;
  return b;
}
/* upper bound of [word mem_size][mem] for an n-symbol table */
PHY_HD u32 tree_blob_cap(u32 n) { return 4u + 9u + ((n - 1) + n * (1 + tree_id_bits(n)) + 7) / 8; }

/* ---- bit sinks ----------------------------------------------------------------------------- */
/* Bits are appended MSB-first and the byte stream is what BitStream produces (bit_stream.h:80-265):
 * stream bit i lives in byte i/8 at bit 7-(i%8).  The output buffer is zero-initialised and viewed
 * as 32-bit words; a writer ORs its bits in.  Words it covers completely are stored plainly, the
 * first and last (possibly shared with the neighbouring writer) are OR-ed atomically. */
struct CountSink {
  u64 bits;
  PHY_HD void init() { bits = 0; }
  PHY_HD void put(u32, u32 n) { bits += n; }
};

struct OrSink {
  u32 *w;     /* word pointer of the next word to flush                                    */
  u64 acc;    /* pending bits, left-aligned: the next stream bit is bit 63                 */
  u32 fill;   /* number of pending bits incl. the leading pad of the first word (< 32)     */
  bool shared_first;
  bool live;  /* false: walk without storing (lanes that only keep a warp converged)       */
  PHY_HD void init(u32 *words, u64 bitpos, bool live_ = true) {
    w = words + (bitpos >> 5);
    fill = (u32)(bitpos & 31);
    acc = 0;
    shared_first = fill != 0;
    live = live_;
  }
  PHY_HD void flush_word(u32 word, bool shared) {
    if (live) {
      u32 v = bswap32(word);
#if defined(__CUDA_ARCH__)
      if (shared) atomicOr(w, v); else *w = v;
#else
      (void)shared; *w |= v;
#endif
    }
    ++w;
  }
  PHY_HD void put(u32 v, u32 n) { /* n <= 32, v < 2^n */
    if (n == 0) return;
    acc |= (u64)v << (64 - fill - n); /* shift in [1, 63] */
    fill += n;
    if (fill >= 32) {
      flush_word((u32)(acc >> 32), shared_first);
      shared_first = false;
      acc <<= 32;
      fill -= 32;
    }
  }
  PHY_HD void finish() {
    if (fill) { flush_word((u32)(acc >> 32), true); fill = 0; acc = 0; }
  }
};

/* ---- titles: tokeniser (phyNGSC.cpp:342-423) ------------------------------------------------- */
struct Tok { u32 start, end, v; bool num; };

PHY_HD u32 count_seps(const u8 *b, u32 ts, u32 te) {
  u32 n = 0;
  for (u32 i = ts; i <= te; ++i) n += is_sep(b[i]) ? 1u : 0u;
  return n;
}

/* character classes for the tokeniser: bit 0 = separator, bit 1 = decimal digit */
PHY_HD void fill_char_lut(u8 *lut) {
  for (u32 c = 0; c < 256; ++c) lut[c] = (u8)((is_sep((u8)c) ? 1u : 0u) | ((c >= '0' && c <= '9') ? 2u : 0u));
}

/* Walks the separators of one title line [ts, te] (te = its '\n').  next() yields the token before the
 * next separator together with utils::is_num / utils::to_num of it (utils.h:107-125); skip(len) steps over
 * a token whose length is already known (constant fields). */
struct TitleCursor {
  const u8 *b; const u8 *lut; u32 pos, lim;
  PHY_HD void init(const u8 *base, u32 ts, u32 te, const u8 *char_lut) { b = base; pos = ts; lim = te; lut = char_lut; }
  PHY_HD void skip(u32 len) { pos += len + 1; }
  PHY_HD bool next(Tok &t) {
    if (pos > lim) return false;
    u32 i = pos, v = 0;
    u32 alld = 2;
    u8 c0 = b[i];
    for (;; ++i) {
      u8 c = b[i];
      u32 fl = lut[c];
      if (fl & 1u) break;
      alld &= fl;
      v = v * 10u + ((u32)c - '0');
    }
    u32 len = i - pos;
    t.start = pos; t.end = i; t.v = v;
    t.num = alld != 0 && len >= 1 && (len == 1 || c0 != '0');
    pos = i + 1;
    return true;
  }
};

/* ---- per-subblock accumulators (zero-initialised, updated with atomicMax / atomicOr / atomicAdd) - */
struct FieldAcc {
  u32 inv_min_len, max_len, not_num, kmax_v, kinvmin_v, kmax_d, kinvmin_d, pad;
  u32 mism[MASKW]; /* bit p: some record differs from record 0 at position p of this field */
};
struct SbAcc {
  i32 status; u32 warnings;
  u32 max_qlen, max_slen;
  u32 inv_min_qlen, pad0; /* ~(shortest read) */
  u32 qpresent[8];
  u32 dna_occ[256];  /* non-zero = symbol occurs in the (compacted) DNA; exact counts are taken later, only if needed */
  FieldAcc f[MAXF];
};

/* ---- per-subblock classification (output of classify_subblock, read by every later stage) ----- */
struct FieldClass {
  u32 off0, len0, min_len, max_len;
  i32 min_v, max_v, min_d, max_d, base;
  u32 diff, bits_num, bits_val, bits_len;
  u32 tab;          /* table id of the numeric Huffman table                   */
  u32 cl_off;       /* arena word offset of that table's code array            */
  u32 freq_off;     /* ... and of its frequency array                          */
  u32 slotmap_off;  /* arena word offset of u16[CHARPOS(+pad)] table ids       */
  u8 sep, kind, is_delta, has_table, is_len_const, pad[3];
  u32 mism[MASKW];
};
struct TableDesc { u32 n, freq_off, cl_off, tree_off, tree_len, dst, maxlen; }; /* maxlen: longest code of the built table */

struct SbClass {
  i32 status; u32 R, nf, P, nnc;
  u32 max_qlen, max_slen, nsym, nq, plain, flags, nb_len;
  u32 varlen;                  /* reads differ in length (picks the walker variant on the GPU; not part of the format) */
  u32 ts0, te0;                /* title line of record 0 (batch-relative positions)                   */
  /* arena layout (word offsets unless stated) */
  u32 ntab, tabdesc_off, tq0, tdna, tchr0, qstat_off, dnastat_off, zero_begin, zero_end;
  u32 chr_cl_off;              /* code array of char table tchr0; table tchr0 + k follows 512 * k words later */
  u32 chr_freq_off;            /* frequency array of char table tchr0; table tchr0 + k follows 256 * k words later */
  u32 qpk_esc;                 /* some packed entry is the escape (set by the Huffman stage) */
  u32 qpk_off, qpk_bad;        /* quality tables packed to 16 bits (len << 12 | code, escape for longer codes); qpk_bad != 0 when a table failed to build */
  u32 nblk, flagbits_off;
  u32 nchunk;                  /* 128-record chunks (work items of the statistics kernels) */
  u32 blk3_off;                /* per 32-record block: [3][nblk] totals -> bases of quality bits, dna bits, title bytes */
  u32 stage_off;               /* header staging: title | quality | dna header bytes                  */
  u32 thdr_cap, qhdr_cap, dhdr_cap;
  u32 arena_used;
  /* layout results */
  u32 info_len, thdr_len, title_len, qhdr_len, qual_len, dhdr_len, dna_len, payload_len;
  u64 qbits_total, dbits_total;
  u64 out_off;
  /* single-walk encoder (phy_fast.cuh): tasks of 256 records with slots in the temporary buffer */
  u64 tmp_base;                /* first word of this subblock's region of the temporary buffer                */
  u32 fast;                    /* 1: encoded by the single-walk kernels, 0: by the two-walk kernels           */
  u32 ntask, task_off;         /* arena: [3][ntask] quality bits / DNA bits / title bytes per task, then [3][ntask] their exclusive scans */
  u32 strd_q, strd_d, strd_t;  /* slot sizes in words                                                         */
  u32 info_words, pad_fast;
  u8 symbols[256], quals[256], sym_code[256], qua_code[256];
  u8 ncf[MAXF];                /* the nnc non-constant fields in title order ...                              */
  u16 ncskip[MAXF];            /* ... and the bytes of constant tokens (with their separators) in front of each */
  FieldClass f[MAXF];
};

/* info header size: three words, three bytes, one word (phyNGSC.cpp:719-730) */
constexpr u32 INFO_FIXED = 19;

/* ---- Huffman (huffman.cpp:18-118, 191-205; huffman.h:57-60, 134-147) ----------------------- */
/* One table is built by `nl` cooperating lanes (a warp on the GPU, 1 on the host).  Phases that are
 * inherently serial run on lane 0.  SYNC() is __syncwarp on the device and nothing on the host. */
template <int N> /* N >= number of symbols; the last slot of in_f doubles as the warp's broadcast word */
struct HuffScratchT {
  static constexpr u32 CAP = N;
  u32 key_f[N];
  u16 key_id[N];
  u32 in_f[N];
  u16 left[N], right[N];
  u32 code[2 * N];
  u8 len[2 * N];
};
typedef HuffScratchT<512> HuffScratch;     /* any table of the format (numeric tables reach 512 symbols) */
typedef HuffScratchT<64> HuffScratchSmall; /* quality and DNA tables: eight times as many tables fit an SM's shared memory */

struct NoSync { PHY_HD void operator()() const {} };

/* freq[n] -> cl[n] (len << 32 | code) and tree blob [word mem_size][mem] at `tree`; returns blob bytes,
 * or 0 when a code would exceed 32 bits (the reference's 32-bit code word would overflow). */
template <class Scratch, class Sync>
PHY_HD u32 huff_table(const u32 *freq, u32 n, u64 *cl, u8 *tree, Scratch &S, u32 lane, u32 nl, Sync sync) {
  constexpr u32 BC = Scratch::CAP - 1; /* broadcast slot: the queue holds at most n - 1 <= CAP - 1 internal nodes, index <= CAP - 2 */
  /* rank sort by the strict total order (frequency, id), huffman.h:57-60 */
  for (u32 i = lane; i < n; i += nl) S.code[i] = freq[i];
  sync();
  u32 myzeros = 0;
  for (u32 i = lane; i < n; i += nl) {
    u32 fi = S.code[i], rank = 0;
    for (u32 j = 0; j < n; ++j) {
      u32 fj = S.code[j];
      rank += (fj < fi || (fj == fi && j < i)) ? 1u : 0u;
    }
    S.key_f[rank] = fi; S.key_id[rank] = (u16)i;
    myzeros += fi == 0 ? 1u : 0u;
  }
  (void)myzeros;
  sync();
  /* zero-frequency compaction, huffman.cpp:44-50: drop smallest while more than two remain */
  u32 lo = 0;
  if (lane == 0) {
    while (n - lo > 2 && S.key_f[lo] == 0) ++lo;
    S.in_f[BC] = lo;
  }
  sync();
  lo = S.in_f[BC];
  sync();
  u32 p = n - lo;
  for (u32 i = lane; i < 2 * n; i += nl) { if (i < 2 * Scratch::CAP) { S.code[i] = 0; S.len[i] = 0; } }
  sync();
  u32 ok = 1;
  if (lane == 0) {
    /* merge loop, huffman.cpp:57-70, as sorted leaves + FIFO of internal nodes (leaf wins ties) */
    u32 li = lo, qi = 0, qn = 0;
    for (u32 i = 0; i + 1 < p; ++i) {
      u32 pf[2], pid[2];
      for (int k = 0; k < 2; ++k) {
        bool take_leaf;
        if (li < n && qi < qn) take_leaf = S.key_f[li] <= S.in_f[qi]; /* equal frequency: leaf id < internal id */
        else take_leaf = li < n;
        if (take_leaf) { pf[k] = S.key_f[li]; pid[k] = S.key_id[li]; ++li; }
        else { pf[k] = S.in_f[qi]; pid[k] = n + qi; ++qi; }
      }
      S.in_f[qn] = pf[0] + pf[1];
      S.left[qn] = (u16)pid[0]; S.right[qn] = (u16)pid[1];
      ++qn;
    }
    /* codes, huffman.cpp:73-79: root = n+p-2, left appends 0, right appends 1 */
    if (p >= 2) {
      for (u32 i = n + p - 2;; --i) {
        u32 l = S.left[i - n], r = S.right[i - n], d = (u32)S.len[i] + 1;
        if (d > 32) { ok = 0; break; }
        S.len[l] = S.len[r] = (u8)d;
        S.code[l] = S.code[i] << 1;
        S.code[r] = (S.code[i] << 1) | 1u;
        if (i == n) break;
      }
    }
    S.in_f[BC] = ok;
  }
  sync();
  ok = S.in_f[BC];
  if (!ok) return 0;
  for (u32 i = lane; i < n; i += nl) cl[i] = ((u64)S.len[i] << 32) | S.code[i];
  /* serialisation, huffman.cpp:88-118 + huffman.h:134-147 + huffman.cpp:191-205 */
  u32 blob = 0;
  if (lane == 0) {
    u32 root = n + p - 2, idb = tree_id_bits(n), min_len = n;
    for (u32 i = 0; i < n; ++i) if (S.len[i] > 0 && S.len[i] < min_len) min_len = S.len[i];
    u8 *m = tree + 4;
    m[0] = (u8)(root >> 24); m[1] = (u8)(root >> 16); m[2] = (u8)(root >> 8); m[3] = (u8)root;
    m[4] = (u8)(n >> 24); m[5] = (u8)(n >> 16); m[6] = (u8)(n >> 8); m[7] = (u8)n;
    m[8] = (u8)min_len;
    u32 o = 9, acc = 0, nb = 0;
    u16 *stack = S.key_id; /* sorted keys are dead after the merge */
    u32 sp = 0;
    stack[sp++] = (u16)root;
    while (sp) {
      u32 v = stack[--sp];
      u32 bits, val;
      if (v < n) { bits = 1 + idb; val = (1u << idb) | v; }
      else { bits = 1; val = 0; stack[sp++] = S.right[v - n]; stack[sp++] = S.left[v - n]; }
      acc = (acc << bits) | val; nb += bits; /* bits <= 10, nb < 8 before -> fits */
      while (nb >= 8) { nb -= 8; m[o++] = (u8)(acc >> nb); }
      acc &= (1u << nb) - 1u;
    }
    if (nb) m[o++] = (u8)(acc << (8 - nb));
    tree[0] = (u8)(o >> 24); tree[1] = (u8)(o >> 16); tree[2] = (u8)(o >> 8); tree[3] = (u8)o;
    blob = o + 4;
    S.in_f[BC] = blob;
  }
  sync();
  blob = S.in_f[BC];
  sync();
  return blob;
}

/* ---- classification of one subblock (serial) ------------------------------------------------ */
/* Turns the reduced statistics into the coding decisions of AnalyzeTitleFields' tail (tasks.cpp:196-222),
 * AnalyzeDNA (tasks.cpp:226-257), the symbol maps (phyNGSC.cpp:659-687) and lays out the subblock's
 * scratch arena.  `b` indexes the batch input; arena is this subblock's word arena. */
struct ArenaAlloc {
  u32 used, cap; bool over;
  PHY_HD u32 take(u32 words) { u32 o = used; used += words; if (used > cap) over = true; return o; }
};

PHY_HDN void classify_subblock(const u8 *b, const u8 *lut, const SbAcc &A, u32 R, u32 ts0, u32 te0, u32 *arena, u32 arena_words, SbClass &C) {
  C.status = A.status; C.R = R; C.ts0 = ts0; C.te0 = te0;
  if (C.status) return;
  /* DNA symbols / quality alphabet ascending, phyNGSC.cpp:669-686 */
  u32 nsym = 0, nq = 0;
  for (u32 c = 0; c < 256; ++c) {
    C.sym_code[c] = 0; C.qua_code[c] = 0;
    if (A.dna_occ[c]) { C.sym_code[c] = (u8)nsym; C.symbols[nsym++] = (u8)c; }
    if ((A.qpresent[c >> 5] >> (c & 31)) & 1u) { C.qua_code[c] = (u8)nq; C.quals[nq++] = (u8)c; }
  }
  if (nsym == 0 || nq == 0) { C.status = E_UNSUPPORTED; return; }
  C.nsym = nsym; C.nq = nq;
  C.plain = nsym <= 4 ? 1u : 0u;                      /* tasks.cpp:239-256 (frequency test is dead code) */
  C.flags = 0x8u | 0x4u | 0x20u | 0x80u | (C.plain ? 0x2u : 0u); /* 0xAE / 0xAC */
  C.max_qlen = A.max_qlen; C.max_slen = A.max_slen;
  C.varlen = (~A.inv_min_qlen != A.max_qlen) ? 1u : 0u;
  C.nb_len = bit_length_u32(C.max_qlen);
  if (C.nb_len > 32) { C.status = E_UNSUPPORTED; return; }
  C.info_len = INFO_FIXED + (u32)(((u64)R * C.nb_len + 7) / 8);

  ArenaAlloc al; al.used = 0; al.cap = arena_words; al.over = false;
  /* histograms first (one contiguous zeroed range) */
  C.zero_begin = al.used;
  C.qstat_off = al.take((C.max_qlen + 1) * nq);

  /* title fields, seeded from record 0 (phyNGSC.cpp:345-379) */
  TitleCursor cur; cur.init(b, ts0, te0, lut);
  Tok t; u32 nf = 0;
  while (cur.next(t)) {
    if (nf < (u32)MAXF) { C.f[nf].off0 = t.start - ts0; C.f[nf].len0 = t.end - t.start; C.f[nf].sep = b[t.end]; }
    ++nf;
  }
  if (nf == 0 || nf > (u32)MAXF) { C.status = E_UNSUPPORTED; return; }
  C.nf = nf;
  u32 P = 0; /* libstdc++ vector growth wipes the seeded value histogram of fields below P */
  if (nf >= 2) { P = 1; while (P * 2 <= nf - 1) P *= 2; }
  C.P = P;
  u32 nnc = 0, ntab_num = 0, ntab_chr = 0;
  for (u32 f = 0; f < nf; ++f) {
    FieldClass &F = C.f[f];
    const FieldAcc &a = A.f[f];
    if (F.len0 > (u32)MAXLEN0) { C.status = E_UNSUPPORTED; return; }
    F.min_len = ~a.inv_min_len; F.max_len = a.max_len;
    for (int k = 0; k < MASKW; ++k) F.mism[k] = a.mism[k];
    bool any_mism = false;
    for (int k = 0; k < MASKW; ++k) any_mism = any_mism || a.mism[k] != 0;
    F.is_len_const = (F.min_len == F.len0 && F.max_len == F.len0) ? 1 : 0;
    F.has_table = 0; F.is_delta = 0; F.tab = NOTAB; F.cl_off = 0; F.freq_off = 0; F.slotmap_off = 0; F.diff = 0; F.base = 0;
    F.bits_num = F.bits_val = F.bits_len = 0;
    F.min_v = F.max_v = F.min_d = F.max_d = 0;
    if (F.is_len_const && !any_mism) { F.kind = K_CONST; continue; }
    ++nnc;
    if (!a.not_num) {
      F.kind = K_NUM;
      F.min_v = val_of(~a.kinvmin_v); F.max_v = val_of(a.kmax_v);
      if (R >= 2) { F.min_d = val_of(~a.kinvmin_d); F.max_d = val_of(a.kmax_d); }
      else { F.min_d = 1; F.max_d = -1; } /* Field::Field defaults, structures.h:103-106 */
      i32 vr = wsub(F.max_v, F.min_v), dr = wsub(F.max_d, F.min_d);
      F.is_delta = !(vr < dr) ? 1 : 0; /* tasks.cpp:208-217 */
      F.bits_num = bit_length_i32(F.is_delta ? dr : vr);
      F.bits_val = bit_length_i32(vr);
      i32 diff = (F.is_delta ? dr : vr) + 1;
      F.base = F.is_delta ? F.min_d : F.min_v;
      bool nonempty = F.is_delta ? (R >= 2) : (f >= P);
      if (diff <= NUMH && nonempty) { /* tasks.cpp:338 */
        if (diff <= 0) { C.status = E_UNSUPPORTED; return; }
        F.has_table = 1; F.diff = (u32)diff; ++ntab_num;
      }
      if (F.bits_val > 32 || (!F.has_table && F.bits_num > 32)) { C.status = E_UNSUPPORTED; return; }
    } else {
      F.kind = K_STR;
      if (F.max_len == 128) { C.status = E_UNSUPPORTED; return; } /* SURVEY Q11 */
      F.bits_len = bit_length_u32(F.max_len - F.min_len);
      u32 ntab = F.max_len < 128 ? F.max_len : 128;
      for (u32 j = 0; j < ntab; ++j) if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) ++ntab_chr;
      if (F.max_len >= 128) ++ntab_chr;
    }
  }
  C.nnc = nnc;
  { /* walkers visit only the non-constant fields: constant tokens in between are stepped over in one addition */
    u32 k = 0, skip = 0;
    for (u32 f = 0; f < nf; ++f) {
      if (C.f[f].kind == K_CONST) { skip += C.f[f].len0 + 1; continue; }
      C.ncf[k] = (u8)f; C.ncskip[k] = (u16)skip; ++k; skip = 0;
    }
  }
  /* numeric + char histograms */
  u32 numhist_off[MAXF];
  for (u32 f = 0; f < nf; ++f) numhist_off[f] = (C.f[f].kind == K_NUM && C.f[f].has_table) ? al.take(C.f[f].diff) : 0;
  u32 chrhist_off = al.take(ntab_chr * 256);
  u32 dnastat_off = al.take(nsym); /* exact symbol counts, only filled (by a later pass) when the DNA is Huffman coded */
  C.dnastat_off = dnastat_off;
  C.zero_end = al.used;
  while (al.used & 3u) al.take(1); /* 16-byte alignment: the packed tables are copied to shared memory as vectors */
  C.qpk_off = al.take(((C.max_qlen + 1) * nq + 1) / 2 + 4);
  C.qpk_bad = 0; C.qpk_esc = 0;
  /* table directory: quality (max_qlen+1), dna (0/1), numeric, char */
  u32 ntab = (C.max_qlen + 1) + (C.plain ? 0 : 1) + ntab_num + ntab_chr;
  C.ntab = ntab;
  C.tabdesc_off = al.take(ntab * (u32)(sizeof(TableDesc) / 4));
  for (u32 f = 0; f < nf; ++f) if (C.f[f].kind == K_STR) C.f[f].slotmap_off = al.take((CHARPOS + 1) / 2 + 1);
  C.nblk = (R + 31) / 32;
  C.flagbits_off = al.take(C.nblk);
  C.nchunk = (R + CHUNK_RECORDS - 1) / CHUNK_RECORDS;
  C.blk3_off = al.take(3 * C.nblk);
  C.ntask = (R + 255) / 256;
  C.task_off = al.take(6 * C.ntask);
  C.fast = 0; C.tmp_base = 0; C.strd_q = C.strd_d = C.strd_t = C.info_words = 0; C.pad_fast = 0;
  if (al.used & 1) al.take(1); /* 8-byte alignment for the code tables */
  u32 cl_off = al.used;
  u64 cl_words = 2ull * ((u64)(C.max_qlen + 1) * nq + (C.plain ? 0 : nsym) + (u64)ntab_chr * 256);
  for (u32 f = 0; f < nf; ++f) if (C.f[f].kind == K_NUM && C.f[f].has_table) cl_words += 2ull * C.f[f].diff;
  if (cl_words > arena_words) { C.status = E_CAPACITY; return; }
  al.take((u32)cl_words);
  u32 tcap_q = align_up(tree_blob_cap(nq), 4) / 4, tcap_c = align_up(tree_blob_cap(256), 4) / 4;
  u32 tree_off = al.used;
  {
    u64 tw = (u64)(C.max_qlen + 1) * tcap_q + (C.plain ? 0 : align_up(tree_blob_cap(nsym), 4) / 4) + (u64)ntab_chr * tcap_c;
    for (u32 f = 0; f < nf; ++f) if (C.f[f].kind == K_NUM && C.f[f].has_table) tw += align_up(tree_blob_cap(C.f[f].diff), 4) / 4;
    if (tw > arena_words) { C.status = E_CAPACITY; return; }
    al.take((u32)tw);
  }
  /* header staging caps */
  u32 thdr = 4;
  for (u32 f = 0; f < nf; ++f) {
    const FieldClass &F = C.f[f];
    thdr += 2;
    if (F.kind == K_CONST) thdr += 4 + F.len0;
    else if (F.kind == K_NUM) thdr += 1 + 16 + (F.has_table ? tree_blob_cap(F.diff) : 0);
    else {
      thdr += 1 + 1 + 12 + F.len0 + (F.len0 + 7) / 8;
      u32 nt = F.max_len < 128 ? F.max_len : 128, k = 0;
      for (u32 j = 0; j < nt; ++j) if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) ++k;
      if (F.max_len >= 128) ++k;
      thdr += k * tree_blob_cap(256);
    }
  }
  C.thdr_cap = align_up(thdr, 16);
  C.qhdr_cap = align_up(nq + (C.max_qlen + 1) * tree_blob_cap(nq), 16);
  C.dhdr_cap = align_up(nsym + (C.plain ? 0 : tree_blob_cap(nsym)), 16);
  C.stage_off = al.take((C.thdr_cap + C.qhdr_cap + C.dhdr_cap) / 4);
  C.arena_used = al.used;
  if (al.over) { C.status = E_CAPACITY; return; }

  /* fill the directory */
  TableDesc *td = (TableDesc *)(arena + C.tabdesc_off);
  u32 tid = 0;
  C.tq0 = 0;
  for (u32 p = 0; p <= C.max_qlen; ++p, ++tid) {
    td[tid].n = nq; td[tid].freq_off = C.qstat_off + p * nq; td[tid].cl_off = cl_off; td[tid].tree_off = tree_off;
    td[tid].tree_len = 0; td[tid].dst = 0; td[tid].maxlen = 0;
    cl_off += 2 * nq; tree_off += tcap_q;
  }
  C.tdna = NOTAB;
  if (!C.plain) {
    /* sym_stats (tasks.cpp:233-236) are counted into arena[dnastat_off + sym_code] after the zeroing pass */
    C.tdna = tid;
    td[tid].n = nsym; td[tid].freq_off = dnastat_off; td[tid].cl_off = cl_off; td[tid].tree_off = tree_off;
    td[tid].tree_len = 0; td[tid].dst = 0; td[tid].maxlen = 0;
    cl_off += 2 * nsym; tree_off += align_up(tree_blob_cap(nsym), 4) / 4; ++tid;
  }
  for (u32 f = 0; f < nf; ++f) {
    FieldClass &F = C.f[f];
    if (F.kind == K_NUM && F.has_table) {
      F.tab = tid; F.cl_off = cl_off; F.freq_off = numhist_off[f];
      td[tid].n = F.diff; td[tid].freq_off = numhist_off[f]; td[tid].cl_off = cl_off; td[tid].tree_off = tree_off;
      td[tid].tree_len = 0; td[tid].dst = 0; td[tid].maxlen = 0;
      cl_off += 2 * F.diff; tree_off += align_up(tree_blob_cap(F.diff), 4) / 4; ++tid;
    }
  }
  u32 slot = 0;
  C.tchr0 = tid; /* char tables are numbered consecutively from here */
  C.chr_cl_off = cl_off; C.chr_freq_off = chrhist_off;
  for (u32 f = 0; f < nf; ++f) {
    FieldClass &F = C.f[f];
    if (F.kind != K_STR) continue;
    u16 *sm = (u16 *)(arena + F.slotmap_off);
    u32 nt = F.max_len < 128 ? F.max_len : 128;
    for (u32 j = 0; j < (u32)CHARPOS; ++j) {
      bool need = (j < nt) ? (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) : (j == 128 && F.max_len >= 128);
      if (!need) { sm[j] = (u16)NOTAB; continue; }
      sm[j] = (u16)tid;
      td[tid].n = 256; td[tid].freq_off = chrhist_off + slot * 256; td[tid].cl_off = cl_off; td[tid].tree_off = tree_off;
      td[tid].tree_len = 0; td[tid].dst = 0; td[tid].maxlen = 0;
      cl_off += 512; tree_off += tcap_c; ++tid; ++slot;
    }
  }
}

/* ---- per-record sequence / quality statistics (phyNGSC.cpp:462-619) ----------------------------- */
struct SeqStat { u32 xfer, kept, err; u32 acgt[4]; bool others; };

/* Decides the ambiguity transfer for one record (phyNGSC.cpp:549-588) and counts A/C/G/T.  Symbols
 * other than ACGT that stay in the DNA string are reported through `other(c)`, every quality byte the
 * record will present to the quality coder through `seen(q)`. */
template <class Other, class Seen>
PHY_HD void seqqual_stat(const u8 *b, u32 ss, u32 L, u32 qs, SeqStat &o, Other other, Seen seen) {
  o.acgt[0] = o.acgt[1] = o.acgt[2] = o.acgt[3] = 0; o.err = 0;
  bool any = false, ok = true, nul = false;
  u32 namb = 0;
  for (u32 j = 0; j < L; ++j) {
    u8 c = b[ss + j], q = b[qs + j];
    nul = nul || c == 0 || q == 0;
    if (c == 'A') ++o.acgt[0]; else if (c == 'C') ++o.acgt[1]; else if (c == 'G') ++o.acgt[2]; else if (c == 'T') ++o.acgt[3];
    else {
      u32 a = amb_code(c);
      if (a == 0 || q < 33 || q > 40) ok = false;
      any = true; ++namb;
    }
  }
  o.xfer = (any && ok) ? 1u : 0u;
  o.kept = o.xfer ? L - namb : L;
  o.others = any;
  if (nul) o.err = 1;
  if (any) {
    for (u32 j = 0; j < L; ++j) {
      u8 c = b[ss + j], q = b[qs + j];
      if (is_acgt(c)) { seen(q); continue; }
      if (o.xfer) seen(xfer_qual(amb_code(c), q));
      else { other(c); seen(q); }
    }
  } else {
    for (u32 j = 0; j < L; ++j) seen(b[qs + j]);
  }
}

/* ---- per-record stream walkers ----------------------------------------------------------------- */
/* Quality table lookups: (table, symbol code) -> (code, len).  QFull reads the 64-bit entries the Huffman stage
 * wrote; QPacked reads a 16-bit copy (len << 12 | code, valid when no code is longer than 12 bits) that the GPU
 * kernels keep in shared memory. */
struct QFull {
  const u64 *cl; u32 nq;
  PHY_HD void get(u32 table, u32 c, u32 &code, u32 &len) const { u64 e = cl[table * nq + c]; code = (u32)e; len = (u32)(e >> 32); }
};
struct QPacked {
  const u16 *pk; u32 nq;
  PHY_HD void get(u32 table, u32 c, u32 &code, u32 &len) const { u32 e = pk[table * nq + c]; code = e & 0xFFFu; len = e >> 12; }
};
PHY_HD bool qpack_entry(u64 e, u16 &out) {
  u32 len = (u32)(e >> 32), code = (u32)e;
  if (len > 12) { out = 0; return false; }
  out = (u16)((len << 12) | code);
  return true;
}

/* Quality codes of one record: position k uses table k+1 (tasks.cpp:609-619).  The table entries of the next
 * three positions are already in flight while one is appended (software pipeline of depth 3), because the
 * lookup chain byte -> symbol code -> table entry is what a thread otherwise waits for. */
template <class Sink, class Q>
PHY_HD void quality_record(const u8 *b, u32 ss, u32 L, u32 qs, bool xfer, const u8 *qua_code, const Q &tab, Sink &s) {
  const u8 *qp = b + qs, *sp = b + ss;
  if (xfer) {
    for (u32 j = 0; j < L; ++j) {
      u8 q = qp[j];
      u32 a = amb_code(sp[j]);
      if (a > 1) q = xfer_qual(a, q);
      u32 code, len;
      tab.get(j + 1, qua_code[q], code, len);
      s.put(code, len);
    }
    return;
  }
  u32 c0 = 0, l0 = 0, c1 = 0, l1 = 0, c2 = 0, l2 = 0;
  if (0 < L) tab.get(1, qua_code[qp[0]], c0, l0);
  if (1 < L) tab.get(2, qua_code[qp[1]], c1, l1);
  if (2 < L) tab.get(3, qua_code[qp[2]], c2, l2);
  u32 j = 0;
  for (; j + 3 <= L; j += 3) {
    u32 n0 = 0, m0 = 0, n1 = 0, m1 = 0, n2 = 0, m2 = 0;
    if (j + 3 < L) tab.get(j + 4, qua_code[qp[j + 3]], n0, m0);
    s.put(c0, l0);
    if (j + 4 < L) tab.get(j + 5, qua_code[qp[j + 4]], n1, m1);
    s.put(c1, l1);
    if (j + 5 < L) tab.get(j + 6, qua_code[qp[j + 5]], n2, m2);
    s.put(c2, l2);
    c0 = n0; l0 = m0; c1 = n1; l1 = m1; c2 = n2; l2 = m2;
  }
  if (j < L) s.put(c0, l0);
  if (j + 1 < L) s.put(c1, l1);
}

template <class Sink, class Q>
PHY_HD void quality_record_simple(const u8 *b, u32 ss, u32 L, u32 qs, bool xfer, const u8 *qua_code, const Q &tab, Sink &s) {
  const u8 *qp = b + qs, *sp = b + ss;
  for (u32 j = 0; j < L; ++j) {
    u8 q = qp[j];
    if (xfer) { u32 a = amb_code(sp[j]); if (a > 1) q = xfer_qual(a, q); }
    u32 code, len;
    tab.get(j + 1, qua_code[q], code, len);
    s.put(code, len);
  }
}

/* DNA codes of one record (tasks.cpp:544-557): 2 bits per kept base (sixteen at a time) or its Huffman code. */
template <class Sink>
PHY_HD void dna_record(const u8 *b, u32 ss, u32 L, bool xfer, bool plain, const u8 *sym_code, const u64 *dcl, Sink &s) {
  const u8 *sp = b + ss;
  if (plain) {
    u32 w = 0, cnt = 0;
    for (u32 j = 0; j < L; ++j) {
      u8 c = sp[j];
      if (xfer && !is_acgt(c)) continue; /* transferred codes are > 1 by construction */
      w = (w << 2) | sym_code[c];
      if (++cnt == 16) { s.put(w, 32); w = 0; cnt = 0; }
    }
    s.put(w, 2 * cnt);
    return;
  }
  for (u32 j = 0; j < L; ++j) {
    u8 c = sp[j];
    if (xfer && !is_acgt(c)) continue;
    u64 e = dcl[sym_code[c]];
    s.put((u32)e, (u32)(e >> 32));
  }
}

/* Title tokens of one record (tasks.cpp:427-506).  `flags` has bit f set when field f's block flag is
 * 1; `first` = first record of its 32-record block; prev(f, v) yields the previous record's numeric value
 * of field f given this record's value v -- it is called for EVERY numeric field of EVERY record, in field
 * order, so that on the GPU it can be a warp shuffle (lane = record of the block; all 32 lanes walk
 * together); tables are reached through `arena` (table directory + slot maps).  FC / ncf / ncskip are the field classes
 * and the list of non-constant fields of C (the GPU passes shared-memory copies). */
template <class Sink, class Prev>
PHY_HD void title_record(const u8 *b, const u8 *lut, u32 ts, u32 te, const SbClass &C, const FieldClass *FC, const u8 *ncf, const u16 *ncskip,
                         const u32 *arena, u32 flags,
                         bool first, Prev prev, Sink &s) {
  TitleCursor cur; cur.init(b, ts, te, lut);
  Tok t;
  for (u32 k = 0; k < C.nnc; ++k) {
    const u32 f = ncf[k];
    const FieldClass &F = FC[f];
    cur.pos += ncskip[k]; /* every record carries record 0's tokens in the constant fields */
    if (!cur.next(t)) break;
    bool flag = (flags >> f) & 1u;
    if (F.kind == K_NUM) {
      i32 v = (i32)t.v;
      i32 pv = prev(f, v);
      if (first) s.put((u32)wsub(v, F.min_v), F.bits_val);
      else if (!flag) {
        u32 x = F.is_delta ? (u32)wsub(wsub(v, pv), F.min_d) : (u32)wsub(v, F.min_v);
        if (F.has_table) { /* x < diff for real records; lanes that only shadow a record may see anything */
          u64 e = ((const u64 *)(arena + F.cl_off))[x < F.diff ? x : 0u];
          s.put((u32)e, (u32)(e >> 32));
        }
        else s.put(x, F.bits_num);
      }
      continue;
    }
    if (!first && flag) continue;
    u32 len = t.end - t.start;
    if (!F.is_len_const) s.put(len - F.min_len, F.bits_len);
    const u16 *sm = (const u16 *)(arena + F.slotmap_off);
    for (u32 j = 0; j < len; ++j) {
      if (j >= F.len0 || ((F.mism[j >> 5] >> (j & 31)) & 1u)) {
        u32 tid = sm[j < 128 ? j : 128];
        u64 e = ((const u64 *)(arena + C.chr_cl_off + (tid - C.tchr0) * 512u))[b[t.start + j]];
        s.put((u32)e, (u32)(e >> 32));
      }
    }
  }
}

/* ---- title / quality / dna header assembly (tasks.cpp:302-390, 576-605, 519-542) ---------------- */
struct ByteWriter {
  u8 *p; u32 n;
  PHY_HD void byte(u8 v) { p[n++] = v; }
  PHY_HD void word(u32 v) { byte((u8)(v >> 24)); byte((u8)(v >> 16)); byte((u8)(v >> 8)); byte((u8)v); }
};

/* Serial part of the layout: writes every scalar of the three stream headers into the staging area and
 * assigns each tree blob its destination (TableDesc::dst = byte offset inside the staging area; the
 * blobs themselves are copied afterwards, possibly in parallel, by copy_tree_blobs).  Returns false when
 * a table failed to build (tree_len == 0). */
/* tlen[t] = blob length of table t (in), tdst[t] = its byte offset inside the staging area (out); the
 * caller moves them from / to the table directory (the GPU keeps them in shared memory meanwhile). */
PHY_HDN bool layout_headers(const u8 *b, SbClass &C, u32 *arena, const u32 *tlen, u32 *tdst) {
  u8 *stage = (u8 *)(arena + C.stage_off);
  for (u32 i = 0; i < C.ntab; ++i) if (tlen[i] == 0) return false;
  ByteWriter w; w.p = stage; w.n = 0;
  w.word(C.nf); /* tasks.cpp:302 */
  for (u32 f = 0; f < C.nf; ++f) {
    const FieldClass &F = C.f[f];
    const u8 *d0 = b + C.ts0 + F.off0;
    w.byte(F.sep);
    w.byte(F.kind == K_CONST ? 1 : 0);
    if (F.kind == K_CONST) { w.word(F.len0); for (u32 j = 0; j < F.len0; ++j) w.byte(d0[j]); continue; }
    w.byte(F.kind == K_NUM ? 1 : 0);
    if (F.kind == K_NUM) {
      w.word((u32)F.min_v); w.word((u32)F.max_v); w.word((u32)F.min_d); w.word((u32)F.max_d);
      if (F.has_table) { tdst[F.tab] = w.n; w.n += tlen[F.tab]; }
      continue;
    }
    w.byte(F.is_len_const);
    w.word(F.len0); w.word(F.max_len); w.word(F.min_len);
    for (u32 j = 0; j < F.len0; ++j) w.byte(d0[j]);
    u32 acc = 0, nb = 0;
    for (u32 j = 0; j < F.len0; ++j) { /* mask bit 1 = every record equals record 0 here */
      acc = (acc << 1) | (((F.mism[j >> 5] >> (j & 31)) & 1u) ? 0u : 1u);
      if (++nb == 8) { w.byte((u8)acc); acc = 0; nb = 0; }
    }
    if (nb) w.byte((u8)(acc << (8 - nb)));
    const u16 *sm = (const u16 *)(arena + F.slotmap_off);
    for (u32 j = 0; j < (u32)CHARPOS; ++j) {
      u32 tid = sm[j];
      if (tid == NOTAB) continue;
      tdst[tid] = w.n; w.n += tlen[tid];
    }
  }
  C.thdr_len = w.n;
  if (w.n > C.thdr_cap) return false;
  /* quality header: alphabet then max_qlen+1 tables (tasks.cpp:576-605) */
  u32 qb = C.thdr_cap;
  w.n = qb;
  for (u32 i = 0; i < C.nq; ++i) w.byte(C.quals[i]);
  for (u32 p = 0; p <= C.max_qlen; ++p) { tdst[C.tq0 + p] = w.n; w.n += tlen[C.tq0 + p]; }
  C.qhdr_len = w.n - qb;
  /* dna header: symbols then the table when not plain (tasks.cpp:519-542) */
  u32 db = C.thdr_cap + C.qhdr_cap;
  w.n = db;
  for (u32 i = 0; i < C.nsym; ++i) w.byte(C.symbols[i]);
  if (!C.plain) { tdst[C.tdna] = w.n; w.n += tlen[C.tdna]; }
  C.dhdr_len = w.n - db;
  return C.qhdr_len <= C.qhdr_cap && C.dhdr_len <= C.dhdr_cap;
}

/* Section sizes once the body sizes are known (phyNGSC.cpp:793-840 concatenates info|title|quality|dna). */
PHY_HD void finish_layout(SbClass &C, u32 title_body_bytes, u64 qbits, u64 dbits) {
  C.qbits_total = qbits; C.dbits_total = dbits;
  C.title_len = C.thdr_len + title_body_bytes;
  C.qual_len = C.qhdr_len + (u32)((qbits + 7) / 8);
  C.dna_len = C.dhdr_len + (u32)((dbits + 7) / 8);
  C.payload_len = C.info_len + C.title_len + C.qual_len + C.dna_len;
}

/* ---- window chaining (phyNGSC.cpp:113-164, 254-331, 744-755) ------------------------------------- */
/* Geometry of one rank's working region and the chaining state that survives from window to window
 * (and from batch to batch).  Positions are relative to the region start (p_wr_start). */
struct PlanState {
  i64 region;       /* p_working_region = file_size / np                                   */
  i64 wr_len;       /* p_wr_end + 1 - p_wr_start                                           */
  i64 rsize;        /* r_buffer_size in force                                              */
  i64 bytes_read;   /* p_bytes_read                                                        */
  i32 overlap;      /* 500 or 0                                                            */
  i32 is_last;      /* last rank                                                           */
  u32 rec_start;    /* rec_start_pos of the next window (non-zero only for the very first) */
  u32 record_cap;
  u32 threads;      /* no_threads of the reference: where in a window its stop rule starts to apply */
  i32 done, status;
  u32 n_subblocks_total;
};

struct SbPlan {
  u64 win_off;      /* region-relative window start                                        */
  u64 win_len;
  u32 rec_start; i32 overlap;
  u32 first_rec, n_records;  /* indices into the batch record table                        */
  u32 warnings; i32 status;
  u64 bytes_consumed;
  u32 chunk_base;   /* first 128-record work item of this subblock                         */
  u32 pad;
};

PHY_HD void plan_init(PlanState &st, u64 file_size, i32 np, i32 rank, u64 window_bytes, u32 overlap, u32 record_cap, u32 first_rec_start, u32 threads = 1) {
  i64 region = (i64)(file_size / (u64)np);
  i64 wr_start = (i64)rank * region;
  i64 wr_end = (rank != np - 1) ? wr_start + region + (i64)overlap - 1 : (i64)file_size - 1;
  st.region = region; st.wr_len = wr_end + 1 - wr_start;
  st.overlap = (i32)overlap; st.is_last = rank == np - 1;
  st.rsize = (i64)window_bytes;
  if (region < st.rsize) { st.rsize = wr_end - wr_start + 1; if (st.is_last) st.overlap = 0; } /* phyNGSC.cpp:119-124 */
  st.bytes_read = 0; st.rec_start = first_rec_start; st.record_cap = record_cap; st.threads = threads ? threads : 1u;
  st.done = region <= 0; st.status = 0; st.n_subblocks_total = 0;
}

}  // namespace phy
