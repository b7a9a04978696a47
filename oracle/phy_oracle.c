/*
 * phy_oracle.c -- TEST INFRASTRUCTURE, not product code (see phy_oracle.h).
 *
 * A from-scratch C restatement of what phyNGSC computes per subblock and per rank.  Every
 * function names the reference lines it follows (paths relative to /root/reference).  The
 * structure is deliberately different from the reference (closed-form per-field reductions,
 * non-mutating ambiguity transfer, two-queue Huffman) so that agreement with the compiled
 * reference is evidence about the *format*, not about shared code.
 *
 * Parity: pinned against oracle/_ref (the unmodified reference built here); see phy_oracle.h.
 */
#include "phy_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * MSB-first bit writer.  Follows the observable behaviour of BitStream / BitMemory
 * (bit_stream.h:80-265, bit_memory.h:149-348): bits concatenate MSB-first, words are big-endian,
 * FlushPartialWordBuffer zero-pads to the next byte.  Whole bytes/words are only ever written at
 * byte-aligned positions by the callers, which is asserted here.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint8_t *p;
  size_t len, cap;
  uint32_t cur; /* pending bits, right-aligned */
  int nb;       /* 0..7 pending bits           */
  int bad;      /* sticky: oom or misuse       */
} bw_t;

static void bw_init(bw_t *w) { memset(w, 0, sizeof(*w)); }
static void bw_free(bw_t *w) { free(w->p); memset(w, 0, sizeof(*w)); }
static void bw_raw_byte(bw_t *w, uint8_t b) {
  if (w->len == w->cap) {
    size_t nc = w->cap ? w->cap * 2 : 4096;
    uint8_t *np_ = (uint8_t *)realloc(w->p, nc);
    if (!np_) { w->bad = 1; return; }
    w->p = np_; w->cap = nc;
  }
  w->p[w->len++] = b;
}
static void bw_bits(bw_t *w, uint32_t v, uint32_t n) { /* n <= 32 */
  while (n) {
    uint32_t take = 8u - (uint32_t)w->nb;
    if (take > n) take = n;
    uint32_t chunk = (n == 32 && take == 32) ? v : ((v >> (n - take)) & ((1u << take) - 1u));
    w->cur = (w->cur << take) | chunk;
    w->nb += (int)take;
    n -= take;
    if (w->nb == 8) { bw_raw_byte(w, (uint8_t)w->cur); w->cur = 0; w->nb = 0; }
  }
}
static void bw_align(bw_t *w) {
  if (w->nb) { bw_raw_byte(w, (uint8_t)(w->cur << (8 - w->nb))); w->cur = 0; w->nb = 0; }
}
static void bw_byte(bw_t *w, uint8_t b) { if (w->nb) w->bad = 1; bw_raw_byte(w, b); }
static void bw_word(bw_t *w, uint32_t v) {
  bw_byte(w, (uint8_t)(v >> 24)); bw_byte(w, (uint8_t)(v >> 16)); bw_byte(w, (uint8_t)(v >> 8)); bw_byte(w, (uint8_t)v);
}
static void bw_bytes(bw_t *w, const uint8_t *s, size_t n) { for (size_t i = 0; i < n; ++i) bw_byte(w, s[i]); }

/* BitStream::BitLength, bit_stream.h:268-277.  Callers pass int32 differences converted to
 * uint64, so a negative difference yields 64. */
static uint32_t bit_length_u64(uint64_t x) {
  for (uint32_t i = 0; i < 32; ++i) if (x < (1ull << i)) return i;
  return 64;
}
static uint32_t bit_length_i32(int32_t d) { return bit_length_u64((uint64_t)(int64_t)d); }

/* ------------------------------------------------------------------------------------------
 * Huffman.  huffman.cpp:18-85 (Complete), :88-118 + huffman.h:134-147 (member StoreTree /
 * EncodeProcess), :191-205 (static StoreTree).  The reference pops a binary heap ordered by the
 * strict total order (frequency, id); because ids are unique the pop sequence equals a sorted
 * list merged with a FIFO of internal nodes (leaf ids < internal ids, older internals first).
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t freq, id; } hitem_t;
static int hitem_cmp(const void *a, const void *b) {
  const hitem_t *x = (const hitem_t *)a, *y = (const hitem_t *)b;
  if (x->freq != y->freq) return x->freq < y->freq ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}

typedef struct {
  uint32_t n, p, root;
  int32_t left[1024], right[1024]; /* n <= 512 -> 2n-1 nodes */
  uint32_t code[1024], len[1024];
} htree_t;

static int huff_build(const uint32_t *freq, uint32_t n, int compact, htree_t *t) {
  if (n == 0 || n > 512) return -1;
  hitem_t leaves[512], inner[512];
  for (uint32_t i = 0; i < n; ++i) { leaves[i].freq = freq[i]; leaves[i].id = i; }
  qsort(leaves, n, sizeof(hitem_t), hitem_cmp);
  uint32_t lo = 0;
  if (compact) while (n - lo > 2 && leaves[lo].freq == 0) ++lo; /* huffman.cpp:44-50 */
  uint32_t p = n - lo;
  t->n = n; t->p = p;
  for (uint32_t i = 0; i < 2 * n - 1; ++i) { t->left[i] = t->right[i] = -1; t->code[i] = 0; t->len[i] = 0; }
  uint32_t li = lo, qi = 0, qn = 0;
  for (uint32_t i = 0; i + 1 < p; ++i) { /* huffman.cpp:57-70 */
    hitem_t pick[2];
    for (int k = 0; k < 2; ++k) {
      int take_leaf;
      if (li < n && qi < qn) take_leaf = hitem_cmp(&leaves[li], &inner[qi]) < 0;
      else take_leaf = li < n;
      pick[k] = take_leaf ? leaves[li++] : inner[qi++];
    }
    inner[qn].freq = pick[0].freq + pick[1].freq;
    inner[qn].id = n + i;
    ++qn;
    t->left[n + i] = (int32_t)pick[0].id;
    t->right[n + i] = (int32_t)pick[1].id;
  }
  t->root = n + p - 2; /* huffman.cpp:82; n == 1 -> 0 (the leaf itself) */
  if (p >= 2)
    for (uint32_t i = n + p - 2; i >= n; --i) { /* huffman.cpp:73-79 */
      uint32_t l = (uint32_t)t->left[i], r = (uint32_t)t->right[i];
      if (t->len[i] + 1 > 32) return -2; /* the reference's 32-bit code would overflow */
      t->len[l] = t->len[r] = t->len[i] + 1;
      t->code[l] = t->code[i] << 1;
      t->code[r] = (t->code[i] << 1) | 1u;
      if (i == 0) break;
    }
  return 0;
}

static void huff_walk(const htree_t *t, uint32_t node, uint32_t bits_per_id, bw_t *w) {
  /* explicit stack pre-order: huffman.h:134-147 */
  uint32_t stack[1024]; int sp = 0;
  stack[sp++] = node;
  while (sp) {
    uint32_t v = stack[--sp];
    if (t->left[v] < 0) { bw_bits(w, 1, 1); bw_bits(w, v, bits_per_id); }
    else { bw_bits(w, 0, 1); stack[sp++] = (uint32_t)t->right[v]; stack[sp++] = (uint32_t)t->left[v]; }
  }
}

/* Appends [align][word mem_size][mem] to w.  huffman.cpp:88-118, 191-205. */
static void huff_store(const htree_t *t, bw_t *w) {
  uint32_t n = t->n, bits_per_id = 0;
  for (uint32_t tmp = 2; tmp <= n; tmp *= 2) ++bits_per_id; /* utils::int_log(n, 2) */
  if (n & (n - 1)) ++bits_per_id;
  uint32_t min_len = n;
  for (uint32_t i = 0; i < n; ++i) if (t->len[i] < min_len && t->len[i] > 0) min_len = t->len[i];
  bw_t m; bw_init(&m);
  bw_word(&m, t->root);
  bw_word(&m, n);
  bw_byte(&m, (uint8_t)min_len);
  huff_walk(t, t->root, bits_per_id, &m);
  bw_align(&m);
  bw_align(w);
  bw_word(w, (uint32_t)m.len);
  bw_bytes(w, m.p, m.len);
  if (m.bad) w->bad = 1;
  bw_free(&m);
}

uint32_t phy_oracle_huffman(const uint32_t *freq, uint32_t n, int compact, uint32_t *code, uint32_t *len,
                            uint8_t *tree_out, uint32_t tree_cap) {
  htree_t *t = (htree_t *)malloc(sizeof(htree_t));
  if (!t) return 0;
  if (huff_build(freq, n, compact, t)) { free(t); return 0; }
  for (uint32_t i = 0; i < n; ++i) { code[i] = t->code[i]; len[i] = t->len[i]; }
  bw_t w; bw_init(&w);
  huff_store(t, &w);
  uint32_t out = 0;
  if (!w.bad && w.len <= tree_cap) { memcpy(tree_out, w.p, w.len); out = (uint32_t)w.len; }
  bw_free(&w); free(t);
  return out;
}

/* ------------------------------------------------------------------------------------------
 * Record splitting, one thread.  phyNGSC.cpp:254-331 with no_threads == 1.
 *   - the first record is located by the "\n+\n" pattern (:276-296) and taken unconditionally;
 *   - every later record is taken iff its title newline lies before the window end (:268, :303);
 *   - the stop rule (:315) and the record cap (:321) are only evaluated after such a later record.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t title_end, seq_end; } rec_t;

static int64_t find_nl(const uint8_t *b, int64_t from, int64_t limit) {
  const uint8_t *q = (from < limit) ? (const uint8_t *)memchr(b + from, '\n', (size_t)(limit - from)) : 0;
  return q ? (int64_t)(q - b) : -1;
}

/* no_threads of the reference (its third CLI argument).  Thread t owns the window bytes [t*size/T, (t+1)*size/T)
 * (phyNGSC.cpp:261-266) and takes every record whose title newline lies there (:268, :303); only the LAST thread
 * evaluates the stop rule (:315), and like every thread not before its own second record.  So with T > 1 every record
 * whose title ends before (T-1)*size/T is taken unconditionally -- the same records as with one thread unless the window
 * is shorter than about T x overlap (the small final window of a rank).  Per-thread record caps are not restated. */
static int g_threads = 1;
void phy_oracle_set_threads(int threads) { g_threads = threads < 1 ? 1 : threads; }

static int split_records(const uint8_t *b, int64_t readable, int64_t size, int64_t rec_start, int64_t overlap,
                         uint32_t cap, rec_t **out, uint32_t *n_out, uint32_t *warn) {
  const int64_t bT = (int64_t)(g_threads - 1) * size / g_threads; /* first byte of the last thread's slice */
  uint32_t n = 0, room = 1024;
  rec_t *r = (rec_t *)malloc(room * sizeof(rec_t));
  if (!r) return PHY_ORACLE_ENOMEM;
  int64_t t = find_nl(b, rec_start, readable);
  int64_t s = t < 0 ? -1 : find_nl(b, t + 1, readable);
  if (s < 0 || s + 2 >= readable || b[s + 1] != '+' || b[s + 2] != '\n' || s == t + 1) { free(r); return PHY_ORACLE_EMALFORMED; }
  r[n].title_end = (uint32_t)t; r[n].seq_end = (uint32_t)s; ++n;
  int64_t i = s + (s - t) + 3;
  for (;;) {
    int64_t lim = size < readable ? size : readable;
    t = find_nl(b, i, lim);
    if (t < 0) break;
    s = find_nl(b, t + 1, readable);
    if (s < 0) { free(r); return PHY_ORACLE_EMALFORMED; }
    if (n == room) {
      room *= 2;
      rec_t *nr = (rec_t *)realloc(r, room * sizeof(rec_t));
      if (!nr) { free(r); return PHY_ORACLE_ENOMEM; }
      r = nr;
    }
    const int64_t prev_t = (int64_t)r[n - 1].title_end;
    r[n].title_end = (uint32_t)t; r[n].seq_end = (uint32_t)s; ++n;
    i = s + (s - t) + 3;
    const int in_last_slice = t >= bT, first_of_last_slice = g_threads > 1 && prev_t < bT;
    if (in_last_slice && !first_of_last_slice && i >= size - overlap) break;
    if (n > cap) { *warn |= 1u; break; }
    ++i;
  }
  *out = r; *n_out = n;
  return PHY_ORACLE_OK;
}

/* ------------------------------------------------------------------------------------------
 * Titles.  Tokeniser phyNGSC.cpp:208, 342-423; statistics tasks.cpp:22-223 restated as
 * reductions over the token table; stream tasks.cpp:289-510.
 * ---------------------------------------------------------------------------------------- */
static int is_sep(uint8_t c) {
  switch (c) { case ' ': case '.': case '_': case ',': case '=': case ':': case '/': case '-': case '#': case '\n': return 1; }
  return 0;
}
static int tok_is_num(const uint8_t *s, uint32_t len) { /* utils.h:107-114 */
  for (uint32_t i = 0; i < len; ++i) if (s[i] < '0' || s[i] > '9') return 0;
  return len == 1 || (len > 1 && s[0] != '0');
}
static int32_t tok_to_num(const uint8_t *s, uint32_t len) { /* utils.h:117-125, wraps mod 2^32 */
  uint32_t r = 0;
  for (uint32_t i = 0; i < len; ++i) r = r * 10u + (uint32_t)(s[i] - '0');
  return (int32_t)r;
}
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

typedef struct {
  uint32_t len0, min_len, max_len;
  uint8_t sep;
  int is_const, is_len_const, is_num, is_delta;
  int32_t min_v, max_v, min_d, max_d;
  uint32_t bits_num, bits_val, bits_len;
  int has_table;          /* numeric Huffman table present                      */
  htree_t *num_tree;      /* over diff symbols                                  */
  htree_t **chr_tree;     /* [129], NULL where no table                         */
  uint8_t *mask;          /* [len0]                                             */
} fld_t;

typedef struct { uint32_t start, end; } tok_t; /* byte range [start,end) in the window */

static void fld_free(fld_t *f, uint32_t nf) {
  if (!f) return;
  for (uint32_t i = 0; i < nf; ++i) {
    free(f[i].num_tree); free(f[i].mask);
    if (f[i].chr_tree) { for (int j = 0; j < 129; ++j) free(f[i].chr_tree[j]); free(f[i].chr_tree); }
  }
  free(f);
}

static int encode_title(const uint8_t *b, const rec_t *rec, uint32_t R, uint32_t rec_start, bw_t *w) {
  int rc = PHY_ORACLE_OK;
  /* n_fields from record 0 (phyNGSC.cpp:345-379) */
  uint32_t nf = 0;
  for (uint32_t i = rec_start; i <= rec[0].title_end; ++i) nf += (uint32_t)is_sep(b[i]);
  tok_t *tok = (tok_t *)malloc((size_t)R * nf * sizeof(tok_t));
  fld_t *fl = (fld_t *)calloc(nf, sizeof(fld_t));
  uint32_t *hist = (uint32_t *)malloc(129u * 256u * sizeof(uint32_t));
  if (!tok || !fl || !hist) { rc = PHY_ORACLE_ENOMEM; goto done; }

  /* token table (phyNGSC.cpp:383-423): record r's title line starts after the previous newline */
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t st = rec_start;
    if (r) { st = rec[r].title_end; while (b[st - 1] != '\n') --st; }
    uint32_t k = 0, fs = st;
    for (uint32_t i = st; i <= rec[r].title_end; ++i) {
      if (!is_sep(b[i])) continue;
      if (k < nf) { tok[(size_t)r * nf + k].start = fs; tok[(size_t)r * nf + k].end = i; }
      ++k; fs = i + 1;
    }
    if (k != nf) { rc = PHY_ORACLE_EFIELDS; goto done; }
  }

  /* P of quirk Q3: libstdc++ vector growth wipes the seeded value histogram of every field whose
   * index is below the last reallocation point (structures.h:126-202, phyNGSC.cpp:350,368). */
  uint32_t P = 0;
  if (nf >= 2) { P = 1; while (P * 2 <= nf - 1) P *= 2; }

  bw_word(w, nf); /* tasks.cpp:302 */
  for (uint32_t f = 0; f < nf; ++f) {
    fld_t *F = &fl[f];
    const tok_t *t0 = &tok[f];
    const uint8_t *d0 = b + t0->start;
    F->len0 = t0->end - t0->start;
    F->sep = b[t0->end];
    F->min_len = F->max_len = F->len0;
    F->is_const = F->is_len_const = 1;
    F->is_num = 1;
    F->mask = (uint8_t *)malloc(F->len0 ? F->len0 : 1);
    if (!F->mask) { rc = PHY_ORACLE_ENOMEM; goto done; }
    memset(F->mask, 1, F->len0);
    int32_t prev = 0;
    F->min_d = 1; F->max_d = -1; /* Field::Field defaults, structures.h:103-106 */
    for (uint32_t r = 0; r < R; ++r) {
      const tok_t *t = &tok[(size_t)r * nf + f];
      const uint8_t *d = b + t->start;
      uint32_t len = t->end - t->start;
      if (len > F->max_len) F->max_len = len;
      if (len < F->min_len) F->min_len = len;
      if (len != F->len0) { F->is_const = 0; F->is_len_const = 0; }
      else if (memcmp(d, d0, len)) F->is_const = 0;
      uint32_t m = len < F->len0 ? len : F->len0;
      for (uint32_t p = 0; p < m; ++p) if (d[p] != d0[p]) F->mask[p] = 0;
      if (F->is_num) {
        if (!tok_is_num(d, len)) F->is_num = 0;
        else {
          int32_t v = tok_to_num(d, len);
          if (r == 0) { F->min_v = F->max_v = v; }
          if (v < F->min_v) F->min_v = v;
          if (v > F->max_v) F->max_v = v;
          if (r >= 1) {
            int32_t dl = wsub(v, prev);
            if (r == 1) { F->min_d = F->max_d = dl; }
            if (dl > F->max_d) F->max_d = dl;
            if (dl < F->min_d) F->min_d = dl;
          }
          prev = v;
        }
      }
    }
    /* header, tasks.cpp:304-390 */
    bw_byte(w, F->sep);
    bw_byte(w, (uint8_t)F->is_const);
    if (F->is_const) { bw_word(w, F->len0); bw_bytes(w, d0, F->len0); continue; }
    bw_byte(w, (uint8_t)F->is_num);
    if (F->is_num) {
      int32_t vr = wsub(F->max_v, F->min_v), dr = wsub(F->max_d, F->min_d);
      F->is_delta = !(vr < dr); /* tasks.cpp:208-217 */
      F->bits_num = bit_length_i32(F->is_delta ? dr : vr);
      F->bits_val = bit_length_i32(vr);
      bw_word(w, (uint32_t)F->min_v); bw_word(w, (uint32_t)F->max_v);
      bw_word(w, (uint32_t)F->min_d); bw_word(w, (uint32_t)F->max_d);
      int32_t diff = (F->is_delta ? dr : vr) + 1, base = F->is_delta ? F->min_d : F->min_v;
      int stats_nonempty = F->is_delta ? (R >= 2) : (f >= P);
      if (diff <= 512 && stats_nonempty) { /* tasks.cpp:338; a negative diff (R == 1) also lands here in the
                                              reference but then the map is empty */
        if (diff <= 0) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
        uint32_t fr[512];
        memset(fr, 0, sizeof(fr));
        prev = 0;
        for (uint32_t r = 0; r < R; ++r) {
          const tok_t *t = &tok[(size_t)r * nf + f];
          int32_t v = tok_to_num(b + t->start, t->end - t->start);
          if (F->is_delta) { if (r >= 1) fr[(uint32_t)wsub(wsub(v, prev), base)]++; }
          else fr[(uint32_t)wsub(v, base)] += (r == 0) ? 2u : 1u; /* seed at phyNGSC.cpp:368 + tasks.cpp:177 */
          prev = v;
        }
        F->num_tree = (htree_t *)malloc(sizeof(htree_t));
        if (!F->num_tree) { rc = PHY_ORACLE_ENOMEM; goto done; }
        if (huff_build(fr, (uint32_t)diff, 1, F->num_tree)) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
        huff_store(F->num_tree, w);
        F->has_table = 1;
      }
      continue;
    }
    /* string field */
    if (F->max_len == 128) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; } /* Q11: reference reads chars[128] out of range */
    F->bits_len = bit_length_u64((uint64_t)(F->max_len - F->min_len));
    bw_byte(w, (uint8_t)F->is_len_const);
    bw_word(w, F->len0); bw_word(w, F->max_len); bw_word(w, F->min_len);
    bw_bytes(w, d0, F->len0);
    for (uint32_t j = 0; j < F->len0; ++j) bw_bits(w, F->mask[j], 1);
    memset(hist, 0, 129u * 256u * sizeof(uint32_t)); /* tasks.cpp:83-93 */
    for (uint32_t r = 0; r < R; ++r) {
      const tok_t *t = &tok[(size_t)r * nf + f];
      const uint8_t *d = b + t->start;
      uint32_t len = t->end - t->start;
      for (uint32_t x = 0; x < len; ++x) hist[(x < 128 ? x : 128) * 256u + d[x]]++;
    }
    F->chr_tree = (htree_t **)calloc(129, sizeof(htree_t *));
    if (!F->chr_tree) { rc = PHY_ORACLE_ENOMEM; goto done; }
    uint32_t ntab = F->max_len < 128 ? F->max_len : 128;
    for (uint32_t j = 0; j <= 128; ++j) {
      int need = (j < ntab) ? (j >= F->len0 || !F->mask[j]) : (j == 128 && F->max_len >= 128);
      if (!need) continue;
      F->chr_tree[j] = (htree_t *)malloc(sizeof(htree_t));
      if (!F->chr_tree[j]) { rc = PHY_ORACLE_ENOMEM; goto done; }
      if (huff_build(hist + j * 256u, 256, 1, F->chr_tree[j])) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
      huff_store(F->chr_tree[j], w);
    }
    bw_align(w);
  }

  /* body, tasks.cpp:393-509 */
  for (uint32_t lo = 0; lo < R; lo += 32) {
    uint32_t hi = lo + 32 < R ? lo + 32 : R;
    uint8_t flag[64 * 16];
    uint8_t *fg = nf <= sizeof(flag) ? flag : (uint8_t *)malloc(nf);
    if (!fg) { rc = PHY_ORACLE_ENOMEM; goto done; }
    for (uint32_t f = 0; f < nf; ++f) {
      fld_t *F = &fl[f];
      if (F->is_const) continue;
      const tok_t *tl = &tok[(size_t)lo * nf + f];
      int bit = 1;
      if (!F->is_num) { /* is_block_constant, tasks.cpp:64-81 */
        for (uint32_t r = lo + 1; r < hi && bit; ++r) {
          const tok_t *t = &tok[(size_t)r * nf + f];
          bit = (t->end - t->start == tl->end - tl->start) && !memcmp(b + t->start, b + tl->start, t->end - t->start);
        }
      } else {
        int32_t v0 = tok_to_num(b + tl->start, tl->end - tl->start), pv = v0, bd = 0;
        for (uint32_t r = lo + 1; r < hi; ++r) {
          const tok_t *t = &tok[(size_t)r * nf + f];
          int32_t v = tok_to_num(b + t->start, t->end - t->start);
          if (F->is_delta) { /* tasks.cpp:127-147 and :415 */
            if (r == lo + 1) bd = wsub(v, pv);
            else if (wsub(v, pv) != bd) bit = 0;
          } else if (v != v0) bit = 0;
          pv = v;
        }
        if (F->is_delta && bd != F->min_d) bit = 0;
      }
      fg[f] = (uint8_t)bit;
      bw_bits(w, (uint32_t)bit, 1);
    }
    for (uint32_t r = lo; r < hi; ++r) {
      for (uint32_t f = 0; f < nf; ++f) {
        fld_t *F = &fl[f];
        if (F->is_const) continue;
        const tok_t *t = &tok[(size_t)r * nf + f];
        const uint8_t *d = b + t->start;
        uint32_t len = t->end - t->start;
        if (F->is_num) {
          int32_t v = tok_to_num(d, len);
          if (F->bits_val > 32) { rc = PHY_ORACLE_EUNSUPPORTED; if (fg != flag) free(fg); goto done; }
          if (r == lo) bw_bits(w, (uint32_t)wsub(v, F->min_v), F->bits_val);
          else if (!fg[f]) {
            const tok_t *tp = &tok[(size_t)(r - 1) * nf + f];
            int32_t pv = tok_to_num(b + tp->start, tp->end - tp->start);
            uint32_t ts = (uint32_t)(F->is_delta ? wsub(wsub(v, pv), F->min_d) : wsub(v, F->min_v));
            if (F->has_table) bw_bits(w, F->num_tree->code[ts], F->num_tree->len[ts]);
            else {
              if (F->bits_num > 32) { rc = PHY_ORACLE_EUNSUPPORTED; if (fg != flag) free(fg); goto done; }
              bw_bits(w, ts, F->bits_num);
            }
          }
          continue;
        }
        if (r != lo && fg[f]) continue;
        if (!F->is_len_const) bw_bits(w, len - F->min_len, F->bits_len);
        for (uint32_t j = 0; j < len; ++j)
          if (j >= F->len0 || !F->mask[j]) {
            const htree_t *ht = F->chr_tree[j < 128 ? j : 128];
            bw_bits(w, ht->code[d[j]], ht->len[d[j]]);
          }
      }
    }
    bw_align(w);
    if (fg != flag) free(fg);
  }
done:
  free(tok); free(hist); fld_free(fl, nf);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * DNA / quality.  Ambiguity transfer phyNGSC.cpp:184-206, 549-588 (restated without mutating the
 * window); statistics :593-619, 669-686 and tasks.cpp:226-286; streams tasks.cpp:513-622.
 * ---------------------------------------------------------------------------------------- */
static uint8_t amb_code(uint8_t c) {
  switch (c) {
    case 'A': case 'C': case 'G': case 'T': return 1;
    case 'Y': return 2; case 'R': return 3; case 'W': return 4; case 'S': return 5; case 'K': return 6;
    case 'M': return 7; case 'D': return 8; case 'V': return 9; case 'H': return 10; case 'B': return 11;
    case 'N': return 12; case 'X': return 13; case 'U': return 14; case '.': return 15; case '-': return 16;
  }
  return 0;
}

int phy_oracle_compress_window(const uint8_t *win, uint64_t readable, int64_t r_buffer_size, uint32_t rec_start,
                               int32_t overlap, uint32_t record_cap, phy_oracle_subblock *out) {
  memset(out, 0, sizeof(*out));
  rec_t *rec = 0; uint32_t R = 0;
  int rc = split_records(win, (int64_t)readable, r_buffer_size, rec_start, overlap, record_cap, &rec, &R, &out->warnings);
  if (rc) return rc;
  bw_t info, title, qual, dna;
  bw_init(&info); bw_init(&title); bw_init(&qual); bw_init(&dna);
  uint8_t *xfer = (uint8_t *)calloc(R, 1);
  uint32_t *qstat = 0;
  htree_t *qt = 0, *dt = 0;
  if (!xfer) { rc = PHY_ORACLE_ENOMEM; goto done; }

  /* validation the reference does not perform but whose violation is undefined behaviour there */
  for (uint32_t r = 0; r < R; ++r) {
    uint64_t qe = 2ull * rec[r].seq_end - rec[r].title_end + 2;
    if (qe >= readable + 1) { rc = PHY_ORACLE_EMALFORMED; goto done; }
    if (win[rec[r].seq_end + 1] != '+' || win[rec[r].seq_end + 2] != '\n') { rc = PHY_ORACLE_EMALFORMED; goto done; }
    if (qe < readable && win[qe] != '\n') { rc = PHY_ORACLE_EMALFORMED; goto done; }
  }
  { /* colour space (phyNGSC.cpp:473-487): not exercised by any workload, reported as unsupported */
    uint32_t s0 = rec[0].title_end + 1;
    if ((win[s0] >= '0' && win[s0] <= '3') || (win[s0 + 1] >= '0' && win[s0 + 1] <= '3')) { rc = PHY_ORACLE_ECOLORSPACE; goto done; }
  }

  uint32_t dna_occ[256] = {0};
  uint8_t qpresent[256] = {0};
  uint32_t max_qlen = 0, max_slen = 0;
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t ss = rec[r].title_end + 1, se = rec[r].seq_end, L = se - ss, qs = se + 3;
    int any = 0, ok = 1;
    for (uint32_t j = 0; j < L; ++j) {
      uint8_t c = amb_code(win[ss + j]);
      if (c == 1) continue;
      if (c == 0 || win[qs + j] < 33 || win[qs + j] > 40) { ok = 0; break; }
      any = 1;
    }
    xfer[r] = (uint8_t)(any && ok);
    uint32_t kept = 0;
    for (uint32_t j = 0; j < L; ++j) {
      uint8_t c = win[ss + j], q = win[qs + j];
      if (c == 0 || q == 0) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
      uint8_t a = amb_code(c);
      if (xfer[r] && a > 1) q = (uint8_t)(128 + (a << 3) - 16 + (q - 33));
      else { dna_occ[c]++; ++kept; }
      qpresent[q] = 1;
    }
    if (L > max_qlen) max_qlen = L;
    if (kept > max_slen) max_slen = kept;
  }
  uint8_t symbols[256], quals[256], sym_code[256] = {0}, qua_code[256] = {0};
  uint32_t nsym = 0, nq = 0;
  for (uint32_t c = 0; c < 256; ++c) { /* phyNGSC.cpp:669-686 */
    if (dna_occ[c]) { sym_code[c] = (uint8_t)nsym; symbols[nsym++] = (uint8_t)c; }
    if (qpresent[c]) { qua_code[c] = (uint8_t)nq; quals[nq++] = (uint8_t)c; }
  }
  int plain = nsym <= 4; /* tasks.cpp:239-256; the frequency test at :248 can never fire */
  uint32_t flags = 0x8u | 0x2u | 0x4u | 0x20u | 0x80u; /* PLUS_ONLY|DNA_PLAIN|CONST_NUM_FIELDS|DELTA_CONSTANT|VARIABLE_LENGTH (Q1) */
  if (!plain) flags &= ~0x2u;
  if (nsym == 0 || nq == 0) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }

  /* info, phyNGSC.cpp:719-742 */
  bw_word(&info, R); bw_word(&info, max_qlen); bw_word(&info, max_slen);
  bw_byte(&info, (uint8_t)nsym); bw_byte(&info, 0); bw_byte(&info, (uint8_t)nq);
  bw_word(&info, flags);
  {
    uint32_t nb = bit_length_u64(max_qlen);
    for (uint32_t r = 0; r < R; ++r) bw_bits(&info, rec[r].seq_end - rec[r].title_end - 1, nb);
    bw_align(&info);
  }

  /* title */
  rc = encode_title(win, rec, R, rec_start, &title);
  if (rc) goto done;

  /* quality, tasks.cpp:260-286, 572-622 */
  qstat = (uint32_t *)calloc((size_t)(max_qlen + 1) * nq, sizeof(uint32_t));
  qt = (htree_t *)malloc((size_t)(max_qlen + 1) * sizeof(htree_t));
  if (!qstat || !qt) { rc = PHY_ORACLE_ENOMEM; goto done; }
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t ss = rec[r].title_end + 1, se = rec[r].seq_end, L = se - ss, qs = se + 3;
    for (uint32_t j = 0; j < L; ++j) {
      uint8_t q = win[qs + j], a = amb_code(win[ss + j]);
      if (xfer[r] && a > 1) q = (uint8_t)(128 + (a << 3) - 16 + (q - 33));
      qstat[(size_t)(j + 1) * nq + qua_code[q]]++;
      qstat[qua_code[q]]++;
    }
  }
  bw_bytes(&qual, quals, nq);
  for (uint32_t p = 0; p <= max_qlen; ++p) {
    if (huff_build(qstat + (size_t)p * nq, nq, 1, &qt[p])) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
    huff_store(&qt[p], &qual);
  }
  bw_align(&qual);
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t ss = rec[r].title_end + 1, se = rec[r].seq_end, L = se - ss, qs = se + 3;
    for (uint32_t j = 0; j < L; ++j) {
      uint8_t q = win[qs + j], a = amb_code(win[ss + j]);
      if (xfer[r] && a > 1) q = (uint8_t)(128 + (a << 3) - 16 + (q - 33));
      const htree_t *t = &qt[j + 1];
      bw_bits(&qual, t->code[qua_code[q]], t->len[qua_code[q]]);
    }
  }
  bw_align(&qual);

  /* dna, tasks.cpp:513-569 */
  bw_bytes(&dna, symbols, nsym);
  if (!plain) {
    uint32_t st[256];
    for (uint32_t i = 0; i < nsym; ++i) st[i] = dna_occ[symbols[i]];
    dt = (htree_t *)malloc(sizeof(htree_t));
    if (!dt) { rc = PHY_ORACLE_ENOMEM; goto done; }
    if (huff_build(st, nsym, 1, dt)) { rc = PHY_ORACLE_EUNSUPPORTED; goto done; }
    huff_store(dt, &dna);
  }
  for (uint32_t r = 0; r < R; ++r) {
    uint32_t ss = rec[r].title_end + 1, se = rec[r].seq_end;
    for (uint32_t j = ss; j < se; ++j) {
      uint8_t c = win[j];
      if (xfer[r] && amb_code(c) > 1) continue;
      if (plain) bw_bits(&dna, sym_code[c], 2);
      else bw_bits(&dna, dt->code[sym_code[c]], dt->len[sym_code[c]]);
    }
  }
  bw_align(&dna);

  if (info.bad || title.bad || qual.bad || dna.bad) { rc = PHY_ORACLE_ENOMEM; goto done; }
  out->n_records = R;
  out->bytes_consumed = 2ull * rec[R - 1].seq_end - rec[R - 1].title_end + 3; /* phyNGSC.cpp:745 */
  out->len[0] = (uint32_t)info.len; out->len[1] = (uint32_t)title.len;
  out->len[2] = (uint32_t)qual.len; out->len[3] = (uint32_t)dna.len;
  out->payload_len = out->len[0] + out->len[1] + out->len[2] + out->len[3];
  out->payload = (uint8_t *)malloc(out->payload_len ? out->payload_len : 1);
  if (!out->payload) { rc = PHY_ORACLE_ENOMEM; goto done; }
  { /* phyNGSC.cpp:809-838: info, title, quality, dna */
    uint8_t *p = out->payload;
    memcpy(p, info.p, info.len); p += info.len;
    memcpy(p, title.p, title.len); p += title.len;
    memcpy(p, qual.p, qual.len); p += qual.len;
    memcpy(p, dna.p, dna.len);
  }
done:
  bw_free(&info); bw_free(&title); bw_free(&qual); bw_free(&dna);
  free(rec); free(xfer); free(qstat); free(qt); free(dt);
  return rc;
}

void phy_oracle_subblock_free(phy_oracle_subblock *sb) { free(sb->payload); memset(sb, 0, sizeof(*sb)); }

/* ------------------------------------------------------------------------------------------
 * Container: block header tasks.cpp:1179-1200 + structures.h:323-333; footer tasks.cpp:1104-1176.
 * ---------------------------------------------------------------------------------------- */
static int ceil_log2_u(uint64_t x) { int b = 0; while ((1ull << b) < x) ++b; return b; }       /* ceil(log2 x), x >= 1 */
static int bitlen_u(uint64_t x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }            /* floor(log2 x) + 1     */

uint32_t phy_oracle_make_header(int wrid, int bewr, int bhs, int beso, int bcss, const uint32_t *sbol, uint32_t nosb,
                                uint8_t *out, uint32_t cap) {
  bw_t w; bw_init(&w);
  bw_bits(&w, (uint32_t)wrid, (uint32_t)bewr);
  bw_bits(&w, (uint32_t)bhs, 12);
  bw_bits(&w, nosb, 6);
  bw_bits(&w, (uint32_t)beso, 5);
  bw_bits(&w, (uint32_t)bcss, 2);
  for (uint32_t i = 0; i < nosb; ++i) bw_bits(&w, sbol[i], (uint32_t)beso);
  bw_align(&w);
  uint32_t n = (w.bad || w.len > cap) ? 0 : (uint32_t)w.len;
  if (n) memcpy(out, w.p, n);
  bw_free(&w);
  return n;
}

int32_t phy_oracle_make_footer(int np, uint64_t fastq_size, uint32_t n_blocks, uint32_t n_subblocks,
                               const int32_t *overlaps, const int32_t *block_order, const uint32_t *lb_sizes,
                               uint8_t *out, uint32_t cap) {
  uint32_t lb_max = 0, lb_min = 0xFFFFFFFFu; int32_t ov_max = 0;
  for (int i = 0; i < np; ++i) {
    if (lb_sizes[i] > lb_max) lb_max = lb_sizes[i];
    if (lb_sizes[i] < lb_min) lb_min = lb_sizes[i];
    if (overlaps[i] > ov_max) ov_max = overlaps[i];
  }
  if (ov_max <= 0 || lb_max == 0 || n_blocks == 0 || n_subblocks == 0) return -1; /* Q12: log2(0) in the reference */
  int BEPS = bitlen_u((uint64_t)np), BEFS = bitlen_u(fastq_size), BEBS = bitlen_u(n_blocks), BESS = bitlen_u(n_subblocks);
  int BELB = bitlen_u(lb_max), BEOV = bitlen_u((uint64_t)ov_max), LBES = lb_max == lb_min;
  bw_t w; bw_init(&w);
  bw_bits(&w, (uint32_t)BEPS, 4); bw_bits(&w, (uint32_t)BEFS, 6); bw_bits(&w, (uint32_t)BEBS, 4);
  bw_bits(&w, (uint32_t)BESS, 4); bw_bits(&w, (uint32_t)BELB, 5); bw_bits(&w, (uint32_t)BEOV, 4);
  bw_bits(&w, (uint32_t)LBES, 1);
  bw_bits(&w, (uint32_t)np, (uint32_t)BEPS);
  if (BEFS > 32) { bw_bits(&w, (uint32_t)(fastq_size >> 32), (uint32_t)(BEFS - 32)); bw_bits(&w, (uint32_t)fastq_size, 32); }
  else bw_bits(&w, (uint32_t)fastq_size, (uint32_t)BEFS);
  bw_bits(&w, n_blocks, (uint32_t)BEBS);
  bw_bits(&w, n_subblocks, (uint32_t)BESS);
  for (int i = 1; i < np; ++i) bw_bits(&w, (uint32_t)overlaps[i], (uint32_t)BEOV);
  int cbo = ceil_log2_u((uint64_t)np);
  for (uint32_t i = 0; i < n_blocks; ++i) bw_bits(&w, (uint32_t)block_order[i], (uint32_t)cbo);
  if (!LBES) for (int i = 0; i < np; ++i) bw_bits(&w, lb_sizes[i], (uint32_t)BELB);
  bw_align(&w);
  uint32_t flen = (uint32_t)w.len;
  bw_byte(&w, (uint8_t)(flen >> 8)); bw_byte(&w, (uint8_t)flen);
  int32_t n = (w.bad || w.len > cap) ? -2 : (int32_t)w.len;
  if (n > 0) memcpy(out, w.p, (size_t)n);
  bw_free(&w);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * One rank: partition phyNGSC.cpp:113-164, window chaining :168-250, 744-755, block assembly
 * :842-928.
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint8_t *p; uint64_t len, cap; } buf_t;
static int buf_put(buf_t *b, const uint8_t *s, uint64_t n) {
  if (b->len + n > b->cap) {
    uint64_t nc = b->cap ? b->cap : 1 << 20;
    while (nc < b->len + n) nc *= 2;
    uint8_t *q = (uint8_t *)realloc(b->p, nc);
    if (!q) return -1;
    b->p = q; b->cap = nc;
  }
  memcpy(b->p + b->len, s, n); b->len += n;
  return 0;
}
#define GROW(ptr, n, type) do { type *q_ = (type *)realloc(ptr, ((size_t)(n) + 1) * sizeof(type)); if (!q_) { rc = PHY_ORACLE_ENOMEM; goto done; } ptr = q_; } while (0)

int phy_oracle_compress_rank(const uint8_t *file, uint64_t size, int np, int rank, uint64_t window_bytes,
                             uint64_t block_bytes, uint32_t record_cap, phy_oracle_rank *out) {
  memset(out, 0, sizeof(*out));
  int rc = PHY_ORACLE_OK;
  int64_t overlap = 500;
  int64_t region = (int64_t)(size / (uint64_t)np);
  int64_t wr_start = (int64_t)rank * region;
  int64_t wr_end = (rank != np - 1) ? wr_start + region + overlap - 1 : (int64_t)size - 1;
  int64_t rsize = (int64_t)window_bytes;
  if (region < rsize) { rsize = wr_end - wr_start + 1; if (rank == np - 1) overlap = 0; }
  int64_t curr = 0; /* r_buffer_curr_pos: window-relative on the first pass, absolute afterwards (as in the reference) */
  if (rank != 0) { /* phyNGSC.cpp:131-156 */
    const uint8_t *b = file + wr_start;
    int64_t c = 0, first_at = 0, lim = (int64_t)size - wr_start;
    while (c < lim && b[c] != '@') ++c;
    first_at = c;
    while (c < lim && b[c] != '\n') ++c;
    if (c + 1 >= lim) return PHY_ORACLE_EMALFORMED;
    curr = (b[c + 1] == '@') ? c + 1 : first_at;
  }
  uint32_t rec_start = (uint32_t)curr;
  out->wr_overlap = (int32_t)curr;
  int bewr = ceil_log2_u((uint64_t)np);

  buf_t sbs = {0, 0, 0}, blks = {0, 0, 0};
  uint8_t *wbuf = (uint8_t *)malloc(block_bytes + 16);
  uint32_t *sbol = 0; uint32_t nsbol = 0, sbol_room = 0;
  int beso = 0, bcss = 0;
  uint64_t written = 0;
  int64_t bytes_read = 0;
  int64_t win_abs = wr_start;
  if (!wbuf) { rc = PHY_ORACLE_ENOMEM; goto done; }
  GROW(out->sb_off, 0, uint64_t); out->sb_off[0] = 0;
  GROW(out->blk_off, 0, uint64_t); out->blk_off[0] = 0;

  while (bytes_read < region) {
    phy_oracle_subblock sb;
    uint64_t readable = size - (uint64_t)win_abs;
    rc = phy_oracle_compress_window(file + win_abs, readable, rsize, rec_start, (int32_t)overlap, record_cap, &sb);
    if (rc) goto done;
    uint32_t k = out->n_subblocks;
    GROW(out->sb_off, k + 1, uint64_t); GROW(out->sb_records, k, uint32_t); GROW(out->sb_win_off, k, uint64_t);
    GROW(out->sb_win_len, k, uint64_t); GROW(out->sb_rec_start, k, uint32_t); GROW(out->sb_overlap, k, int32_t);
    { uint32_t(*q_)[4] = (uint32_t(*)[4])realloc(out->sb_len, ((size_t)k + 1) * sizeof(*q_)); if (!q_) { rc = PHY_ORACLE_ENOMEM; phy_oracle_subblock_free(&sb); goto done; } out->sb_len = q_; }
    if (buf_put(&sbs, sb.payload, sb.payload_len)) { rc = PHY_ORACLE_ENOMEM; phy_oracle_subblock_free(&sb); goto done; }
    out->sb_off[k + 1] = sbs.len; out->sb_records[k] = sb.n_records; out->sb_win_off[k] = (uint64_t)win_abs;
    out->sb_win_len[k] = (uint64_t)rsize; out->sb_rec_start[k] = rec_start; out->sb_overlap[k] = (int32_t)overlap;
    memcpy(out->sb_len[k], sb.len, sizeof(sb.len));
    out->n_subblocks = k + 1;

    /* phyNGSC.cpp:745-755: note bytes_read counts from the window start of the first pass, i.e. it
     * includes rec_start of ranks > 0 only through the seq/title positions being window-relative */
    bytes_read += (int64_t)sb.bytes_consumed;
    int64_t next_abs = bytes_read + wr_start;
    if (next_abs + rsize > wr_end) { if (rank == np - 1) overlap = 0; rsize = wr_end - next_abs; }
    rec_start = 0;

    /* block assembly, phyNGSC.cpp:842-903 */
    uint32_t n = sb.payload_len;
    if (nsbol == sbol_room) { sbol_room = sbol_room ? sbol_room * 2 : 64; uint32_t *q = (uint32_t *)realloc(sbol, sbol_room * sizeof(uint32_t)); if (!q) { rc = PHY_ORACLE_ENOMEM; phy_oracle_subblock_free(&sb); goto done; } sbol = q; }
    sbol[nsbol++] = n;
    { uint32_t mx = 0; for (uint32_t i = 0; i < nsbol; ++i) if (sbol[i] > mx) mx = sbol[i]; beso = bitlen_u(mx); }
    uint64_t hsz = ((uint64_t)bewr + 18 + (uint64_t)beso * nsbol + 7 + 7) / 8; /* ceil((BEWR+6+12 + BESO*n + 5+2)/8) */
    if (written + n + hsz > block_bytes) {
      bcss |= 1; /* LSBS */
      uint64_t fill = block_bytes - (written + hsz);
      sbol[nsbol - 1] = (uint32_t)fill;
      uint8_t hdr[4096];
      uint32_t hl = phy_oracle_make_header(rank, bewr, (int)hsz, beso, bcss, sbol, nsbol, hdr, sizeof(hdr));
      if (hl != hsz) { rc = PHY_ORACLE_EUNSUPPORTED; phy_oracle_subblock_free(&sb); goto done; }
      if (buf_put(&blks, hdr, hl) || buf_put(&blks, wbuf, written) || buf_put(&blks, sb.payload, fill)) { rc = PHY_ORACLE_ENOMEM; phy_oracle_subblock_free(&sb); goto done; }
      GROW(out->blk_off, out->n_blocks + 1, uint64_t);
      out->blk_off[++out->n_blocks] = blks.len;
      written = n - fill;
      memcpy(wbuf, sb.payload + fill, written);
      bcss |= 2; bcss &= ~1; /* FSBS stays set for all later blocks of the rank */
      nsbol = 0; sbol[nsbol++] = (uint32_t)written;
    } else {
      memcpy(wbuf + written, sb.payload, n);
      written += n;
    }
    phy_oracle_subblock_free(&sb);
    win_abs = next_abs;
  }
  out->last_block_size = (uint32_t)written;
  if (written > 0) { /* phyNGSC.cpp:910-928 */
    uint64_t hsz = ((uint64_t)bewr + 18 + (uint64_t)beso * nsbol + 7 + 7) / 8;
    uint8_t hdr[4096];
    uint32_t hl = phy_oracle_make_header(rank, bewr, (int)hsz, beso, bcss, sbol, nsbol, hdr, sizeof(hdr));
    if (hl == 0 || buf_put(&blks, hdr, hl) || buf_put(&blks, wbuf, written)) { rc = PHY_ORACLE_ENOMEM; goto done; }
    GROW(out->blk_off, out->n_blocks + 1, uint64_t);
    out->blk_off[++out->n_blocks] = blks.len;
    out->last_block_size = (uint32_t)(written + hl);
  }
  out->sb_bytes = sbs.p; sbs.p = 0;
  out->blk_bytes = blks.p; blks.p = 0;
done:
  free(sbs.p); free(blks.p); free(wbuf); free(sbol);
  if (rc) phy_oracle_rank_free(out);
  return rc;
}

void phy_oracle_rank_free(phy_oracle_rank *r) {
  free(r->sb_bytes); free(r->sb_off); free(r->sb_records); free(r->sb_win_off); free(r->sb_win_len);
  free(r->sb_rec_start); free(r->sb_overlap); free(r->sb_len); free(r->blk_bytes); free(r->blk_off);
  memset(r, 0, sizeof(*r));
}
