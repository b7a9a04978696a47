/*
 * fastq_synth.c -- deterministic synthetic FASTQ generator for the workloads named in
 * BASELINE.json / SURVEY.md section 8(d).  Plain C, no dependencies; the same seed and shape
 * always give the same bytes, so the CPU reference, the oracle and the CUDA path can all be
 * fed identical input without shipping large files.
 *
 * Shapes (SURVEY.md 8(d) table and Appendix A regression shapes):
 *   1  36 bp, ERR005195-like titles, 1 % reads with one N (quality '!'), 70 % 'I' quality
 *   3  100 bp, SRR/HWI titles, 41-symbol quality, 5 % reads with an N run whose quality is '#'
 *  30  as 3 but the N run gets quality 'F' (> 40: no transfer, 5-symbol Huffman DNA)
 *   4  150 bp paired-style titles (mates interleaved), 0.5 % N, 4-bin quality
 *   5  50-250 bp variable length, 17-field titles, skewed quality
 *  50  as 5 but read length capped at 205 bp (records stay below the 500-byte overlap)
 *  60  title stress: delta-coded numeric with small deltas, block-constant variable-length strings,
 *      a long (120-140 char) field, a zero-padded "numeric", a small-range value-coded numeric
 *  61  numeric field degrading to string part-way, 1-symbol quality, 1-symbol DNA
 *  62  3-symbol DNA (no G), 2-symbol quality, mixed transferable / non-transferable ambiguity codes
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { /* splitmix64 */
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32); }
static inline uint32_t rng_range(rng_t *r, uint32_t lo, uint32_t hi) { return lo + rng_below(r, hi - lo + 1); }

static inline uint8_t *put_str(uint8_t *p, const char *s) { size_t n = strlen(s); memcpy(p, s, n); return p + n; }
static inline uint8_t *put_u(uint8_t *p, uint64_t v) {
  char tmp[24]; int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = (uint8_t)tmp[--n];
  return p;
}
static inline uint8_t *put_u_pad(uint8_t *p, uint32_t v, int width) {
  for (int i = width - 1; i >= 0; --i) { p[i] = (uint8_t)('0' + v % 10); v /= 10; }
  return p + width;
}
static inline uint8_t *put_bases(uint8_t *p, rng_t *r, uint32_t n, const char *alpha, uint32_t na) {
  uint32_t i = 0;
  while (i < n) {
    uint64_t w = rng_next(r);
    for (int k = 0; k < 16 && i < n; ++k, ++i, w >>= 4) *p++ = (uint8_t)alpha[(w & 15) % na];
  }
  return p;
}

#define MAX_REC 2048

/* Writes one record for `shape` with running index idx (0-based) at p; returns bytes written. */
static size_t gen_record(int shape, rng_t *r, uint64_t idx, uint8_t *p0) {
  uint8_t *p = p0;
  uint8_t seq[512], qual[512];
  uint32_t L = 0;
  switch (shape) {
    case 1: {
      p = put_str(p, "@ERR005195."); p = put_u(p, idx + 1);
      p = put_str(p, " BGI-FC30BFTAAXX_5_1_"); p = put_u_pad(p, rng_below(r, 301), 3);
      *p++ = ':'; p = put_u(p, rng_range(r, 1, 2047)); p = put_str(p, "/2\n");
      L = 36;
      put_bases(seq, r, L, "ACGT", 4);
      for (uint32_t i = 0; i < L; ++i) qual[i] = (rng_below(r, 100) < 70) ? 'I' : (uint8_t)rng_range(r, '#', 'I');
      if (rng_below(r, 100) == 0) { uint32_t k = rng_below(r, L); seq[k] = 'N'; qual[k] = '!'; }
      break;
    }
    case 3: case 30: {
      p = put_str(p, "@SRR1770413."); p = put_u(p, idx + 1);
      p = put_str(p, " HWI-ST1234:100:C0ABCACXX:3:"); p = put_u(p, rng_range(r, 1101, 2316));
      *p++ = ':'; p = put_u(p, rng_range(r, 1000, 20000));
      *p++ = ':'; p = put_u(p, rng_range(r, 1000, 200000)); p = put_str(p, " length=100\n");
      L = 100;
      put_bases(seq, r, L, "ACGT", 4);
      for (uint32_t i = 0; i < L; ++i) qual[i] = (uint8_t)rng_range(r, '!', 'I');
      if (rng_below(r, 100) < 5) {
        uint32_t run = rng_range(r, 1, 10), at = rng_below(r, L - run + 1);
        for (uint32_t i = 0; i < run; ++i) { seq[at + i] = 'N'; qual[at + i] = (shape == 3) ? '#' : 'F'; }
      }
      break;
    }
    case 4: {
      static const char *bc[4] = {"ACGTACGT", "TTAGGCAA", "CCGGAATT", "GATCGATC"};
      static uint32_t lane, tile, x, y, code; /* mates share coordinates */
      if ((idx & 1) == 0) {
        lane = rng_range(r, 1, 4); tile = rng_range(r, 1101, 2678); x = rng_range(r, 1000, 32000);
        y = rng_range(r, 1000, 36000); code = rng_below(r, 4);
      }
      p = put_str(p, "@A00123:45:HXXXXDSXX:"); p = put_u(p, lane); *p++ = ':'; p = put_u(p, tile);
      *p++ = ':'; p = put_u(p, x); *p++ = ':'; p = put_u(p, y); *p++ = ' ';
      *p++ = (idx & 1) ? '2' : '1'; p = put_str(p, ":N:0:"); p = put_str(p, bc[code]); *p++ = '\n';
      L = 150;
      put_bases(seq, r, L, "ACGT", 4);
      for (uint32_t i = 0; i < L; ++i) {
        uint32_t u = rng_below(r, 100);
        qual[i] = u < 85 ? 'F' : u < 93 ? ':' : u < 98 ? ',' : '#';
      }
      if (rng_below(r, 200) == 0) { uint32_t k = rng_below(r, L); seq[k] = 'N'; qual[k] = '#'; }
      break;
    }
    case 5: case 50: {
      L = rng_range(r, 50, shape == 5 ? 250 : 205);
      p = put_str(p, "@M00123:45:000000000-A1B2C:1:"); p = put_u(p, rng_range(r, 1101, 2119));
      *p++ = ':'; p = put_u(p, rng_range(r, 1000, 29000)); *p++ = ':'; p = put_u(p, rng_range(r, 1000, 29000));
      *p++ = ' '; *p++ = (uint8_t)('1' + rng_below(r, 2)); p = put_str(p, ":N:0:");
      p = put_bases(p, r, 6, "ACGT", 4);
      p = put_str(p, " sample=XYZ_"); p = put_u(p, rng_range(r, 1, 99));
      p = put_str(p, " len="); p = put_u(p, L); *p++ = '\n';
      put_bases(seq, r, L, "ACGT", 4);
      for (uint32_t i = 0; i < L; ++i) {
        if (rng_below(r, 100) < 60) qual[i] = 'I';
        else { uint32_t q = 'H'; while (q > '#' && (rng_next(r) & 1)) --q; qual[i] = (uint8_t)q; }
      }
      break;
    }
    case 60: {
      /* @run7:<idx*3 + small jitter>_<word repeated in runs of 40>.<long field>,007#<100..140> */
      static const char *words[5] = {"alpha", "be", "gammaray", "d", "epsilon"};
      static const char *w;
      static uint64_t ctr;
      if (idx == 0) ctr = 1000;
      ctr += rng_range(r, 1, 4);
      if (idx % 40 == 0) w = words[rng_below(r, 5)];
      p = put_str(p, "@run7:"); p = put_u(p, ctr); *p++ = '_';
      p = put_str(p, w);
      *p++ = '.';
      { uint32_t n = rng_range(r, 120, 140); p = put_bases(p, r, n, "abcdefghij", 10); }
      *p++ = ','; p = put_u_pad(p, rng_below(r, 50), 3);
      *p++ = '#'; p = put_u(p, rng_range(r, 100, 140)); *p++ = '\n';
      L = rng_range(r, 30, 60);
      put_bases(seq, r, L, "ACGT", 4);
      for (uint32_t i = 0; i < L; ++i) qual[i] = (uint8_t)rng_range(r, '5', '?');
      break;
    }
    case 61: {
      p = put_str(p, "@deg ");
      if (idx == 12345 % 1000 + 200) p = put_str(p, "0042"); else p = put_u(p, 40 + (idx % 7));
      *p++ = ':'; p = put_u(p, 7 + idx * 2); *p++ = '\n';
      L = 51;
      memset(seq, 'A', L); memset(qual, 'B', L);
      break;
    }
    case 62: {
      static const char amb[] = "NYRWSKMDVHBXU.-";
      p = put_str(p, "@mix_"); p = put_u(p, idx / 3); *p++ = '/'; p = put_u(p, 1 + idx % 3); *p++ = '\n';
      L = rng_range(r, 20, 45);
      put_bases(seq, r, L, "ACT", 3);
      for (uint32_t i = 0; i < L; ++i) qual[i] = (rng_next(r) & 1) ? '%' : ';';
      uint32_t kind = rng_below(r, 20);
      if (kind < 3) { /* transferable: ambiguity codes whose quality is within '!'..'(' */
        uint32_t n = rng_range(r, 1, 4);
        for (uint32_t i = 0; i < n; ++i) { uint32_t k = rng_below(r, L); seq[k] = (uint8_t)amb[rng_below(r, 15)]; qual[k] = '%'; }
      } else if (kind == 3) { /* not transferable: quality too high at one ambiguous base */
        uint32_t k = rng_below(r, L); seq[k] = 'N'; qual[k] = ';';
      }
      break;
    }
    default: return 0;
  }
  memcpy(p, seq, L); p += L; *p++ = '\n'; *p++ = '+'; *p++ = '\n';
  memcpy(p, qual, L); p += L; *p++ = '\n';
  return (size_t)(p - p0);
}

/*
 * Fill out[0..cap) with whole records of `shape` until at least target_bytes are produced
 * (or max_records, if non-zero, is reached).  Returns the byte count, 0 on bad shape / tiny cap.
 * *n_records (optional) receives the record count.
 */
uint64_t phy_synth_fastq(int shape, uint64_t seed, uint64_t target_bytes, uint64_t max_records, uint8_t *out,
                         uint64_t cap, uint64_t *n_records) {
  rng_t r = {seed * 0x2545F4914F6CDD1Dull + (uint64_t)shape};
  uint64_t pos = 0, idx = 0;
  uint8_t tmp[MAX_REC];
  while (pos < target_bytes && (max_records == 0 || idx < max_records)) {
    size_t n = gen_record(shape, &r, idx, tmp);
    if (n == 0 || pos + n > cap) break;
    memcpy(out + pos, tmp, n);
    pos += n; ++idx;
  }
  if (n_records) *n_records = idx;
  return pos;
}
