/*
 * phy_decode.hpp -- decoder of one phyNGSC subblock payload (info | title | quality | dna), the inverse of the
 * compression path in csrc/.  SURVEY section 8(f) "next #2": the reference ships Fetch* functions but no program that
 * calls them (phyNGSD.cpp is missing from the reference tree), and those functions cannot read what the reference's
 * own encoder writes for value-coded numeric title fields without a table.  This decoder mirrors the
 * ENCODER's rule for "is there a table?" (phy_core.cuh classify_subblock, tasks.cpp:338 + structures.h:126-202), so
 * every payload the compressor can produce round-trips.
 *
 * Stream layout: SURVEY Appendix A (phyNGSC.cpp:717-742, 809-838; tasks.cpp:302-509, 519-559, 576-621;
 * huffman.cpp:88-118, 191-205).  Host-only C++, no CUDA: decoding one subblock is a sequential walk over three bit
 * streams; subblocks decode independently (the callers run them on a thread pool).
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace phydec {

struct Error { const char *what; };

struct BitReader { /* MSB-first, bit_stream.h:203-208, 325-355 */
  const uint8_t *p; size_t n, pos; uint32_t cur, left;
  BitReader(const uint8_t *d, size_t len) : p(d), n(len), pos(0), cur(0), left(0) {}
  uint32_t byte_raw() { if (pos >= n) throw Error{"payload truncated"}; return p[pos++]; }
  void align() { left = 0; }
  uint32_t bit() { if (!left) { cur = byte_raw(); left = 8; } return (cur >> --left) & 1u; }
  uint32_t bits(uint32_t k) { uint32_t v = 0; while (k--) v = (v << 1) | bit(); return v; }
  uint32_t byte() { align(); return byte_raw(); }
  uint32_t word() { align(); uint32_t v = 0; for (int i = 0; i < 4; ++i) v = (v << 8) | byte_raw(); return v; }
};

inline uint32_t bit_length(uint32_t x) { /* BitStream::BitLength, bit_stream.h:268-277 (values >= 2^31 do not occur in valid streams) */
  for (uint32_t i = 0; i < 32; ++i) if (x < (1u << i)) return i;
  return 64;
}
inline uint32_t id_bits(uint32_t n) { /* huffman.cpp:96-98 */
  uint32_t b = 0;
  for (uint32_t t = 2; t <= n; t *= 2) ++b;
  if (n & (n - 1)) ++b;
  return b;
}

struct Tree { /* huffman.cpp:88-118, 191-205: [word mem_size][word root][word n][byte min_len][pre-order bits] */
  std::vector<int32_t> left, right; /* per internal node: child >= 0 internal index, < 0: ~leaf id */
  int32_t root = 0;                 /* >= 0 internal, < 0 ~leaf (single-symbol alphabet: codes of zero bits) */
  void load(BitReader &r) {
    const uint32_t mem = r.word();
    const size_t end = r.pos + mem;
    if (end > r.n || mem < 9) throw Error{"bad Huffman table"};
    BitReader t(r.p + r.pos, mem);
    t.word();
    const uint32_t n = t.word();
    t.byte();
    if (n == 0 || n > 65536) throw Error{"bad Huffman table"};
    const uint32_t ib = id_bits(n);
    left.clear(); right.clear();
    root = parse(t, ib, n, 0);
    r.pos = end; r.align();
  }
  int32_t parse(BitReader &t, uint32_t ib, uint32_t n, int depth) {
    if (depth > 40) throw Error{"bad Huffman table"};
    if (t.bit()) { const uint32_t id = ib ? t.bits(ib) : 0u; if (id >= n) throw Error{"bad Huffman table"}; return ~(int32_t)id; }
    const int32_t me = (int32_t)left.size();
    left.push_back(0); right.push_back(0);
    const int32_t l = parse(t, ib, n, depth + 1);
    left[me] = l;
    const int32_t rr = parse(t, ib, n, depth + 1);
    right[me] = rr;
    return me;
  }
  uint32_t decode(BitReader &r) const {
    int32_t v = root;
    while (v >= 0) v = r.bit() ? right[v] : left[v];
    return (uint32_t)~v;
  }
};

struct Field {
  uint8_t sep = 0; bool constant = false, numeric = false, len_const = false, is_delta = false, has_table = false;
  std::string tok0;                 /* constant token / record 0's token of a string field */
  int32_t min_v = 0, max_v = 0, min_d = 0, max_d = 0;
  uint32_t bits_num = 0, bits_val = 0, bits_len = 0, len0 = 0, max_len = 0, min_len = 0;
  Tree num_tab;
  std::vector<bool> same;           /* Hamming mask: position equals record 0's everywhere */
  std::vector<Tree> chr; std::vector<int> chr_of; /* per-position tables: chr_of[min(j,128)] -> index into chr or -1 */
};

inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

/* Decodes one subblock payload; appends the FASTQ text of its records to `out`. */
/* `limit`: most text bytes a subblock may decode to (a window of the reference is 8 MiB of FASTQ; a damaged payload -- e.g. a
 * long constant token times a huge record count -- must not be allowed to ask for unbounded memory) */
inline void decode_subblock(const uint8_t *payload, size_t len, std::string &out, size_t limit = (size_t)1 << 30) {
  static const char amb_of_code[17] = {0, 0, 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N', 'X', 'U', '.', '-'}; /* phyNGSC.cpp:184-206 */
  BitReader r(payload, len);
  /* ---- info (phyNGSC.cpp:717-742) */
  const uint32_t R = r.word(), max_qlen = r.word();
  r.word(); /* longest kept DNA string */
  const uint32_t nsym = r.byte(); r.byte(); const uint32_t nq = r.byte();
  const uint32_t flags = r.word();
  if (R == 0 || (uint64_t)R > (uint64_t)len * 8) throw Error{"bad record count"}; /* every record costs at least one payload bit */
  if (max_qlen > (1u << 20) || (uint64_t)max_qlen * 13 > (uint64_t)len) throw Error{"bad read length"}; /* one table per read position */
  std::vector<uint32_t> qlen(R, max_qlen);
  if (flags & 0x8u) { /* FLAG_VARIABLE_LENGTH: always set by the encoder */
    const uint32_t nb = bit_length(max_qlen);
    for (uint32_t i = 0; i < R; ++i) qlen[i] = nb ? r.bits(nb) : 0u;
    r.align();
  }
  /* ---- title header (tasks.cpp:302-390) */
  const uint32_t nf = r.word();
  if (nf == 0 || nf > 4096) throw Error{"bad field count"};
  uint32_t P = 0; /* value-coded numeric fields below index P carry no table */
  if (nf >= 2) { P = 1; while (P * 2 <= nf - 1) P *= 2; }
  std::vector<Field> F(nf);
  uint32_t nnc = 0;
  for (uint32_t f = 0; f < nf; ++f) {
    Field &x = F[f];
    x.sep = (uint8_t)r.byte();
    x.constant = r.byte() != 0;
    if (x.constant) {
      const uint32_t l = r.word();
      if (l > r.n - r.pos) throw Error{"payload truncated"}; /* the token's bytes follow: a damaged length must not size a buffer */
      x.tok0.resize(l);
      for (uint32_t j = 0; j < l; ++j) x.tok0[j] = (char)r.byte();
      continue;
    }
    ++nnc;
    x.numeric = r.byte() != 0;
    if (x.numeric) {
      x.min_v = (int32_t)r.word(); x.max_v = (int32_t)r.word(); x.min_d = (int32_t)r.word(); x.max_d = (int32_t)r.word();
      const int32_t vr = wsub(x.max_v, x.min_v), dr = wsub(x.max_d, x.min_d);
      x.is_delta = !(vr < dr); /* tasks.cpp:208-217 */
      x.bits_num = bit_length((uint32_t)(x.is_delta ? dr : vr));
      x.bits_val = bit_length((uint32_t)vr);
      const int32_t diff = (x.is_delta ? dr : vr) + 1;
      const bool nonempty = x.is_delta ? R >= 2 : f >= P;
      x.has_table = diff > 0 && diff <= 512 && nonempty; /* tasks.cpp:338 */
      if (x.has_table) x.num_tab.load(r);
      continue;
    }
    x.len_const = r.byte() != 0;
    x.len0 = r.word(); x.max_len = r.word(); x.min_len = r.word();
    if (x.len0 > (1u << 20) || x.max_len > (1u << 20) || x.min_len > x.max_len || x.len0 > r.n - r.pos) throw Error{"bad string field"};
    x.tok0.resize(x.len0);
    for (uint32_t j = 0; j < x.len0; ++j) x.tok0[j] = (char)r.byte();
    x.same.resize(x.len0);
    r.align();
    for (uint32_t j = 0; j < x.len0; ++j) x.same[j] = r.bit() != 0;
    r.align();
    x.bits_len = bit_length(x.max_len - x.min_len);
    const uint32_t nt = x.max_len < 128 ? x.max_len : 128;
    x.chr_of.assign(129, -1);
    for (uint32_t j = 0; j < nt; ++j)
      if (j >= x.len0 || !x.same[j]) { x.chr_of[j] = (int)x.chr.size(); x.chr.emplace_back(); x.chr.back().load(r); }
    if (x.max_len >= 128) { x.chr_of[128] = (int)x.chr.size(); x.chr.emplace_back(); x.chr.back().load(r); }
    r.align();
  }
  /* ---- title body (tasks.cpp:393-509): blocks of 32 records, byte-aligned */
  std::vector<std::string> titles(R);
  std::vector<int32_t> prev(nf, 0);
  std::vector<std::string> blk_tok(nf);
  std::vector<uint8_t> flag(nf, 0);
  size_t title_bytes = 0;
  for (uint32_t lo = 0; lo < R; lo += 32) {
    const uint32_t hi = lo + 32 < R ? lo + 32 : R;
    if (title_bytes > limit) throw Error{"subblock text exceeds the limit"};
    for (uint32_t f = 0; f < nf; ++f) if (!F[f].constant) flag[f] = (uint8_t)r.bit();
    for (uint32_t i = lo; i < hi; ++i) {
      std::string &t = titles[i];
      const bool first = i == lo;
      for (uint32_t f = 0; f < nf; ++f) {
        Field &x = F[f];
        if (x.constant) { t += x.tok0; t += (char)x.sep; continue; }
        if (x.numeric) {
          int32_t v;
          if (first) v = (int32_t)((uint32_t)x.min_v + (x.bits_val ? r.bits(x.bits_val) : 0u));
          else if (flag[f]) v = x.is_delta ? (int32_t)((uint32_t)prev[f] + (uint32_t)x.min_d) : prev[f]; /* block-constant delta == min_delta / value */
          else {
            const uint32_t tcode = x.has_table ? x.num_tab.decode(r) : (x.bits_num ? r.bits(x.bits_num) : 0u);
            v = x.is_delta ? (int32_t)((uint32_t)prev[f] + tcode + (uint32_t)x.min_d) : (int32_t)((uint32_t)x.min_v + tcode);
          }
          prev[f] = v;
          char buf[16]; int k = 0; uint32_t u = (uint32_t)v;
          do { buf[k++] = (char)('0' + u % 10); u /= 10; } while (u);
          while (k) t += buf[--k];
          t += (char)x.sep;
          continue;
        }
        if (!first && flag[f]) { t += blk_tok[f]; t += (char)x.sep; continue; }
        const uint32_t l = x.len_const ? x.len0 : x.min_len + (x.bits_len ? r.bits(x.bits_len) : 0u);
        std::string &tok = blk_tok[f];
        tok.resize(l);
        for (uint32_t j = 0; j < l; ++j) {
          if (j >= x.len0 || !x.same[j]) {
            const int ti = x.chr_of[j < 128 ? j : 128];
            if (ti < 0) throw Error{"character without a table"};
            tok[j] = (char)x.chr[ti].decode(r);
          } else tok[j] = x.tok0[j];
        }
        t += tok; t += (char)x.sep;
      }
      title_bytes += t.size();
    }
    r.align();
  }
  (void)nnc;
  /* ---- quality (tasks.cpp:576-621) */
  std::vector<uint8_t> quals(nq);
  for (uint32_t i = 0; i < nq; ++i) quals[i] = (uint8_t)r.byte();
  std::vector<Tree> qt(max_qlen + 1);
  for (uint32_t p = 0; p <= max_qlen; ++p) qt[p].load(r);
  r.align();
  std::vector<std::string> qual(R);
  std::vector<uint32_t> namb(R, 0);
  for (uint32_t i = 0; i < R; ++i) {
    if (qlen[i] > max_qlen) throw Error{"read longer than the longest read"};
    std::string &q = qual[i];
    q.resize(qlen[i]);
    for (uint32_t k = 0; k < qlen[i]; ++k) {
      const uint32_t c = qt[k + 1].decode(r);
      if (c >= nq) throw Error{"quality code out of range"};
      q[k] = (char)quals[c];
      if (quals[c] >= 128) ++namb[i];
    }
  }
  r.align();
  /* ---- dna (tasks.cpp:519-559) and reassembly with the transferred ambiguity codes (phyNGSC.cpp:549-588) */
  std::vector<uint8_t> syms(nsym);
  for (uint32_t i = 0; i < nsym; ++i) syms[i] = (uint8_t)r.byte();
  r.align();
  const bool plain = (flags & 0x2u) != 0;
  Tree dt;
  if (!plain) dt.load(r);
  r.align();
  for (uint32_t i = 0; i < R; ++i) {
    if (out.size() > limit) throw Error{"subblock text exceeds the limit"};
    out += titles[i]; /* ends with its '\n' separator */
    std::string &q = qual[i];
    for (uint32_t k = 0; k < qlen[i]; ++k) {
      const uint8_t qb = (uint8_t)q[k];
      if (qb >= 128) {
        const uint32_t x = (uint32_t)qb - 128u + 16u;
        if ((x >> 3) > 16 || (x >> 3) < 2) throw Error{"bad transferred quality byte"};
        out += amb_of_code[x >> 3];
        q[k] = (char)(33 + (x & 7));
      } else {
        const uint32_t c = plain ? r.bits(2) : dt.decode(r);
        if (c >= nsym) throw Error{"base code out of range"};
        out += (char)syms[c];
      }
    }
    out += "\n+\n";
    out += q;
    out += '\n';
  }
}

}  // namespace phydec
