// ref_kat.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" hooks into the reference's own
// HuffmanEncoder / BitStream classes (compiled from /root/reference, never copied), used to pin
// oracle/phy_oracle.c's Huffman build and tree serialisation against the real thing.
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <vector>
#include "defs.h"
#include "bit_stream.h"
#include "huffman.h"

extern "C" int ref_huffman(const uint32_t *freq, uint32_t n, int compact, uint32_t *code, uint32_t *len,
                           uint8_t *tree_out, uint32_t tree_cap, uint32_t *tree_len) {
  HuffmanEncoder enc(n);
  for (uint32_t i = 0; i < n; ++i) enc.Insert(freq[i]);
  HuffmanEncoder::Code *c = enc.Complete(compact != 0);
  if (!c) return -1;
  for (uint32_t i = 0; i < n; ++i) { code[i] = c[i].code; len[i] = c[i].len; }
  BitStream bs;
  bs.Create(1);
  HuffmanEncoder::StoreTree(bs, enc);
  std::vector<uchar> v = bs.GetIO_Buffer();
  *tree_len = (uint32_t)v.size();
  if (v.size() > tree_cap) return -2;
  memcpy(tree_out, v.data(), v.size());
  return 0;
}

// ops: sequence of (kind, value, nbits): 0=PutBits 1=PutBit 2=Put2Bits 3=PutByte 4=PutWord 5=Flush
extern "C" int ref_bitstream(const uint32_t *ops, uint32_t n_ops, uint8_t *out, uint32_t cap, uint32_t *out_len) {
  BitStream bs;
  bs.Create(1);
  for (uint32_t i = 0; i < n_ops; ++i) {
    uint32_t k = ops[3 * i], v = ops[3 * i + 1], b = ops[3 * i + 2];
    switch (k) {
      case 0: bs.PutBits(v, (int32)b); break;
      case 1: bs.PutBit(v); break;
      case 2: bs.Put2Bits(v); break;
      case 3: bs.PutByte((uchar)v); break;
      case 4: bs.PutWord(v); break;
      case 5: bs.FlushPartialWordBuffer(); break;
      default: return -1;
    }
  }
  bs.FlushPartialWordBuffer();
  std::vector<uchar> v = bs.GetIO_Buffer();
  *out_len = (uint32_t)v.size();
  if (v.size() > cap) return -2;
  memcpy(out, v.data(), v.size());
  return 0;
}

// ---- round trip through the reference's own decoder (tasks.cpp:625-1101) ---------------------------------
// The reference ships the Fetch* functions but no main that calls them (phyNGSD.cpp is missing); this hook strings
// them together the way the encoder's stream order dictates (info | title | quality | dna, phyNGSC.cpp:717-742,
// 809-838) and re-inserts the ambiguity codes that the encoder moved into the quality bytes (phyNGSC.cpp:549-588).
// Returns the number of FASTQ bytes written to `out`, or a negative code.  Inputs that hit the reference's
// "tree written by the encoder?" asymmetry cannot be decoded by the reference itself.
#include "structures.h"
#include "tasks.h"

static long long decode_subblock(const uint8_t *payload, uint32_t len, uint8_t *out, unsigned long long cap);
extern "C" long long ref_decode_subblock(const uint8_t *payload, uint32_t len, uint8_t *out, unsigned long long cap) {
  try { return decode_subblock(payload, len, out, cap); } catch (...) { return -5; } /* the reference decoder ran off its input */
}
static long long decode_subblock(const uint8_t *payload, uint32_t len, uint8_t *out, unsigned long long cap) {
  BitStream bs;
  bs.Create(1);
  bs.SetIO_Buffer((uchar *)payload, len, 0);
  bs.SetIO_Buffer_Pos(0);
  uint32 no_records, max_qlen, max_slen, no_symbols, qmode, no_qualities, flags;
  bs.GetWord(no_records); bs.GetWord(max_qlen); bs.GetWord(max_slen);
  bs.GetByte(no_symbols); bs.GetByte(qmode); bs.GetByte(no_qualities);
  bs.GetWord(flags);
  bs.FlushInputWordBuffer();
  (void)max_slen; (void)qmode;
  Record *records = new Record[no_records];
  uint32 qbits = BitStream::BitLength(max_qlen);
  if ((flags & FLAG_VARIABLE_LENGTH) != 0) {
    for (uint32 i = 0; i < no_records; ++i) { uint32 q = 0; if (qbits) bs.GetBits(q, qbits); records[i].qua_len = (int32)q; }
    bs.FlushInputWordBuffer();
  } else {
    for (uint32 i = 0; i < no_records; ++i) records[i].qua_len = (int32)max_qlen;
  }
  std::vector<Field> fields;
  std::vector<uchar> symbols, qualities;
  std::vector<uint32> no_ambiguity;
  FetchTitleHeader(bs, fields);
  FetchTitleBody(bs, fields, records, no_records, flags);
  FetchQuality(bs, qualities, no_qualities, records, no_records, flags, no_ambiguity, max_qlen);
  FetchDNA(bs, symbols, no_symbols, flags, records, no_records, no_ambiguity);
  static const char amb_of_code[17] = {0, 0, 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N', 'X', 'U', '.', '-'};
  unsigned long long o = 0;
  long long rc = 0;
  for (uint32 i = 0; i < no_records && rc == 0; ++i) {
    Record &r = records[i];
    const size_t ql = r.quality.size() ? r.quality.size() - 1 : 0; /* without the '\n' */
    if (o + r.title.size() + 2 * (ql + 1) + 2 > cap) { rc = -2; break; }
    memcpy(out + o, r.title.data(), r.title.size()); o += r.title.size(); /* ends with its '\n' separator */
    size_t k = 0;
    for (size_t j = 0; j < ql; ++j) {
      uchar q = r.quality[j];
      if (q >= 128) {
        uint32 x = (uint32)q - 128 + 16;
        if ((x >> 3) > 16) { rc = -3; break; }
        out[o++] = (uint8_t)amb_of_code[x >> 3];
        r.quality[j] = (uchar)(33 + (x & 7));
      } else {
        if (k + 1 >= r.dna_seq.size() + 0 && k >= r.dna_seq.size()) { rc = -4; break; }
        out[o++] = r.dna_seq[k++];
      }
    }
    out[o++] = '\n'; out[o++] = '+'; out[o++] = '\n';
    memcpy(out + o, r.quality.data(), ql); o += ql;
    out[o++] = '\n';
  }
  delete[] records;
  return rc ? rc : (long long)o;
}
