// ref_kat.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" hooks into the reference's own
// HuffmanEncoder / BitStream classes (compiled from /root/reference, never copied), used to pin
// oracle/phy_oracle.c's Huffman build and tree serialisation against the real thing.
#include <stdint.h>
#include <string.h>
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
