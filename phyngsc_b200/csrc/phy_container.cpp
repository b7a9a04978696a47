/*
 * phy_container.cpp -- block header and footer of the .ngsc container, host side.
 * The layouts are the ones MakeHeader / MakeFooter write (tasks.cpp:1104-1200, structures.h:310-333);
 * the implementation is a plain MSB-first bit packer over a caller-provided buffer.
 */
#include <stdint.h>
#include <string.h>

#include "../../include/phyngsc_b200.h"

namespace {
struct Packer {
  uint8_t *p; uint32_t cap; uint64_t bit; bool over;
  Packer(uint8_t *out, uint32_t c) : p(out), cap(c), bit(0), over(false) { memset(out, 0, c); }
  void put(uint64_t v, unsigned n) { /* n <= 64, MSB first */
    for (unsigned i = n; i-- > 0;) {
      uint64_t byte = bit >> 3;
      if (byte >= cap) { over = true; return; }
      if ((v >> i) & 1u) p[byte] |= (uint8_t)(0x80u >> (bit & 7));
      ++bit;
    }
  }
  void align() { bit = (bit + 7) & ~7ull; }
  uint32_t bytes() const { return (uint32_t)((bit + 7) >> 3); }
};
unsigned bitlen(uint64_t x) { unsigned b = 0; while (x) { ++b; x >>= 1; } return b; }          /* floor(log2 x) + 1 */
unsigned ceil_log2(uint64_t x) { unsigned b = 0; while ((1ull << b) < x) ++b; return b; }     /* x >= 1 */
}  // namespace

/* tasks.cpp:1179-1200: WRID (BEWR bits) | BHS 12 | NOSB 6 | BESO 5 | BCSS 2 | NOSB x SBOL (BESO bits) | align */
extern "C" uint32_t phy_make_block_header(int32_t wrid, int32_t bewr, int32_t bhs, int32_t beso, int32_t bcss,
                                          const uint32_t *sbol, uint32_t nosb, uint8_t *out, uint32_t cap) {
  if (!out || (!sbol && nosb) || bewr < 0 || bewr > 31 || beso < 0 || beso > 31) return 0;
  Packer w(out, cap);
  w.put((uint32_t)wrid, (unsigned)bewr);
  w.put((uint32_t)bhs, 12);
  w.put(nosb, 6);
  w.put((uint32_t)beso, 5);
  w.put((uint32_t)bcss, 2);
  for (uint32_t i = 0; i < nosb; ++i) w.put(sbol[i], (unsigned)beso);
  w.align();
  return w.over ? 0 : w.bytes();
}

/* tasks.cpp:1104-1176 */
extern "C" int32_t phy_make_footer(int32_t np, uint64_t fastq_size, uint32_t n_blocks, uint32_t n_subblocks,
                                   const int32_t *overlaps, const int32_t *block_order, const uint32_t *lb_sizes,
                                   uint8_t *out, uint32_t cap) {
  if (np < 1 || !overlaps || !block_order || !lb_sizes || !out) return PHY_ERR_ARG;
  uint32_t lb_max = 0, lb_min = 0xFFFFFFFFu; int32_t ov_max = 0;
  for (int i = 0; i < np; ++i) {
    if (lb_sizes[i] > lb_max) lb_max = lb_sizes[i];
    if (lb_sizes[i] < lb_min) lb_min = lb_sizes[i];
    if (overlaps[i] > ov_max) ov_max = overlaps[i];
  }
  /* the reference evaluates log2(0) when every rank starts exactly on a record */
  if (ov_max <= 0 || lb_max == 0 || n_blocks == 0 || n_subblocks == 0) return PHY_ERR_UNSUPPORTED;
  const unsigned BEPS = bitlen((uint64_t)np), BEFS = bitlen(fastq_size), BEBS = bitlen(n_blocks), BESS = bitlen(n_subblocks);
  const unsigned BELB = bitlen(lb_max), BEOV = bitlen((uint64_t)ov_max), LBES = lb_max == lb_min ? 1u : 0u;
  Packer w(out, cap);
  w.put(BEPS, 4); w.put(BEFS, 6); w.put(BEBS, 4); w.put(BESS, 4); w.put(BELB, 5); w.put(BEOV, 4); w.put(LBES, 1);
  w.put((uint64_t)np, BEPS);
  w.put(fastq_size, BEFS);
  w.put(n_blocks, BEBS);
  w.put(n_subblocks, BESS);
  for (int i = 1; i < np; ++i) w.put((uint32_t)overlaps[i], BEOV);
  const unsigned cbo = ceil_log2((uint64_t)np);
  for (uint32_t i = 0; i < n_blocks; ++i) w.put((uint32_t)block_order[i], cbo);
  if (!LBES) for (int i = 0; i < np; ++i) w.put(lb_sizes[i], BELB);
  w.align();
  const uint32_t flen = w.bytes();
  w.put(flen, 16);
  if (w.over) return PHY_ERR_CAPACITY;
  return (int32_t)w.bytes();
}

/* ---- subblock decoder (host/phy_decode.hpp) behind the C ABI ------------------------------------------------- */
#include "../host/phy_decode.hpp"

extern "C" int64_t phy_decode_subblock(const uint8_t *payload, uint64_t len, uint8_t *out, uint64_t cap) {
  if (!payload || !out) return PHY_ERR_ARG;
  try {
    std::string text;
    text.reserve((size_t)len * 6);
    phydec::decode_subblock(payload, (size_t)len, text);
    if (text.size() > cap) return PHY_ERR_CAPACITY;
    memcpy(out, text.data(), text.size());
    return (int64_t)text.size();
  } catch (const phydec::Error &) {
    return PHY_ERR_MALFORMED;
  } catch (...) {
    return PHY_ERR_MALFORMED;
  }
}
