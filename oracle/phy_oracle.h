/*
 * phy_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C) of phyNGSC's per-subblock FASTQ compression path and of the host
 * plumbing around it (working-region partition, window chaining, block assembly, header, footer).
 * It is the checker the CUDA path is compared against; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library never does.
 *
 * Parity status: PINNED against the reference itself, compiled unmodified from /root/reference
 * into oracle/_ref/ (see oracle/Makefile) and run in this container -- the reference ships no
 * golden vectors or tests of its own.  tests/test_oracle_vs_reference.py
 * holds the comparison; tests/golden/ holds reference-generated fixtures for boxes without
 * /root/reference.
 */
#ifndef PHY_ORACLE_H
#define PHY_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  PHY_ORACLE_OK = 0,
  PHY_ORACLE_EMALFORMED = -1, /* '+' line missing / truncated record                      */
  PHY_ORACLE_EFIELDS = -2,    /* separator count differs between titles (reference: UB)    */
  PHY_ORACLE_ECOLORSPACE = -3,
  PHY_ORACLE_EUNSUPPORTED = -4, /* max_len == 128 (Q11), code length > 32, NUL bytes ...    */
  PHY_ORACLE_ENOMEM = -5,
  PHY_ORACLE_ECAP = -6
};

typedef struct {
  uint32_t n_records;
  uint64_t bytes_consumed; /* window offset of the byte after the last taken record         */
  uint32_t len[4];         /* info, title, quality, dna section lengths                      */
  uint8_t *payload;        /* malloc'ed: info | title | quality | dna                        */
  uint32_t payload_len;
  uint32_t warnings;       /* bit0: record cap hit                                           */
} phy_oracle_subblock;

/* Huffman: build + serialise exactly like HuffmanEncoder::Complete / StoreTree.
 * tree_out receives [word mem_size][mem bytes]; returns bytes written (0 on error). */
uint32_t phy_oracle_huffman(const uint32_t *freq, uint32_t n, int compact, uint32_t *code, uint32_t *len,
                            uint8_t *tree_out, uint32_t tree_cap);

/* no_threads of the reference for the calls that follow (default 1); see split_records in phy_oracle.c */
void phy_oracle_set_threads(int threads);

/* One subblock.  win[0..readable) must be addressable; r_buffer_size is the reference's window
 * size (drives the stop rule), readable >= r_buffer_size allows the read-slack semantics (Q4). */
int phy_oracle_compress_window(const uint8_t *win, uint64_t readable, int64_t r_buffer_size, uint32_t rec_start,
                               int32_t overlap, uint32_t record_cap, phy_oracle_subblock *out);
void phy_oracle_subblock_free(phy_oracle_subblock *sb);

typedef struct {
  /* subblocks of the rank, in order */
  uint32_t n_subblocks;
  uint8_t *sb_bytes;     /* concatenated payloads                                         */
  uint64_t *sb_off;      /* n_subblocks + 1 offsets into sb_bytes                          */
  uint32_t *sb_records;
  uint64_t *sb_win_off;  /* absolute file offset of each window                           */
  uint64_t *sb_win_len;  /* r_buffer_size used for each window                            */
  uint32_t *sb_rec_start;
  int32_t *sb_overlap;
  uint32_t (*sb_len)[4];
  /* blocks of the rank, in order */
  uint32_t n_blocks;
  uint8_t *blk_bytes;
  uint64_t *blk_off;     /* n_blocks + 1                                                   */
  /* footer inputs */
  uint32_t last_block_size;
  int32_t wr_overlap;
} phy_oracle_rank;

/* Everything rank `rank` of `np` does to file[0..size): partition, chain windows, compress,
 * assemble 8 MiB blocks.  window_bytes/block_bytes are the reference's READ/WRITE_BUFFER_SIZE
 * (8 MiB); smaller values give small multi-subblock / multi-block cases for tests. */
int phy_oracle_compress_rank(const uint8_t *file, uint64_t size, int np, int rank, uint64_t window_bytes,
                             uint64_t block_bytes, uint32_t record_cap, phy_oracle_rank *out);
void phy_oracle_rank_free(phy_oracle_rank *r);

/* Block header as MakeHeader writes it; returns bytes written. */
uint32_t phy_oracle_make_header(int wrid, int bewr, int bhs, int beso, int bcss, const uint32_t *sbol, uint32_t nosb,
                                uint8_t *out, uint32_t cap);
/* Footer as MakeFooter writes it for the given block order (rank id per block, file order). */
int32_t phy_oracle_make_footer(int np, uint64_t fastq_size, uint32_t n_blocks, uint32_t n_subblocks,
                               const int32_t *overlaps, const int32_t *block_order, const uint32_t *lb_sizes,
                               uint8_t *out, uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif
