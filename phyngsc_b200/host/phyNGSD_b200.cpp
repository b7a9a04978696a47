/*
 * phyNGSD_b200 -- decompressor for .ngsc files: `./phyNGSD_b200 in.ngsc out.fastq [threads]`.
 *
 * The reference's Makefile names a phyNGSD.cpp (Makefile:7,18-19) that is not in its tree; its building blocks are
 * ReadFooter (tasks.cpp:1203-1292), the block header layout of MakeHeader (tasks.cpp:1179-1200) and the Fetch*
 * functions (tasks.cpp:625-1101).  This program reads the container the same way -- footer at the end of the file,
 * blocks in file order keyed by WRID, subblocks that were split across two blocks of a rank (BCSS bits LSBS / FSBS,
 * phyNGSC.cpp:857-894) joined again -- and decodes every subblock with phy_decode.hpp on a pool of threads
 * (subblocks are independent; rank order, then subblock order, is file order of the FASTQ).
 * The container is memory-mapped, subblocks are decoded at most a bounded window ahead of the writer and written in order
 * as they complete, so neither the .ngsc nor the FASTQ is ever held in memory as a whole.
 * Plain C++: decoding needs no GPU and no MPI.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "phy_decode.hpp"

namespace {

struct Image { const uint8_t *p; uint64_t n; uint64_t size() const { return n; } uint8_t operator[](uint64_t i) const { return p[i]; } const uint8_t *data() const { return p; } };

struct Bits { /* MSB-first reader over the whole file image */
  const Image &d; uint64_t bit;
  Bits(const Image &data, uint64_t byte_pos) : d(data), bit(byte_pos * 8) {}
  uint64_t get(unsigned n) {
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i, ++bit) {
      if ((bit >> 3) >= d.size()) throw phydec::Error{"container truncated"};
      v = (v << 1) | ((d[bit >> 3] >> (7 - (bit & 7))) & 1u);
    }
    return v;
  }
  void align() { bit = (bit + 7) & ~7ull; }
  uint64_t byte_pos() const { return bit >> 3; }
};

int ceil_log2(uint64_t x) { int b = 0; while ((1ull << b) < x) ++b; return b; }

struct Piece { uint64_t off, len; };

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.ngsc out.fastq [threads]\n", argv[0]); return 1; }
  const int threads = argc > 3 ? atoi(argv[3]) : (int)std::thread::hardware_concurrency();
  const int fd = open(argv[1], O_RDONLY);
  struct stat sb;
  if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size < 4) { fprintf(stderr, "[E] cannot read %s\n", argv[1]); return 2; }
  void *map = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map == MAP_FAILED) { fprintf(stderr, "[E] cannot map %s\n", argv[1]); return 2; }
  const Image data = {(const uint8_t *)map, (uint64_t)sb.st_size};
  try {
    /* footer (tasks.cpp:1104-1176 / 1203-1292): the last two bytes hold its length */
    const uint64_t flen = ((uint64_t)data[data.size() - 2] << 8) | data[data.size() - 1];
    if (flen + 2 > data.size()) throw phydec::Error{"bad footer length"};
    const uint64_t fstart = data.size() - 2 - flen;
    Bits r(data, fstart);
    const unsigned BEPS = (unsigned)r.get(4), BEFS = (unsigned)r.get(6), BEBS = (unsigned)r.get(4), BESS = (unsigned)r.get(4), BELB = (unsigned)r.get(5),
                   BEOV = (unsigned)r.get(4), LBES = (unsigned)r.get(1);
    const uint64_t np = r.get(BEPS);
    const uint64_t fastq_size = BEFS > 32 ? (r.get(BEFS - 32) << 32) | r.get(32) : r.get(BEFS);
    const uint64_t n_blocks = r.get(BEBS), n_subblocks = r.get(BESS);
    (void)BELB; (void)BEOV; (void)LBES; (void)n_subblocks;
    if (np == 0 || np > 32768) throw phydec::Error{"bad rank count"};
    /* blocks in file order; a block is header + sum(SBOL) (tasks.cpp:1179-1200) */
    std::vector<std::vector<Piece>> subs(np); /* per rank: subblocks as lists of byte ranges ... */
    std::vector<std::vector<uint32_t>> parts(np); /* ... how many ranges form each (1, or 2+ when split across blocks) */
    std::vector<char> open_split(np, 0);
    uint64_t pos = 0, nb = 0;
    const int bewr = ceil_log2(np);
    while (pos < fstart) {
      Bits h(data, pos);
      const uint64_t wrid = h.get((unsigned)bewr), bhs = h.get(12), nosb = h.get(6), beso = h.get(5), bcss = h.get(2);
      if (wrid >= np) throw phydec::Error{"bad block header"};
      std::vector<uint64_t> sbol(nosb);
      for (auto &x : sbol) x = h.get((unsigned)beso);
      h.align();
      if (h.byte_pos() - pos != bhs) throw phydec::Error{"bad block header"};
      uint64_t p = pos + bhs;
      for (uint64_t i = 0; i < nosb; ++i) {
        if (p + sbol[i] > fstart) throw phydec::Error{"block runs into the footer"};
        if (i == 0 && open_split[wrid]) { subs[wrid].push_back({p, sbol[i]}); parts[wrid].back()++; }
        else { subs[wrid].push_back({p, sbol[i]}); parts[wrid].push_back(1); }
        open_split[wrid] = (i + 1 == nosb) && (bcss & 1u); /* LSBS: the last subblock continues in the rank's next block */
        p += sbol[i];
      }
      if (nosb == 0) open_split[wrid] = open_split[wrid] && (bcss & 1u);
      pos = p; ++nb;
    }
    if (pos != fstart || nb != n_blocks) throw phydec::Error{"blocks do not tile the file"};
    for (uint64_t w = 0; w < np; ++w) if (open_split[w]) throw phydec::Error{"dangling split subblock"};
    /* flatten: payload list in rank order */
    struct Job { std::vector<Piece> pieces; std::string text; const char *err = nullptr; bool ready = false; };
    std::vector<Job> jobs;
    for (uint64_t w = 0; w < np; ++w) {
      size_t k = 0;
      for (uint32_t n : parts[w]) { Job j; for (uint32_t i = 0; i < n; ++i) j.pieces.push_back(subs[w][k++]); jobs.push_back(std::move(j)); }
    }
    /* workers decode at most `window` subblocks ahead of the writer; the writer (this thread) writes them in order and frees them */
    const int nthreads = threads < 1 ? 1 : threads;
    const size_t window = (size_t)nthreads * 3 + 2;
    std::mutex m; std::condition_variable cv;
    size_t next = 0, written = 0;
    bool stop = false;
    auto work = [&]() {
      std::vector<uint8_t> joined;
      for (;;) {
        size_t i;
        {
          std::unique_lock<std::mutex> g(m);
          cv.wait(g, [&] { return stop || next >= jobs.size() || next < written + window; });
          if (stop || next >= jobs.size()) return;
          i = next++;
        }
        Job &j = jobs[i];
        const uint8_t *p = data.data() + j.pieces[0].off;
        size_t n = j.pieces[0].len;
        if (j.pieces.size() > 1) {
          joined.clear();
          for (auto &pc : j.pieces) joined.insert(joined.end(), data.data() + pc.off, data.data() + pc.off + pc.len);
          p = joined.data(); n = joined.size();
        }
        try { j.text.reserve(n * 6); phydec::decode_subblock(p, n, j.text, (size_t)256 << 20); } catch (const phydec::Error &e) { j.err = e.what; } catch (...) { j.err = "decoder failure"; }
        { std::lock_guard<std::mutex> g(m); j.ready = true; }
        cv.notify_all();
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(work);
    FILE *o = fopen(argv[2], "wb");
    uint64_t total = 0;
    const char *fail = o ? nullptr : "cannot create the output";
    size_t fail_at = 0;
    for (size_t i = 0; i < jobs.size() && !fail; ++i) {
      { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return jobs[i].ready; }); }
      Job &j = jobs[i];
      if (j.err) { fail = j.err; fail_at = i; break; }
      size_t n = j.text.size();
      /* a FASTQ that does not end in a newline: the decoder closes its last record with one; the footer knows the size */
      if (i + 1 == jobs.size() && total + n == fastq_size + 1 && n && j.text[n - 1] == '\n') --n;
      if (fwrite(j.text.data(), 1, n, o) != n) { fail = "write error"; fail_at = i; break; }
      total += n;
      std::string().swap(j.text);
      { std::lock_guard<std::mutex> g(m); written = i + 1; }
      cv.notify_all();
    }
    { std::lock_guard<std::mutex> g(m); stop = true; }
    cv.notify_all();
    for (auto &t : pool) t.join();
    if (o) fclose(o);
    if (fail) { fprintf(stderr, "[E] subblock %zu: %s\n", fail_at, fail); return o ? 4 : 2; }
    printf("[I] %s: %llu rank(s), %llu block(s), %zu subblock(s) -> %llu bytes of FASTQ%s\n", argv[1], (unsigned long long)np, (unsigned long long)nb,
           jobs.size(), (unsigned long long)total, total == fastq_size ? "" : " (footer states a different size!)");
    return total == fastq_size ? 0 : 5;
  } catch (const phydec::Error &e) {
    fprintf(stderr, "[E] %s: %s\n", argv[1], e.what);
    return 4;
  }
}
