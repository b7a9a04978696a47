"""INTEGRATION.md patch B as an executable recipe (test infrastructure, like the rest of oracle/).

Applies the reference-side binding a phyNGSC maintainer would add -- the loop body of phyNGSC.cpp:127-840 replaced by ONE call into
the C ABI (include/phyngsc_b200.h), the reference's own block assembly (:842-906), final block (:908-928), footer gathers and
timestamp writer (:930-1057) left as they are -- to the reference's main file WHERE IT LIES under /root/reference, writes the patched
translation unit to a temporary directory (never into the repo), and builds

    oracle/_ref/phyNGSC_patchB      patched phyNGSC.cpp + the reference's tasks.cpp / huffman.cpp / bit_stream.cpp (MakeHeader,
                                    MakeFooter, BitStream are still the reference's) + -lphyngsc_b200, MPI = the fork-based stand-in

The edit is keyed by the reference's line numbers and checked against anchor text, so a different reference revision fails loudly.
tests/test_gpu_driver.py runs the binary on the GPU box and compares its blocks, keyed by rank, with the unmodified reference's.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PHY_REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "phyNGSC_patchB")

# (first line, last line, anchor text that must occur in the first line, replacement) -- 1-based, inclusive
INCLUDE = '#include "phyngsc_b200.h" /* patch B: C ABI of the B200 path, link with -lphyngsc_b200 */\n'

CALL = r'''  /* ---- patch B: the rank's working region goes through the B200 library in one call (replaces :127-160 and the loop body :170-840) */
  phy_ctx *gpu = NULL;
  {
    int ndev = phy_device_count();
    if (ndev < 1 || phy_ctx_create(&gpu, p_rank % ndev, 0, 0) != PHY_OK)
    {
      printf("\n[E] ERROR: p_Rank %d cannot create the GPU context.\n", p_rank);
      MPI_Finalize();
      exit(3);
    }
  }
  uint64_t region_len = (uint64_t)(p_wr_end - p_wr_start + 1);
  if (p_rank != g_size - 1) /* read a little past p_wr_end where the file allows: records longer than the overlap */
  {
    MPI_Offset extra = FASTQ_size - (p_wr_end + 1);
    region_len += (uint64_t)(extra > 65536 ? 65536 : (extra > 0 ? extra : 0));
  }
  uchar *region = (uchar*) phy_host_alloc(region_len + 64);
  uchar *payloads = (uchar*) phy_host_alloc(region_len / 2 + (1 << 20));
  MPI_File_read_at(input_FASTQ, p_wr_start, region, (int)region_len, MPI_CHAR, MPI_STATUS_IGNORE);
  phy_region_params prm;
  prm.file_size = (uint64_t)FASTQ_size; prm.np = g_size; prm.rank = p_rank; prm.window_bytes = READ_BUFFER_SIZE; prm.overlap = 500;
  prm.record_cap = 100000; prm.threads = (uint32_t)no_threads; prm.reserved = 0;
  std::vector<phy_subblock_desc> sb(region_len / (READ_BUFFER_SIZE / 2) + 64);
  uint32_t n_sb = (uint32_t)sb.size();
  phy_region_result res;
  {
    int rc = phy_compress_region(gpu, region, region_len, &prm, payloads, region_len / 2 + (1 << 20), sb.data(), &n_sb, &res);
    if (rc != PHY_OK)
    {
      printf("\n[E] ERROR: p_Rank %d: %s: %s\n", p_rank, phy_strerror(rc), phy_last_error(gpu));
      MPI_Finalize();
      exit(4);
    }
  }
  wr_ov_used = res.wr_overlap;
'''

LOOP_HEAD = r'''  for (uint32_t i_sb = 0; i_sb < n_sb; ++i_sb) /* patch B: one iteration == one iteration of the old while loop */
  {
    ++p_subblock_count;
    if (sb[i_sb].warnings & 1)
      printf("\n[!] WARNING: p_Rank %d subblock %u: records_per_th exceeded.\n", p_rank, i_sb);
    copy_buffer = payloads + sb[i_sb].out_off;   /* info | title | quality | dna, as at :809-838 */
    p_bytes_to_copy = sb[i_sb].out_len;          /* :799 */
    p_bytes_read += sb[i_sb].bytes_consumed;     /* :745 */
'''

LOOP_TAIL = r'''  } /* patch B: copy_buffer is not freed here, it points into `payloads` */
  phy_ctx_destroy(gpu); phy_host_free(region); phy_host_free(payloads);
'''

EDITS = [
    (10, 10, "#include <mpi.h>", "#include <mpi.h>\n" + INCLUDE),
    (127, 160, "read_buffer  = (uchar*) malloc", CALL),
    (166, 841, "// Begin processing FASTQ file", LOOP_HEAD),
    (905, 906, "free(copy_buffer);", LOOP_TAIL),
]


def patched_source():
    src = open(os.path.join(REF, "phyNGSC.cpp")).read().splitlines(keepends=True)
    out, pos = [], 1
    for first, last, anchor, repl in EDITS:
        if anchor not in src[first - 1]:
            raise SystemExit(f"patch B: line {first} of phyNGSC.cpp is not the expected one ({anchor!r}); different reference revision?")
        out += src[pos - 1:first - 1]
        out.append(repl)
        pos = last + 1
    out += src[pos - 1:]
    return "".join(out)


def build(force=False):
    """-> path of the patched binary, or None when the reference sources are not there (the GPU box uses the prebuilt file)."""
    if not os.path.exists(os.path.join(REF, "phyNGSC.cpp")):
        return OUT if os.path.exists(OUT) else None
    lib_dir = os.path.join(ROOT, "phyngsc_b200", "csrc")
    lib = os.path.join(lib_dir, "libphyngsc_b200.so")
    if not os.path.exists(lib):
        return None
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(lib), os.path.getmtime(__file__)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "phyNGSC_patchB.cpp")
        with open(cpp, "w") as f:
            f.write(patched_source())
        cmd = ["g++", "-O3", "-m64", "-fopenmp", "-std=c++11", "-w", "-I" + os.path.join(ROOT, "phyngsc_b200", "host", "mpi_shim"), "-I" + REF,
               "-I" + os.path.join(ROOT, "include"), "-o", OUT, cpp] + [os.path.join(REF, f) for f in ("tasks.cpp", "huffman.cpp", "bit_stream.cpp")] + \
              ["-L" + lib_dir, "-lphyngsc_b200", "-Wl,-rpath,$ORIGIN/../../phyngsc_b200/csrc", "-lpthread"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--show":
        sys.stdout.write(patched_source())
    else:
        print(build(force=True))
