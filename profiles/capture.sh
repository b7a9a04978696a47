#!/bin/bash
# Round capture on the GPU box (run through gpurun): GPU parity tests, both bench arms, the ncu launch list of the
# bench command and one `ncu --set full` pass over the heavy kernels.  usage: profiles/capture.sh <tag>
# Everything lands in gpurun_out/<tag>_*; profiles/summarize_capture.sh <tag> turns it into the tracked summaries.
TAG=${1:-cap}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${TAG}_bench.json
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tests/gpu_prof_target.py 36bp 256 2 > $O/${TAG}_plain.log 2>&1 &&
# one subblock group, so that every launch covers the whole 256 MB shard (the traffic figures in ncu_traffic.json are per shard)
PHY_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(stat1|stat2|qhist|lengths|emit|nl_emit|nl_count)' -s 7 -c 7 -f -o $O/${TAG}_full \
  python tests/gpu_prof_target.py 36bp 256 2 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
