#!/bin/bash
# Round capture on a one-GPU box (run through gpurun): GPU parity tests, both bench arms, the ncu launch list of the bench
# command and `ncu --set full` passes over the record kernels of three shapes.  usage: profiles/capture.sh <tag>
# Everything lands in gpurun_out/<tag>_*; profiles/summarize_capture.sh <tag> turns it into the tracked summaries.
TAG=${1:-cap}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${TAG}_bench.json
# launch list of the bench command on a 2 GB image (same code path, two batches per step; the 16 GB default would replay
# ~7000 launches one by one under ncu), without the legs that launch no kernels of ours
LIST_ARGS="--mb 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-shapes --driver-mb 0"
timeout 600 python bench.py $LIST_ARGS > $O/${TAG}_bench_short.json 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py $LIST_ARGS > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
# full captures are summarised here on the box (gpurun brings back at most 64 MiB): only the 100 bp report itself travels
for sh in 36bp 100bp var50_205; do
  profiles/capture_full.sh $TAG $sh
  python profiles/ncu_table.py $O/${TAG}_full_${sh}.ncu-rep > $O/${TAG}_ncu_full_summary_${sh}.txt
  python profiles/ncu_summary.py $O/${TAG}_full_${sh}.ncu-rep > $O/${TAG}_ncu_full_metrics_${sh}.txt
  [ $sh = 100bp ] || rm -f $O/${TAG}_full_${sh}.ncu-rep
done
