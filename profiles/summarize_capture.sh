#!/bin/bash
# Turns the files profiles/capture.sh left in gpurun_out/ into the tracked summaries.  usage: profiles/summarize_capture.sh <tag>
TAG=${1:-cap}
O=gpurun_out
cp $O/${TAG}_bench.json profiles/${TAG}_bench_n1.json
cp $O/${TAG}_bench_ref.json profiles/${TAG}_bench_reference_arm.json
cp $O/${TAG}_launches.csv profiles/${TAG}_launches.csv
python profiles/summarize_launches.py $O/${TAG}_launches.csv "${TAG}: ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --mb 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-shapes --driver-mb 0 (2 GB image of the default 100 bp shape = two batches per step; the list starts with the kernel-only leg: 5 resident steps, then the 128 MiB batches of the end-to-end leg)" 640 > profiles/${TAG}_launch_list_summary.txt
for sh in 36bp 100bp var50_205; do
  cp $O/${TAG}_ncu_full_summary_${sh}.txt $O/${TAG}_ncu_full_metrics_${sh}.txt profiles/
done
tail -2 $O/${TAG}_pytest_gpu.log > profiles/${TAG}_pytest_gpu.txt
