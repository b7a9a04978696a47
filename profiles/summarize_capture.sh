#!/bin/bash
# Turns the files profiles/capture.sh left in gpurun_out/ into the tracked summaries.  usage: profiles/summarize_capture.sh <tag>
TAG=${1:-cap}
O=gpurun_out
cp $O/${TAG}_bench.json profiles/${TAG}_bench_1gb.json
cp $O/${TAG}_bench_ref.json profiles/${TAG}_bench_reference_arm.json
cp $O/${TAG}_launches.csv profiles/${TAG}_launches.csv
python profiles/summarize_launches.py $O/${TAG}_launches.csv "${TAG}: ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu-baseline (1 GB 36 bp shard); first 205 launches = the kernel-only region (3 warm-up + 2 timed resident steps x (5 + 3 groups x 12) launches), the rest of the list is the 64 MiB batches of the end-to-end leg" 205 > profiles/${TAG}_launch_list_summary.txt
python profiles/ncu_summary.py $O/${TAG}_full.ncu-rep > profiles/${TAG}_ncu_full_summary.txt
tail -2 $O/${TAG}_pytest_gpu.log > profiles/${TAG}_pytest_gpu.txt
