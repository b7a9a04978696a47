TAG=r02t; O=gpurun_out; mkdir -p $O
LIST_ARGS="--mb 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-shapes --driver-mb 0"
timeout 600 python bench.py $LIST_ARGS > $O/${TAG}_bench_short.json 2>$O/${TAG}_bench_short.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py $LIST_ARGS > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
for sh in 36bp 100bp var50_205; do
  profiles/capture_full.sh $TAG $sh
  python profiles/ncu_table.py $O/${TAG}_full_${sh}.ncu-rep > $O/${TAG}_ncu_full_summary_${sh}.txt
  python profiles/ncu_summary.py $O/${TAG}_full_${sh}.ncu-rep > $O/${TAG}_ncu_full_metrics_${sh}.txt
  [ $sh = 100bp ] || rm -f $O/${TAG}_full_${sh}.ncu-rep
done
du -sh $O
