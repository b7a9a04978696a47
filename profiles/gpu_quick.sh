# quick GPU check: parity cases + per-stage probe (shape list in $1, default "36bp 100bp")
timeout 300 python tests/gpu_diag.py > gpurun_out/diag.log 2>&1; grep -v " OK " gpurun_out/diag.log | tail -8
for sh in ${1:-36bp 100bp}; do
  timeout 200 python tests/gpu_probe.py $sh 1000 > gpurun_out/probe_$sh.log 2>&1
  grep -h "resident run 2" gpurun_out/probe_$sh.log
  grep -h " ms " gpurun_out/probe_$sh.log | grep -v "resident\|e2e" | awk '{printf "%s=%s ", $1, $2} END {print ""}'
done
