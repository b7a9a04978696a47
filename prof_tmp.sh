set -e
python tests/gpu_prof_target.py 36bp 256 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_(nl_emit|stat1|qhist|stat2|lengths|emit)' -s 6 -c 6 -f -o gpurun_out/prof_r01e python tests/gpu_prof_target.py 36bp 256 2 > gpurun_out/ncu_e.log 2>&1
