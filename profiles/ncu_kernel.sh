# one ncu --set full capture of the kernels matching $1 (regex) on shape $2: bash profiles/ncu_kernel.sh "k_(emit|stat1)" 36bp <skip> <count>
set -e
K=${1:-'k_(lengths|emit)'}
SHAPE=${2:-36bp}
python tests/gpu_prof_target.py $SHAPE 256 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${3:-2} -c ${4:-2} -f -o gpurun_out/prof_cur python tests/gpu_prof_target.py $SHAPE 256 2 > gpurun_out/ncu_cur.log 2>&1
