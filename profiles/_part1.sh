TAG=r02t; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $O/${TAG}_bench.json
