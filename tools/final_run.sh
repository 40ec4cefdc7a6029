set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s15_tests.log 2>&1; tail -3 gpurun_out/s15_tests.log
timeout 300 python bench.py > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err; tail -c 600 gpurun_out/s15_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s15_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s15_ncu_b.log 2>&1; tail -2 gpurun_out/s15_ncu_b.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
