# Round-end GPU pass: full GPU test suite, bench line, smoke, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
T=${1:-s18}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
timeout 200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json
python -c "import __graft_entry__ as g; g.smoke()"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_b.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_b.log
