# Round-2 GPU evidence pass (one B200): full GPU test suite, the three bench configurations + the reference arm, smoke,
# the ncu launch list of the cfg-3 bench command and ncu --set full captures of the top kernels.  Outputs: gpurun_out/r2z_*
mkdir -p gpurun_out
T=r2z
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_tests.log 2>&1; grep -E "passed|failed" gpurun_out/${T}_tests.log | tail -2
for c in cfg3 cfg2 cfg4; do
  timeout 400 python bench.py --config $c > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; tail -c 200 gpurun_out/${T}_bench_$c.json; echo
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lp > gpurun_out/${T}_ncu_cfg3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_cfg4.csv python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edgeconv2 -c 4 -o gpurun_out/${T}_edgeconv python tools/prof_edgeconv.py 32 4096 20 1 > gpurun_out/${T}_ncu_ec.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 1 -o gpurun_out/${T}_knn_k20 python tools/time_knn.py 16 ncu > gpurun_out/${T}_ncu_knn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 5 -c 1 -o gpurun_out/${T}_knn_k40 python tools/time_knn.py 16 ncu > gpurun_out/${T}_ncu_knn40.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowgemm_ws -s 27 -c 1 -o gpurun_out/${T}_rowgemm_ws python tools/time_rowgemm.py > gpurun_out/${T}_ncu_rowgemm.log 2>&1
timeout 200 python tools/time_rowgemm.py > gpurun_out/${T}_time_rowgemm_ws.log 2>&1
WSPC_ROWGEMM_KERNEL=serial timeout 200 python tools/time_rowgemm.py > gpurun_out/${T}_time_rowgemm_serial.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:lpb_matvec -s 4 -c 1 -o gpurun_out/${T}_lp python tools/time_lp_blocks.py 8 4096 2.0 > gpurun_out/${T}_ncu_lp.log 2>&1
timeout 200 python tools/time_lp_blocks.py 64 4096 2.0 > gpurun_out/${T}_time_lp.log 2>&1
timeout 200 python tools/prof_edgeconv.py 128 4096 20 3 > gpurun_out/${T}_time_edgeconv.log 2>&1
ls -la gpurun_out | tail -30
