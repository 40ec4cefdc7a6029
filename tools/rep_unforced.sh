for i in 1 2 3 4 5 6 7 8 9 10 11 12; do timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -q -s -k "unforced" 2>&1 | grep -E "un-forced|quantiles|passed|failed|^E  " | cut -c1-330; done
