for i in 1 2 3 4 5 6; do timeout 600 python -m pytest tests/test_baseline_shapes_gpu.py -q -s -k "forced" 2>&1 | grep -E "gradient errors|passed|failed" | python -c "
import sys,re
for l in sys.stdin:
    v=[float(x) for x in re.findall(r\"'([0-9.]+e-[0-9]+)'\", l)]
    print(l[:40].strip(), ('max %.1e' % max(v)) if v else l.strip())
"; done
