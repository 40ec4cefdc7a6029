for i in 1 2 3 4 5 6; do timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED|^E  " | cut -c1-300 | tail -6; done
