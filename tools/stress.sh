#!/bin/bash
# repeat the contention-heavy parity tests (integer times: few hot cells)
for i in 1 2 3 4 5 6; do
  timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "hello or lanl or burst" --timeout 60 2>&1 | tail -1
done
