#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/c12_pytest.log; cat gpurun_out/c12_pytest.log
timeout 600 python tools/ab_variants.py --steps 20 nbprod0 base 2>&1 | tee gpurun_out/c12_ab.log | grep -E "^==|fused|gemm|tail"
