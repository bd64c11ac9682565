#!/bin/bash
# correctness of the checked cases, then timing of the model's GEMM shapes and the big square
timeout 120 ./build/test_gemm 2>&1 | grep -E "FAIL|error|rc=|ok" | awk '{print "   " $0}' | tail -16
for shape in "16384 1536 384 1 0 0 0" "16384 384 1536 0 1 0 0" "16384 1152 384 0 0 0 0" "4096 3072 768 1 0 0 0" "4096 768 3072 0 1 0 0" "4096 2304 768 0 0 0 0" "4096 768 768 0 1 0 0" \
             "65536 192 384 0 0 0 0" "262144 96 192 0 0 0 0" "65536 384 192 2 0 0 0" "8192 7680 8192 0 0 0 256" "8192 7680 8192 0 0 0 192" "8192 7680 8192 0 0 0 128"; do
  timeout 60 ./build/test_gemm one $shape 20 2>&1 | grep "us " | sed 's/bias1 //'
done
