# pipeline-depth sensitivity of the tcgen05 GEMM: M N K act res f32 bn iters, ring capped by SUNET_GEMM_STAGES
for shape in "8192 7680 8192 0 0 0" "16384 1536 384 1 0 0" "4096 768 3072 0 1 0"; do
  for cfg in "256 3" "256 2" "192 4" "192 3" "128 5" "128 4" "128 3" "128 2"; do
    set -- $cfg
    SUNET_GEMM_STAGES=$2 ./build/test_gemm one $shape $1 20 2>/dev/null | grep "us " | sed 's/bias1.*f320//'
  done
done
