for shape in "8192 7680 8192 0 0 0 256" "8192 7680 8192 0 0 0 192" "16384 1536 384 1 0 0 256" "4096 3072 768 1 0 0 256" "4096 768 3072 0 1 0 256" "4096 2304 768 0 0 0 256"; do
  for pair in 1 2; do SUNET_GEMM_PAIR=$pair timeout 60 ./build/test_gemm one $shape 20 2>&1 | grep "us " | sed 's/bias1 //'; done
done
SUNET_GEMM_PAIR=2 SUNET_GEMM_DBG=4 ./build/test_gemm one 8192 7680 8192 0 0 0 256 10 | grep "us "
