# timing of the GEMM shapes of the SUNet forward (B=64) through build/test_gemm: M N K act res f32 bn iters
for shape in "262144 96 96 0 1 0 0" "65536 192 192 0 1 0 0" "16384 1152 384 0 0 0 0" "16384 384 384 0 1 0 0" "16384 1536 384 1 0 0 0" "16384 384 1536 0 1 0 0" "4096 2304 768 0 0 0 0" "4096 768 768 0 1 0 0" "4096 3072 768 1 0 0 0" "4096 768 3072 0 1 0 0" "262144 1536 96 2 0 0 0" "8192 8192 8192 0 0 0 0"; do
  ./build/test_gemm one $shape 20 | grep "us "
done
