# pair (cta_group::2) vs single-CTA tiles on the stage-2/3 shapes: M N K act res f32 bn iters [pair]
for shape in "16384 1152 384 0 0 0" "16384 384 384 0 1 0" "16384 1536 384 1 0 0" "16384 384 1536 0 1 0" "4096 2304 768 0 0 0" "4096 768 768 0 1 0" "4096 3072 768 1 0 0" "4096 768 3072 0 1 0" "8192 8192 8192 0 0 0"; do
  ./build/test_gemm one $shape 0 20 | grep "us "
  for bn in 256 192 128; do SUNET_GEMM_PAIR=2 ./build/test_gemm one $shape $bn 20 | grep "us "; done
done
