# single-CTA vs CTA-pair (cta_group::2) tiles: M N K act res f32 bn iters
for shape in "8192 7680 8192 0 0 0" "16384 1536 384 1 0 0" "16384 384 1536 0 1 0" "4096 3072 768 1 0 0" "4096 768 3072 0 1 0" "4096 2304 768 0 0 0" "16384 1152 384 0 0 0"; do
  for bn in 256 192 128; do
    for pair in 0 2; do
      SUNET_GEMM_PAIR=$pair ./build/test_gemm one $shape $bn 20 2>/dev/null | grep "us " | sed 's/bias1.*f320//'
    done
  done
done
