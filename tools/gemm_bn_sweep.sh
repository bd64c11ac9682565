# tile-width sweep on the stage-2/3 GEMM shapes: M N K act res f32 bn iters
for shape in "16384 1152 384 0 0 0" "16384 384 384 0 1 0" "16384 1536 384 1 0 0" "16384 384 1536 0 1 0" "4096 2304 768 0 0 0" "4096 768 768 0 1 0" "4096 3072 768 1 0 0" "4096 768 3072 0 1 0" "16384 768 1536 0 0 0" "65536 384 768 0 0 0"; do
  for bn in 256 192 128 96 64; do ./build/test_gemm one $shape $bn 30 2>/dev/null | grep "us " | sed 's/bias1.*f320//'; done
done
