# bring-up ablations of the tcgen05 GEMM pipeline (timing only): SUNET_GEMM_DBG bit 1 = no epilogue, 2 = no MMA, 4 = no TMA loads
for shape in "16384 1536 384 1 0 0 256" "262144 96 96 0 1 0 96" "65536 192 192 0 1 0 192"; do
  for dbg in 0 1 2 4 7; do
    echo -n "dbg=$dbg "; SUNET_GEMM_DBG=$dbg ./build/test_gemm one $shape 20 | grep "us " | sed 's/bias1.*grid=148//'
  done
done
