for bn in 256 192 128; do for dbg in 0 4 5; do echo -n "bn=$bn dbg=$dbg "; SUNET_GEMM_DBG=$dbg ./build/test_gemm one 8192 7680 8192 0 0 0 $bn 10 | grep "us " | sed 's/bias1.*grid=148//'; done; done
