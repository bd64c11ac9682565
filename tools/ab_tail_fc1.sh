# A/B: tail_up_fused with two fc1-issuing threads (even / odd sub-pixels; default) vs one (-DSUNET_TAIL_FC1_SPLIT=0).  r03: 182 vs 207 us
for i in 1 2; do python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench_fc1s1_$i.json 2>/dev/null; done
SUNET_NVCC_EXTRA=-DSUNET_TAIL_FC1_SPLIT=0 python -m sunet_tf_b200._build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q -k "model or upsample or u8 or eval" 2>&1 | tail -2
for i in 1 2; do python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench_fc1s0_$i.json 2>/dev/null; done
python -c "
import json
for f in ('0_1','0_2','1_1','1_2'):
    d=json.load(open('gpurun_out/bench_fc1s'+f+'.json')); print(f, round(d['value'],1), round(d['e2e']['value'],1), d['kernels']['tail_up_fused']['ms'])
"
