# A/B: mlp_proj_fused with two MMA-issuing threads (default build: warp 1 fc1, 19th warp proj + fc2) vs one (-DSUNET_MLP_SPLIT=0)
# r03: 3.21 ms vs 3.30 ms per forward (6710 vs 6644 images/s)
python bench.py --no-cpu-baseline > gpurun_out/bench_mlp_split1.json 2>/dev/null
SUNET_NVCC_EXTRA=-DSUNET_MLP_SPLIT=0 python -m sunet_tf_b200._build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/bench_mlp_split0.json 2>/dev/null
python -c "
import json
for f in ('0','1'):
    d=json.load(open('gpurun_out/bench_mlp_split'+f+'.json')); print(f, round(d['value'],1), d['kernels']['mlp_fused']['ms'])
"
