# A/B of the split MMA issuers (fc1 and fc2 issued by different threads) in tail_up_fused.  r03: 295 -> 224 us.  The same split in
# mlp_proj_fused (plus the 19th warp it needs) measured 3.31 ms against 3.29 ms for the single-issuer 18-warp kernel: not kept.
python bench.py --no-cpu-baseline > gpurun_out/bench_ab_split.json 2> gpurun_out/bench_ab.err
SUNET_TAIL_NO_SPLIT=1 python bench.py --no-cpu-baseline > gpurun_out/bench_ab_nosplit.json 2>/dev/null
python -c "
import json
for f in ('split','nosplit'):
    d=json.load(open('gpurun_out/bench_ab_'+f+'.json')); print(f, round(d['value'],1), d['kernels']['tail_up_fused']['ms'], d['kernels']['mlp_fused']['ms'])
"
