# A/B: attn_fused<384> (RING mode) with a separate ring-producer thread (default) vs one producer + issuer thread (-DSUNET_AF_SPLIT=0)
# r03: 49.6 vs 55.9 us per launch (6820 vs 6694 images/s)
python bench.py --no-cpu-baseline --profile-json gpurun_out/kernels_af1.json > gpurun_out/bench_af1.json 2>/dev/null
SUNET_NVCC_EXTRA=-DSUNET_AF_SPLIT=0 python -m sunet_tf_b200._build --force > /dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --profile-json gpurun_out/kernels_af0.json > gpurun_out/bench_af0.json 2>/dev/null
python -c "
import json
for f in ('0','1'):
    d=json.load(open('gpurun_out/bench_af'+f+'.json'))
    ll=json.load(open('gpurun_out/kernels_af'+f+'.json'))['launch_list']
    agg={}
    for n,ms,fl,by in ll:
        if n=='attn_fused':
            a=agg.setdefault(by,[0,0.0]); a[0]+=1; a[1]+=ms*1000
    print(f, round(d['value'],1), {k:round(v[1]/v[0],1) for k,v in agg.items()})
"
