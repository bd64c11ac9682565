"""Turns an ncu CSV launch log (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of one bench.py run
into profiles/ncu_traffic.json: per kernel family, launches, mean duration and mean DRAM bytes per launch.

  python tools/ncu_traffic.py gpurun_out/launches_dram.csv profiles/ncu_traffic.json
"""
import csv
import json
import sys

FAMILIES = (("gemm_tn_f16", "gemm_tcgen05"), ("proj_ln_kernel", "gemm_tcgen05"), ("patch_embed_mma_kernel", "patch_embed_conv"), ("attn_fused_kernel", "attn_fused"), ("attn_core_kernel", "attn_core"), ("attn_core_tc_kernel", "attn_core"), ("mlp_proj_fused", "mlp_fused"), ("mlp_row_kernel", "mlp_fused"),
            ("mlp_fused_kernel", "mlp_fused"), ("layernorm_kernel", "layernorm"), ("patch_embed_kernel", "patch_embed_conv"),
            ("upsample_combine", "upsample_combine"), ("tail_stencil", "tail_stencil"), ("tail_finish", "tail_stencil"), ("tail_up_fused", "tail_up_fused"))

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    rows.append(r)
acc = {}
for r in rows:
    name = r.get("Kernel Name", "")
    fam = next((f for pat, f in FAMILIES if pat in name), None)
    if fam is None:
        continue
    a = acc.setdefault(fam, {"ids": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["ids"].add(r["ID"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
    if r["Metric Name"] == "gpu__time_duration.sum":
        a["ns"] += v * scale
    elif r["Metric Name"] == "dram__bytes_read.sum":
        a["rd"] += v * scale
    elif r["Metric Name"] == "dram__bytes_write.sum":
        a["wr"] += v * scale
out = {"source": sys.argv[1], "note": "ncu replays each kernel cold-cache and serialised: durations are for the share of the step only", "kernels": {}}
tot = sum(a["ns"] for a in acc.values())
for fam, a in sorted(acc.items(), key=lambda kv: -kv[1]["ns"]):
    n = len(a["ids"])
    out["kernels"][fam] = {"launches": n, "mean_us": a["ns"] / n / 1e3, "share": a["ns"] / tot, "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n,
                           "dram_read_per_launch": a["rd"] / n, "dram_write_per_launch": a["wr"] / n}
with open(sys.argv[2], "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out["kernels"], indent=1))
