#!/bin/bash
# A/B of tuning knobs on the embed bench (20-step burst + >= 1.2 s sustained region in one run).
#   bash profiles/ab_embed.sh "CLIPB200_PDL=1" "CLIPB200_PDL=0 CLIPB200_GEMM_STAGES=4" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --workload embed --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$i.json 2>/dev/null
  python - "$cfg" gpurun_out/ab_$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r, s = d["roofline"], d["sustained"]
print(f"{sys.argv[1]:48s} burst {d['value']:8.0f} img/s  e2e {d['e2e']['value']:8.0f}  sustained {s['value']:8.0f} @ {s['clocks']['sm_mhz']:.0f} MHz  "
      f"one-lane gemm {r['gemm_ms_per_step']:.3f} ms ({r['achieved']:.0f} TF)  sust one-lane gemm {s['roofline']['gemm_ms_per_step']:.3f} ms")
PY
done
