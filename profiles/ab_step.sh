run() { env "$@" python bench.py --workload embed --steps 400 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['roofline']['achieved']), d['clocks']['sm_mhz'])
"; }
echo "baseline      $(run A=1)"
echo "mainloop-only $(run CLIPB200_GEMM_DEBUG=1)"
echo "no-stores     $(run CLIPB200_GEMM_DEBUG=2)"
echo "no-attention  $(run CLIPB200_SKIP=2)"
echo "no-attn+mainloop $(run CLIPB200_SKIP=2 CLIPB200_GEMM_DEBUG=1)"
echo "baseline2     $(run A=1)"
