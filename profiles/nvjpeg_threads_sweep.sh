for t in 1 4 8; do
  CLIPB200_NVJPEG_THREADS=$t timeout 300 python bench_configs.py --configs 1 --images 6000 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('threads', $t, {k:round(v['images_per_s']) for k,v in d['clipb200'].items()})"
done
