SWEEP_BUFFERS=pinned SWEEP_SEGS=65536 python tools/chunk_sweep.py 256 1,8 2>&1 | cut -c60-160
SWEEP_BUFFERS=pinned SWEEP_SEGS=65536 python tools/chunk_sweep.py 1024 1,8 2>&1 | cut -c60-160
for i in 1 2; do python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"; done
