timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
SWEEP_BUFFERS=pinned SWEEP_SEGS=4096,65536,1048576 python tools/chunk_sweep.py 256 1,2,8 2>&1 | cut -c1-160
SWEEP_BUFFERS=pinned SWEEP_SEGS=65536 python tools/chunk_sweep.py 1024 1,8 2>&1 | cut -c1-160
python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_l.json; python -c "
import json
d=json.load(open('gpurun_out/bench_l.json')); print({k:d[k] for k in ('value','deflate_gbps','inflate_gbps','gpu_launches') if k in d}, d['e2e']['value'], d['e2e']['pcie'])"
