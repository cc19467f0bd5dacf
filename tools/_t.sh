python -m pytest tests/test_gpu_deflate.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -1
for i in 1 2; do python tools/gpu_deflate_prof.py 1024 2>&1 | head -4 | cut -c1-215; done
python tools/gpu_deflate_prof.py 256 4096 2>&1 | head -1 | cut -c1-200
timeout 300 python tools/gpu_fuzz.py 3 3000 2>&1 | tail -1
