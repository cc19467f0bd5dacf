set -x
python tools/chunk_sweep.py 256 1,8 > gpurun_out/sweep_k.jsonl 2> gpurun_out/sweep_k.err
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/gpu_roundtrip_once.py 256 59460 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_indexed' -c 2 -o gpurun_out/r1k_kernels -f python tools/gpu_roundtrip_once.py 256 59460 > gpurun_out/ncu_full1.log 2>&1
python tools/gpu_roundtrip_once.py 256 4096 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_indexed_kernel<9, 864, 7, 256, 128, 11' -c 2 -o gpurun_out/r1k_small -f python tools/gpu_roundtrip_once.py 256 4096 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
