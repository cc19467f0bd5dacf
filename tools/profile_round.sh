#!/bin/bash
# The round's profile set, run on the GPU box (gpurun -- 'bash tools/profile_round.sh').  Every ncu pass runs the
# same command that has just exited 0 without ncu; outputs land in gpurun_out/ and are summarised into profiles/.
set -x
python tools/chunk_sweep.py 256 1,8 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/gpu_roundtrip_once.py 256 59460 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_indexed' -c 2 -o gpurun_out/r1_kernels -f python tools/gpu_roundtrip_once.py 256 59460 > gpurun_out/ncu_full1.log 2>&1
python tools/gpu_roundtrip_once.py 256 4096 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_indexed' -c 2 -o gpurun_out/r1_small -f python tools/gpu_roundtrip_once.py 256 4096 > gpurun_out/ncu_full2.log 2>&1
fi
tail -c 600 gpurun_out/bench.json
