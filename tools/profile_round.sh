#!/bin/bash
# The round's profile set, run on the GPU box (gpurun -- 'bash tools/profile_round.sh [ncu]').  Every ncu pass runs the
# same command that has just exited 0 without ncu; outputs land in gpurun_out/ and are summarised into profiles/ here
# (tools/ncu_traffic.py, on the CPU box) under the round's tag.
TAG=${TAG:-r02e}
set -x
python tools/chunk_sweep.py 256 1,8 > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
if [ "$1" = "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-extras --no-config4 --no-cpu > gpurun_out/${TAG}_bench_pre_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-config4 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/gpu_roundtrip_once.py 256 59460 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_tok' -c 2 -o gpurun_out/${TAG}_kernels -f python tools/gpu_roundtrip_once.py 256 59460 > gpurun_out/${TAG}_ncu_full1.log 2>&1
python tools/gpu_foreign_inflate.py 1024 1 0 > gpurun_out/${TAG}_foreign_1GiB.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:inflate_spec -c 1 -o gpurun_out/${TAG}_spec -f python tools/gpu_foreign_inflate.py 1024 1 0 > gpurun_out/${TAG}_ncu_full3.log 2>&1
python tools/gpu_roundtrip_once.py 256 4096 && ncu --set full --clock-control none --import-source on -k regex:'deflate_kernel|inflate_tok' -c 2 -o gpurun_out/${TAG}_small -f python tools/gpu_roundtrip_once.py 256 4096 > gpurun_out/${TAG}_ncu_full2.log 2>&1
fi
tail -c 600 gpurun_out/${TAG}_bench.json
