"""GPU tuning sweep: device-resident deflate / inflate throughput of the kernels on the synthetic
lineitem workload, for every inflate variant.  Prints one line per configuration.
usage: python tools/gpu_sweep.py [MiB] [seg]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import _capi as capi  # noqa: E402
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    seg = int(sys.argv[2]) if len(sys.argv) > 2 else 59460
    reps = 3
    gens = {"lineitem": lambda n: synth.lineitem_like(n)}
    for name in synth.COLUMNS:
        gens[name] = (lambda nm: (lambda n: synth.column(nm, n)))(name)
    only = os.environ.get("SWEEP_ONLY")
    for wname, gen in gens.items():
        if only and wname != only:
            continue
        n_bytes = mib << 20 if wname == "lineitem" else (mib << 20) // 4
        data = gen(n_bytes)
        n = (data.size + seg - 1) // seg
        dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 8))
        src = torch.from_numpy(data).cuda()
        out = torch.zeros(n * seg + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ops, slots = dev.compress_ops(src.data_ptr(), data.size)
        best = 1e9
        for _ in range(reps + 1):
            res = dev.enqueue("deflate", 0, ops)
            dev.wait(0)
            k, t = dev.last_ms(0)
            best = min(best, k)
        comp = int(res["produced"].sum())
        print(f"[{wname}] deflate seg={seg} n={n} U={data.size} C={comp} ratio={data.size / comp:.3f} "
              f"kernel={best:.3f} ms  {data.size / best / 1e6:.1f} GB/s", flush=True)
        iops = dev.decompress_ops(slots, res["produced"], out.data_ptr())
        for v in [int(x) for x in os.environ.get("SWEEP_VARIANTS", "0,5").split(",")]:
            capi.lib().bitar_tune_inflate_variant(v)
            best = 1e9
            for _ in range(reps + 1):
                ires = dev.enqueue("inflate", 0, iops)
                dev.wait(0)
                k, t = dev.last_ms(0)
                best = min(best, k)
            ok = bool((out[:data.size] == src).all().item()) and int(ires["produced"].sum()) == data.size
            print(f"[{wname}] inflate variant={v} kernel={best:.3f} ms  {data.size / best / 1e6:.1f} GB/s ok={ok}", flush=True)
        capi.lib().bitar_tune_inflate_variant(0)
        dev.close()
        del src, out


if __name__ == "__main__":
    main()
