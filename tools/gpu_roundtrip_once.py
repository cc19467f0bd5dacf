"""One Compress() + one Decompress() of the lineitem-like workload, device-resident (the program ncu captures).
usage: python tools/gpu_roundtrip_once.py [MiB] [seg]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bitar_b200 import synth  # noqa: E402
from bitar_b200.engine import CompressDevice, Configuration  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 59460
data = synth.lineitem_like(mib << 20)
n = (data.size + seg - 1) // seg
dev = CompressDevice(0, 1).Initialize(Configuration(decompressed_seg_size=seg, max_preallocate_memzones=n + 8))
src = torch.from_numpy(data).cuda()
out = torch.zeros(n * seg + 64, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ops, slots = dev.compress_ops(src.data_ptr(), data.size)
res = dev.enqueue("deflate", 0, ops)
dev.wait(0)
kd, _ = dev.last_ms(0)
iops = dev.decompress_ops(slots, res["produced"], out.data_ptr())
dev.enqueue("inflate", 0, iops)
dev.wait(0)
ki, _ = dev.last_ms(0)
ok = bool((out[:data.size] == src).all().item())
print(f"{mib} MiB seg {seg}: deflate {kd:.3f} ms ({data.size / kd / 1e6:.1f} GB/s), inflate {ki:.3f} ms "
      f"({data.size / ki / 1e6:.1f} GB/s), ratio {data.size / float(res['produced'].sum()):.3f}, ok={ok}")
dev.close()
sys.exit(0 if ok else 1)
