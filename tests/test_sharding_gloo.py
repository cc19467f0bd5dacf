"""CPU, world_size 2 over gloo: the multi-GPU path of bench.py / SURVEY.md 8(e) without GPUs.

Every rank owns a contiguous range of chunks (engine.shard_range), compresses it on its own (here with the
oracle standing in for the device, which is the checker's job in a CPU test), and no collective touches the
data path: only sizes and timings are gathered.  Rank 0 then checks that the concatenation of the ranks'
outputs in rank order is exactly what one Compress() over the whole buffer produces, and that it inflates
back to the input."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
SEG = 59460


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_bytes, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import oracle_lib as O
    from bitar_b200 import synth
    from bitar_b200.engine import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.lineitem_like(n_bytes)                # same bytes on every rank
    n = (data.size + SEG - 1) // SEG
    first, last = shard_range(n, rank, world)
    mine = data[first * SEG:min(data.size, last * SEG)]
    slots, produced = O.compress_buffer(mine, SEG, threads=1)
    # the only exchange: per-rank sizes (and, in bench.py, the timing) -- never the data
    counts = torch.tensor([last - first, int(produced.sum())], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, counts)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks, as bench.py does for the time
    assert float(t[0]) == world
    np.save(os.path.join(ret, f"slots{rank}.npy"), slots)
    np.save(os.path.join(ret, f"prod{rank}.npy"), produced)
    if rank == 0:
        np.save(os.path.join(ret, "gathered.npy"), torch.stack(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_chunks_without_data_collective(tmp_path):
    sys.path.insert(0, HERE)
    import oracle_lib as O
    from bitar_b200 import synth
    from bitar_b200.engine import shard_range
    world, n_bytes = 2, 41 * SEG + 1234
    mp.spawn(_worker, args=(world, _free_port(), n_bytes, str(tmp_path)), nprocs=world, join=True)
    data = synth.lineitem_like(n_bytes)
    n = (n_bytes + SEG - 1) // SEG
    whole_slots, whole_prod = O.compress_buffer(data, SEG, threads=2)
    g = np.load(tmp_path / "gathered.npy")
    assert g[:, 0].sum() == n and g[:, 1].sum() == whole_prod.sum()
    prod = np.concatenate([np.load(tmp_path / f"prod{r}.npy") for r in range(world)])
    slots = np.concatenate([np.load(tmp_path / f"slots{r}.npy") for r in range(world)])
    assert np.array_equal(prod, whole_prod)
    for i in range(n):
        assert np.array_equal(slots[i, :prod[i]], whole_slots[i, :prod[i]])
    out, _ = O.decompress_buffer(slots, prod, SEG)
    assert np.array_equal(out, data)
    # ranges tile [0, n) for every world size the bench is launched with
    for w in (1, 2, 4, 8):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
