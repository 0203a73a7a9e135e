"""N>1 host logic on CPU: world_size-2 gloo processes run the row sharding + reference
replication plumbing that bench.py uses under torchrun with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdsctk_b200 import sharding, synth


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, n_total, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    atoms = 7
    begin, count = sharding.shard_range(n_total, world, rank)
    mine = synth.traj_frames(n_total, atoms, 4, 123, begin, count)
    bpf = [atoms * 12, 4]
    arrays = [torch.zeros(n_total * b, dtype=torch.uint8) for b in bpf]
    arrays[0][begin * bpf[0]:(begin + count) * bpf[0]] = torch.from_numpy(mine.reshape(-1).view(np.uint8).copy())
    gid = np.arange(begin, begin + count, dtype=np.int32)
    arrays[1][begin * 4:(begin + count) * 4] = torch.from_numpy(gid.view(np.uint8).copy())
    sharding.replicate_frame_major(arrays, bpf, n_total, world, rank, dist)
    full = synth.traj_frames(n_total, atoms, 4, 123)
    ok = np.array_equal(arrays[0].numpy().view(np.float32).reshape(n_total, atoms, 3), full)
    ok &= np.array_equal(arrays[1].numpy().view(np.int32), np.arange(n_total, dtype=np.int32))
    # every rank's step rows stay inside its shard and all shards tile [0, n_total)
    for s in range(5):
        b, n = sharding.step_rows(begin, count, 3, s)
        ok &= begin <= b and b + n <= begin + count
    t = torch.tensor([count], dtype=torch.int64)
    dist.all_reduce(t)
    ok &= int(t) == n_total
    # bench.py's cross-N output check: every rank hashes the k-lists of ITS rows with their global row numbers, the
    # 64-bit sums are added over the ranks as two 32-bit halves -- the result must not depend on the sharding
    import bench
    rng = np.random.default_rng(5)
    all_d = np.sort(rng.random((n_total, 9)), axis=1)
    all_i = rng.integers(0, n_total, (n_total, 9)).astype(np.int32)
    h = bench.lists_hash(all_d[begin:begin + count], all_i[begin:begin + count], begin) if count else 0
    hv = torch.tensor([h & 0xFFFFFFFF, h >> 32], dtype=torch.int64)
    dist.all_reduce(hv)
    lo, hi = hv.tolist()
    ok &= ((lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF) == bench.lists_hash(all_d, all_i, 0)
    results[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])   # even shards (all_gather) and ragged shards (broadcasts)
def test_world2_replication(n_total):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(2, free_port(), n_total, results), nprocs=2, join=True)
    assert results[0] and results[1]


def test_shard_ranges_tile_the_rows():
    for n in (1, 7, 64, 1000, 100001):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, w, r) for r in range(w)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for b, c in spans:
                assert c >= 0 and (c == 0 or b == pos)
                pos += c
