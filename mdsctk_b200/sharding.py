"""Row sharding of the kNN stage across the GPUs of one box (one process per GPU).

Fit rows are independent (every OpenMP iteration of the reference owns one row,
knn_rms.cpp:268-279), so the frames are split into contiguous row blocks, one per rank; every
rank needs the WHOLE reference set, which each rank packs for its own block and which is then
replicated once with NCCL (all-gather over NVLink / NVSwitch).  No reduction, no further
communication: a rank's k-lists are final and the host concatenates them in frame order.
"""


def shard_range(n_total, world, rank):
    """Contiguous block [begin, begin+count) of rank `rank`; blocks differ by at most `per` rows."""
    per = (n_total + world - 1) // world
    begin = min(rank * per, n_total)
    return begin, min(per, n_total - begin)


def replicate_frame_major(arrays, bytes_per_frame, n_total, world, rank, dist):
    """All-gathers frame-major byte arrays in place.

    arrays[i] is a flat uint8 torch tensor of n_total * bytes_per_frame[i] bytes of which this rank
    has filled its own shard_range() slice.  Uses one all_gather_into_tensor per array when the
    shards are equal, otherwise one broadcast per (array, rank).
    """
    begin, count = shard_range(n_total, world, rank)
    even = n_total % world == 0
    for t, b in zip(arrays, bytes_per_frame):
        if even:
            dist.all_gather_into_tensor(t, t[begin * b:(begin + count) * b].clone())
        else:
            for r in range(world):
                rb, rc = shard_range(n_total, world, r)
                if rc > 0:
                    dist.broadcast(t[rb * b:(rb + rc) * b], src=r)


def step_rows(begin, count, rows_per_step, step):
    """Row block a rank pushes through the sweep at `step` (bench.py): cycles over its shard."""
    rows = min(rows_per_step, count)
    nblk = max(1, count // rows)
    return begin + (step % nblk) * rows, rows
