"""Host logic of bench.py that needs no GPU: the partition-independent hash of k-lists."""
import numpy as np


def test_lists_hash_is_partition_independent_and_bit_sensitive():
    import bench
    rng = np.random.default_rng(0)
    n, k = 5000, 64
    dist = np.sort(rng.random((n, k)), axis=1)
    idx = rng.integers(0, 1_000_000, (n, k)).astype(np.int32)
    whole = bench.lists_hash(dist, idx, 0)
    for cuts in ([0, 2500, 5000], [0, 1250, 2500, 3750, 5000], [0, 17, 4000, 5000]):
        parts = sum(bench.lists_hash(dist[a:b], idx[a:b], a) for a, b in zip(cuts[:-1], cuts[1:])) & 0xFFFFFFFFFFFFFFFF
        assert parts == whole
    d2 = dist.copy(); d2[1234, 5] = np.nextafter(d2[1234, 5], 2.0)
    assert bench.lists_hash(d2, idx, 0) != whole
    i2 = idx.copy(); i2[77, 0] += 1
    assert bench.lists_hash(dist, i2, 0) != whole
    i3 = idx.copy(); i3[[10, 11]] = i3[[11, 10]]                # same rows in another order: position matters
    assert bench.lists_hash(dist, i3, 0) != whole
    assert bench.lists_hash(dist, idx, 1) != whole
