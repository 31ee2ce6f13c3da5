"""Host-side multi-GPU logic (model3d_b200/distributed.py) on CPU: world_size-2 gloo process
group.  Each rank produces the per-pixel SUMS of its shard from a stand-in per-(pixel, sample)
colour function (the device kernels key their Philox streams the same way), the shards are
reduced to rank 0 and must equal the single-process result for both partitionings."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model3d_b200 import distributed as D


def test_split_even_properties():
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            parts = D.split_even(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        D.split_even(5, 0)
    assert D.row_band(10, 1, 4) == (3, 6, 0)
    assert D.sample_shard(10, 3, 4) == ((0, 0, 8), 2)
    assert D.ray_slice(1 << 24, 7, 8) == (7 << 21, 8 << 21)


def sample_colour(W, H, s):
    """Deterministic stand-in for one sample of every pixel, keyed by (pixel, sample)."""
    pix = np.arange(W * H, dtype=np.uint64)
    h = (pix * np.uint64(0x9E3779B97F4A7C15) + np.full(1, s, np.uint64) * np.uint64(0xD2511F53CD9E8D57)) >> np.uint64(40)
    base = (h.astype(np.float64) / float(1 << 24)).reshape(H, W, 1)
    return (base * np.array([1.0, 0.5, 0.25])).astype(np.float32)


def _worker(rank, world, port, W, H, spp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # path tracers: sample shards
        (rb, re, s0), cnt = D.sample_shard(spp, rank, world)
        acc = torch.zeros((H, W, 3), dtype=torch.float32)
        for s in range(s0, s0 + cnt):
            acc += torch.from_numpy(sample_colour(W, H, s))
        D.reduce_sums(acc, dst=0)
        # RayCaster: row bands (disjoint, zero elsewhere)
        b, e, _ = D.row_band(H, rank, world)
        band = torch.zeros((H, W, 3), dtype=torch.float32)
        band[b:e] = torch.from_numpy(sample_colour(W, H, 0))[b:e]
        D.reduce_sums(band, dst=0)
        if rank == 0:
            np.savez(out_path, mean=D.finalize_mean(acc, spp).numpy(), band=band.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_reduce_matches_single_process(tmp_path):
    W, H, spp, world = 24, 10, 13, 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, port, W, H, spp, out), nprocs=world, join=True)
    got = np.load(out)
    full = np.zeros((H, W, 3), np.float32)
    for s in range(spp):
        full += sample_colour(W, H, s)
    assert np.allclose(got["mean"], full / spp, rtol=1e-6, atol=1e-7)
    assert np.array_equal(got["band"], sample_colour(W, H, 0))
