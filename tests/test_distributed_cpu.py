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


# ---- SharedAccumulator: the fused flush+reduce protocol of the one-process-per-GPU launch -------
# On the GPU box the accumulators are cudaMalloc blocks of rank 0 mapped by CUDA IPC and the adds
# are red.add inside path_flush; here POSIX shared memory and a lock-protected numpy add stand in
# for them, so that the handle exchange, the double buffering and the one-barrier-per-frame
# ordering run with world_size 2 on CPU.

class _ShmBackend:
    def __init__(self):
        from multiprocessing import shared_memory
        self.shm_mod = shared_memory
        self.blocks = {}

    def alloc(self, nbytes):
        b = self.shm_mod.SharedMemory(create=True, size=nbytes)
        b.buf[:nbytes] = bytes(nbytes)
        self.blocks[id(b)] = b
        return id(b)

    def export(self, ptr):
        return self.blocks[ptr].name.encode()

    def open(self, handle):
        b = self.shm_mod.SharedMemory(name=handle.decode())
        self.blocks[id(b)] = b
        return id(b)

    def array(self, ptr, shape):
        return np.ndarray(shape, np.float32, buffer=self.blocks[ptr].buf)

    def close(self, ptr):
        self.blocks.pop(ptr).close()

    def free(self, ptr):
        b = self.blocks.pop(ptr)
        b.close()
        b.unlink()


def _shared_worker(rank, world, port, W, H, spp, frames, out_path, lock):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = _ShmBackend()
        sa = D.SharedAccumulator(W * H * 3 * 4, rank, world, be, nbuf=2)
        (rb, re, s0), cnt = D.sample_shard(spp, rank, world)
        means = []
        for k in range(frames):
            part = np.zeros((H, W, 3), np.float32)
            for s in range(s0, s0 + cnt):
                part += sample_colour(W, H, s + 1000 * k)
            with lock:  # stands in for the atomicity of red.add
                be.array(sa.ptr(k), (H, W, 3))[...] += part
            sa.barrier()
            if rank == 0:
                a = be.array(sa.ptr(k), (H, W, 3))
                means.append((a / spp).copy())
                a[...] = 0
        dist.barrier()
        sa.close()
        if rank == 0:
            np.save(out_path, np.stack(means))
    finally:
        dist.destroy_process_group()


def test_shared_accumulator_protocol_two_ranks(tmp_path):
    W, H, spp, world, frames = 16, 6, 9, 2, 5
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "means.npy")
    lock = mp.get_context("spawn").Lock()
    mp.spawn(_shared_worker, args=(world, port, W, H, spp, frames, out, lock), nprocs=world, join=True)
    got = np.load(out)
    for k in range(frames):
        full = np.zeros((H, W, 3), np.float32)
        for s in range(spp):
            full += sample_colour(W, H, s + 1000 * k)
        assert np.allclose(got[k], full / spp, rtol=1e-5, atol=1e-6), k
