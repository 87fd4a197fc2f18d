"""CPU: the N>1 host logic (contiguous shards + one-frame halo, no data-path collective) with two gloo ranks.
Each rank runs its shard through the CPU oracle chain (the stand-in for the device path on a box without a
GPU); the concatenation must equal the single-process run over the whole sequence."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_ranges_cover_the_sequence():
    from psl_slam_b200.shard import shard_range, shard_with_halo
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                got += list(range(a, b))
                s, e, halo = shard_with_halo(n, r, world)
                assert e == b and s == a - halo and halo == (1 if a > 0 and b > a else 0)
            assert got == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import orc
    from psl_slam_b200 import synth
    from psl_slam_b200.shard import drop_halo, shard_with_halo
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K = synth.ICL
    gray, depth, T = synth.sequence(5, n_frames, 320, 240, poster_size=1024)
    T12 = np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n_frames, 12))
    cam6 = np.array([K["fx"] / 2, K["fy"] / 2, K["cx"] / 2, K["cy"] / 2, K["bf"], np.float32(1) / np.float32(5000)],
                    np.float32)
    s, e, halo = shard_with_halo(n_frames, rank, world)
    p = orc.params(300, 1.2, 5, 20, 7)
    n, nm, nl, lnm = orc.frontend_batch_mt(gray[s:e], depth[s:e], T12[s:e], cam6, p, nthreads=1)
    res = drop_halo(dict(n=n, nm=nm, nl=nl, lnm=lnm), halo)
    dist.barrier()   # the only collective: timing / completion, never data
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path, orc):
    import torch.multiprocessing as mp
    from psl_slam_b200 import synth
    n_frames, world = 5, 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.start_processes(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    K = synth.ICL
    gray, depth, T = synth.sequence(5, n_frames, 320, 240, poster_size=1024)
    T12 = np.ascontiguousarray(T.astype(np.float32)[:, :3, :4].reshape(n_frames, 12))
    cam6 = np.array([K["fx"] / 2, K["fy"] / 2, K["cx"] / 2, K["cy"] / 2, K["bf"], np.float32(1) / np.float32(5000)],
                    np.float32)
    want = orc.frontend_batch_mt(gray, depth, T12, cam6, orc.params(300, 1.2, 5, 20, 7), nthreads=2)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for k, w in zip(("n", "nm", "nl", "lnm"), want):
        got = np.concatenate([p[k] for p in parts])
        assert np.array_equal(got, w), k
    assert want[1][1:].min() > 50  # frames really track across the shard seam
