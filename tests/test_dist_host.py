"""Host-side logic of the multi-GPU path on CPU, world_size 2 over gloo: every rank must derive
the same count-balanced slab boundaries from an all-reduced x-plane histogram, and the ownership
rule (global x cell in [x_begin, x_end)) must give every particle exactly one owner."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nprsph_b200.dist import slab_partition
    from oracle import oracle as O
    nx, ny, nz = 48, 10, 8
    p = O.dam_break_params(nx, ny, nz)
    g = O.grid_setup(p, 2)
    # each rank looks at half of the particles, the histogram is all-reduced
    P = O.jitter(O.make_block(nx, ny, nz), 2e-4, 5)
    mine = np.ascontiguousarray(P[rank::world])
    keys = O.cell_keys(mine, g).astype(np.int64)
    cx = keys // (g.dim[1] * g.dim[2])
    hist = torch.from_numpy(np.bincount(cx, minlength=g.dim[0]).astype(np.int64))
    dist.all_reduce(hist)
    bounds = slab_partition(hist.numpy().astype(np.uint64), world, 2 * g.reach)
    # every rank must hold identical boundaries
    gathered = [torch.zeros(world + 1, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(bounds.copy()))
    assert all(torch.equal(gathered[0], t) for t in gathered)
    # ownership: exactly one owner per particle, counts balanced
    all_cx = O.cell_keys(P, g).astype(np.int64) // (g.dim[1] * g.dim[2])
    owner = np.searchsorted(bounds[1:], all_cx, side="right")
    counts = np.bincount(owner, minlength=world)
    assert counts.sum() == len(P) and (owner < world).all()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([bounds, counts]))
    dist.destroy_process_group()


def test_slab_partition_agrees_across_ranks(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"r{r}.npy") for r in range(world))
    assert np.array_equal(r0, r1)
    bounds, counts = r0[:world + 1], r0[world + 1:]
    assert bounds[0] == 0 and (np.diff(bounds) >= 2).all()
    assert abs(int(counts[0]) - int(counts[1])) <= 0.15 * counts.sum(), counts


def test_slab_partition_properties(sph):
    from nprsph_b200.dist import slab_partition
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for _ in range(20):
            dimx = int(rng.integers(world * 4, 400))
            hist = rng.integers(0, 1000, dimx).astype(np.uint64)
            hist[rng.integers(0, dimx):] = 0 if rng.random() < 0.3 else hist[-1]   # empty tail, like a dam break
            b = slab_partition(hist, world, 4)
            assert b[0] == 0 and b[-1] == dimx and (np.diff(b) >= 4).all()
    with pytest.raises(ValueError):
        slab_partition(np.ones(10, np.uint64), 4, 4)      # 4 slabs x 4 cells do not fit in 10
    # balanced when the histogram allows it
    b = slab_partition(np.full(64, 100, np.uint64), 4, 2)
    assert list(b) == [0, 16, 32, 48, 64]
