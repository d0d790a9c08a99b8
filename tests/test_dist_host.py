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


def _rebalance_worker(rank, world, port, out_dir):
    """Each rank knows its own counter block, exchanges it with the neighbour over gloo (what the
    per-step ncclSend/ncclRecv of 12 words does) and evaluates the face rule: both must agree."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nprsph_b200 import dist as D
    rng = np.random.default_rng(7)
    decisions = []
    for trial in range(50):
        layer = int(rng.integers(100, 5000))
        blocks = np.zeros((world, D.SLAB_COUNTER_WORDS), np.uint32)
        for r in range(world):
            own = int(rng.integers(20, 200)) * layer
            blocks[r, [D.CNT_HALO_L, D.CNT_HALO_R]] = 2 * layer
            blocks[r, D.CNT_OWN], blocks[r, D.CNT_FREE] = own, int(rng.integers(0, 400)) * layer
            blocks[r, D.CNT_WIDTH], blocks[r, D.CNT_CAP_MIGRATE] = own // layer, int(rng.integers(1, 8)) * layer
        mine = torch.from_numpy(blocks[rank].astype(np.int64))
        other = torch.zeros_like(mine)
        peer = 1 - rank
        ops = [dist.P2POp(dist.isend, mine, peer), dist.P2POp(dist.irecv, other, peer)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        theirs = other.numpy().astype(np.uint32)
        assert np.array_equal(theirs, blocks[peer])
        a, b = (blocks[rank], theirs) if rank == 0 else (theirs, blocks[rank])
        decisions.append(D.slab_face_move(a, b, 2, 64 * layer))
    np.save(os.path.join(out_dir, f"moves{rank}.npy"), np.array(decisions))
    dist.destroy_process_group()


def test_rebalancing_rule_agrees_across_ranks(tmp_path):
    mp.spawn(_rebalance_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    m0, m1 = np.load(tmp_path / "moves0.npy"), np.load(tmp_path / "moves1.npy")
    assert np.array_equal(m0, m1), "both ranks of a face must take the same decision"
    assert set(m0.tolist()) == {-1, 0, 1}, "the random trials should exercise all three outcomes"


def test_rebalancing_rule_properties(sph):
    from nprsph_b200 import dist as D
    def block(own, layer, free=10**7, width=40, cap_mig=10**6):
        b = np.zeros(D.SLAB_COUNTER_WORDS, np.uint32)
        b[[D.CNT_HALO_L, D.CNT_HALO_R]] = 2 * layer
        b[D.CNT_OWN], b[D.CNT_FREE], b[D.CNT_WIDTH], b[D.CNT_CAP_MIGRATE] = own, free, width, cap_mig
        return b
    L = 1000
    assert D.slab_face_move(block(50 * L, L), block(50 * L, L), 2, 10**6) == 0          # balanced
    assert D.slab_face_move(block(50 * L, L), block(51 * L, L), 2, 10**6) == 0          # within two layers
    assert D.slab_face_move(block(60 * L, L), block(40 * L, L), 2, 10**6) == -1         # left rank is heavier
    assert D.slab_face_move(block(40 * L, L), block(60 * L, L), 2, 10**6) == +1
    assert D.slab_face_move(block(60 * L, L, width=5), block(40 * L, L), 2, 10**6) == 0  # giver too narrow (2R+2 = 6)
    assert D.slab_face_move(block(60 * L, L), block(40 * L, L, free=1000), 2, 10**6) == 0   # receiver full
    assert D.slab_face_move(block(60 * L, L, cap_mig=1500), block(40 * L, L), 2, 10**6) == 0  # layer exceeds the migration buffer
    assert D.slab_face_move(block(60 * L, L), block(40 * L, L), 2, 3 * L) == 0          # ... or the ghost capacity
    assert D.slab_face_move(block(60 * L, L), block(40 * L, L), 1, 10**6) == 0          # reach 1: no room in the boundary layer


def test_rebalancing_by_measured_time(sph):
    """rebalance_every < 0: equal particle counts but unequal density-pass times (the disordered front
    of a dam break) still move the face, towards the rank that takes longer."""
    from nprsph_b200 import dist as D
    def block(own, layer, cost_us):
        b = np.zeros(D.SLAB_COUNTER_WORDS, np.uint32)
        b[[D.CNT_HALO_L, D.CNT_HALO_R]] = 2 * layer
        b[D.CNT_OWN], b[D.CNT_FREE], b[D.CNT_WIDTH], b[D.CNT_CAP_MIGRATE], b[D.CNT_COST_US] = own, 10**7, 40, 10**6, cost_us
        return b
    L = 65536
    slow, fast = block(256 * L, L, 1950), block(256 * L, L, 1690)
    assert D.slab_face_move(slow, fast, 2, 10**7) == 0                         # by count: balanced
    assert D.slab_face_move(slow, fast, 2, 10**7, by_time=True) == -1          # the slow rank hands a layer over
    assert D.slab_face_move(fast, slow, 2, 10**7, by_time=True) == +1
    assert D.slab_face_move(block(256 * L, L, 1700), fast, 2, 10**7, by_time=True) == 0   # within 3 layers' worth
    assert D.slab_face_move(block(256 * L, L, 0), fast, 2, 10**7, by_time=True) == 0      # no measurement yet: by count
