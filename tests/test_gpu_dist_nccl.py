"""The slab-decomposed step over real NCCL (one process per GPU) against the single-GPU run.
Needs at least two GPUs; the single-GPU round-end box skips it (the LOCAL-transport tests in
test_gpu_dist.py cover the same protocol there)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, dims, steps):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import nprsph_b200 as sph
    from nprsph_b200.dist import SlabGroup, unique_id
    from oracle import oracle as O
    nx, ny, nz = dims
    p = O.dam_break_params(nx, ny, nz)
    p.gravity[0] = 300.0                                  # push fluid across the slab faces
    idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{rank}")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    grp = SlabGroup.nccl(rank, world, idt.cpu().numpy().tobytes(), device=rank, cell_subdiv=2,
                         rebalance_every=4)
    grp.apply_params(p)
    grp.scene_block(nx, ny, nz, 0.005, None, 2e-4, 21)
    grp.set_paused(False)
    grp.step(steps)
    rec, ids = grp.download()
    info = grp.info()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), rec=rec, ids=ids, migrated=info.migrated_total,
             ghosts=info.ghosts_left + info.ghosts_right, rebalanced=info.rebalanced)
    if rank == 0:                                          # the same scene on one GPU
        ref = sph.Simulation(device=0, cell_subdiv=2)
        ref.apply_params(p)
        ref.scene_block(nx, ny, nz, 0.005, None, 2e-4, 21)
        ref.set_paused(False)
        ref.step(steps)
        np.save(os.path.join(out_dir, "ref.npy"), ref.download())
    dist.barrier()
    grp.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_nccl_slabs_match_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from conftest import assert_field_close
    world = min(_ngpus(), 4)
    dims, steps = (96, 32, 24), 100
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), dims, steps), nprocs=world, join=True)
    n = dims[0] * dims[1] * dims[2]
    ref = np.load(tmp_path / "ref.npy")
    got = np.full((n, 16), np.nan, np.float32)
    migrated = ghosts = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.isnan(got[z["ids"], 0]).all(), "a particle is owned by two ranks"
        got[z["ids"]] = z["rec"]
        migrated += int(z["migrated"]); ghosts += int(z["ghosts"])
    assert not np.isnan(got[:, 0]).any(), "a particle is owned by no rank"
    assert migrated > 0 and ghosts > 0, (migrated, ghosts)
    for name, cols in (("pos", slice(0, 3)), ("vel", slice(4, 7)), ("force", slice(8, 11)), ("rho", 12)):
        assert_field_close(got[:, cols], ref[:, cols], name, elementwise=False)
    print(f"\n[nccl] {world} ranks, {migrated} hand-overs, result matches the single-GPU run")
