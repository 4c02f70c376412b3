"""CPU suite, world_size 2 on gloo: the host-side logic of the multi-GPU mode (frame sharding, mask fold, the single
all-gather).  The per-frame masks come from the oracle here; on the GPUs they come from erasor_process_frames."""
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
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, vois, masks, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from erasor_b200 import dist as D
    mine = D.shard_frames(len(vois), rank, world)
    final = D.static_map_mask(n_global, [vois[f] for f in mine], [masks[f] for f in mine])
    np.save(os.path.join(out_dir, f"final_{rank}.npy"), final.numpy())
    # the streaming form bench.py uses: fold per batch, then ONE all-gather into a caller-owned buffer; repeating the
    # collective with the same buffer must give the same mask
    keep_g = D.fold_masks(n_global, [torch.from_numpy(np.asarray(vois[f]).astype(np.int64)) for f in mine],
                          [torch.from_numpy(np.asarray(masks[f]).astype(np.uint8)) for f in mine])
    buf = torch.empty((world, n_global), dtype=torch.uint8)
    again = D.allgather_and(keep_g, buf)
    again2 = D.allgather_and(keep_g, buf)
    assert torch.equal(again, final) and torch.equal(again2, final)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frames_partition():
    from erasor_b200 import dist as D
    for n in (0, 1, 7, 20, 161):
        for world in (1, 2, 3, 8):
            got = [f for r in range(world) for f in D.shard_frames(n, r, world)]
            assert got == list(range(n))
            sizes = [len(D.shard_frames(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_mask_exchange(tmp_path, oracle_mod, small_workload):
    from erasor_b200 import params as P
    p = P.preset("seq_05").replace(skip_voxelize=1)
    n_global = len(small_workload["map_world"])
    vois, masks = [], []
    o = oracle_mod.Oracle(p)
    for voi, q, k, idx in small_workload["frames"][:5]:
        r2 = voi[:, 0].astype(np.float64) ** 2 + voi[:, 1].astype(np.float64) ** 2
        sel = r2 < p.max_range ** 2
        o.run(voi[sel], q)
        _, rej = o.cloud(o.MAP_REJECTED)
        keep = np.ones(int(sel.sum()), dtype=np.uint8)
        keep[rej] = 0
        vois.append(idx[sel])
        masks.append(keep)
    expect = np.ones(n_global, dtype=np.uint8)
    for v, k in zip(vois, masks):
        np.minimum.at(expect, v, k)
    assert expect.min() == 0, "the workload must reject something"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_global, vois, masks, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(tmp_path / f"final_{r}.npy")
        assert np.array_equal(got, expect), f"rank {r}"
