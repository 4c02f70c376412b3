"""2-rank check of the path's single collective behind the C ABI: every rank folds a different mask, then
erasor_allgather_and_keep (bit-pack -> one ncclAllGather -> AND + unpack) must give numpy's AND on every rank.
Launched by tests/test_gpu_nodes.py::test_two_gpu_exchange_through_the_c_abi under torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from erasor_b200 import capi, params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # control plane only: ships the 128-byte NCCL id
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    h = capi.Handle(params.preset("seq_05"), device=local)
    h.comm_init(bytes(idt.numpy().tobytes()), world, rank)
    for n in (1, 33, 336860, 2_000_003):
        masks = [(np.random.default_rng(100 * n + r).random(n) > 0.2).astype(np.uint8) for r in range(world)]
        expect = np.minimum.reduce(masks)
        d = torch.from_numpy(masks[rank]).cuda()
        for _ in range(2):
            d.copy_(torch.from_numpy(masks[rank]))
            torch.cuda.synchronize()
            h.allgather_and_keep(d.data_ptr(), n)
            h.synchronize()
            assert np.array_equal(d.cpu().numpy(), expect), (rank, n)
    h.comm_destroy()
    h.close()
    dist.barrier()
    if rank == 0:
        print("exchange ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
