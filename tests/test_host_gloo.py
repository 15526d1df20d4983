"""world_size-2 (and 4) CPU test of the multi-rank host logic over torch.distributed/gloo: communicator
creation from the process group (accfft_create_comm semantics), coordinate mapping, and the rank-ordered
all-gather that carries the CUDA-IPC blobs at plan time.  No GPU, no compute calls."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import accfft_b200 as acc
from oracle import accfft_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = [0, 0]
        comm = acc.accfft_create_comm(dims)          # picks rank / size from the default group
        assert comm.rank == rank and comm.nprocs == world
        assert tuple(dims) == orc.dims_create(world)
        assert comm.coords == orc.rank_to_coords(rank, dims)
        blob = bytes([rank]) * 128
        got = comm.all_gather_bytes(blob)
        assert [g[0] for g in got] == list(range(world)) and all(len(g) == 128 for g in got)
        # every rank computes the same global picture of the decomposition
        n = (12, 10, 14)
        mine = acc.accfft_local_size_dft_r2c_gpu(n, comm)
        sizes = [None] * world
        dist.all_gather_object(sizes, mine)
        total_in = sum(s[1][0] * s[1][1] * s[1][2] for s in sizes)
        total_out = sum(s[3][0] * s[3][1] * s[3][2] for s in sizes)
        assert total_in == 12 * 10 * 14 and total_out == 12 * 10 * 8
        comm.barrier()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_comm_and_blob_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
