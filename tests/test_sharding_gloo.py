"""Host-side multi-rank logic on CPU (gloo, world_size 2): each rank owns a contiguous shard of the
block range, no data-path collective, timing reduced with MAX, results identical to the single-rank
run.  The per-rank work here is the CPU oracle port (the CUDA path needs a GPU); what is under test is
the sharding / reduction logic bench.py uses."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_port, ptr, probagen

BLOCK, SLOT = 32768, 33548


def _compress(data):
    port = load_port()
    nb = (len(data) + BLOCK - 1) // BLOCK
    cbuf = np.zeros(nb * SLOT, np.uint8); cs = np.zeros(nb, np.uint64)
    port.orc_compress_blocks(1, ptr(data), len(data), BLOCK, ptr(cbuf), SLOT, ptr(cs), 255, 12)
    return cbuf, cs


def _worker(rank, world, port_no, total_blocks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = total_blocks // world
    # rank r owns bytes [r*S, (r+1)*S) of the generator stream (bench.py: FSEB200_probagen(offset = rank*n))
    stream = probagen(total_blocks * BLOCK, 0.14)
    shard = np.ascontiguousarray(stream[rank * per * BLOCK:(rank + 1) * per * BLOCK])
    _, cs = _compress(shard)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)       # stand-in for a per-rank time
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sizes = [torch.zeros(per, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.from_numpy(cs.astype(np.int64)))
    if rank == 0:
        q.put((float(t[0]), torch.cat(sizes).numpy()))
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    world, total_blocks = 2, 8
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, total_blocks, q)) for r in range(world)]
    [p.start() for p in procs]
    tmax, sizes = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert tmax == float(world)                                   # MAX over ranks
    _, want = _compress(probagen(total_blocks * BLOCK, 0.14))
    assert np.array_equal(sizes.astype(np.uint64), want)          # shards concatenate to the single-rank result
