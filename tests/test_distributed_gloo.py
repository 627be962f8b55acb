"""CPU, world_size 2 over gloo: the N>1 host path -- chunk = rank sharding, weight broadcast, output gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wmar_b200.distributed import broadcast_state, gather_rows
    from wmar_b200.generate import expand_conditionings, plan_batches
    torch.manual_seed(100 + rank)
    state = {"b.weight": torch.randn(4, 3), "a.bias": torch.randn(5)}
    broadcast_state(state)
    inputs = expand_conditionings("1,9,232", 5)
    mine = plan_batches(inputs, 4, chunk_id=rank, num_chunks=world)
    rows = [[c, k] for _, b, ci in mine for c, k in zip(b, ci)]
    local = torch.tensor(rows, dtype=torch.long).reshape(-1, 2)
    parts = gather_rows(local)
    q.put((rank, {k: v.clone() for k, v in state.items()}, [p.tolist() for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_shard_broadcast_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (r0, s0, g0), (r1, s1, g1) = res
    for k in s0:
        assert torch.equal(s0[k], s1[k])          # replicas hold rank 0's weights
    assert g0 == g1                               # both ranks see the same gathered result
    allrows = sorted(tuple(r) for part in g0 for r in part)
    assert allrows == sorted((c, k + 1) for c in (1, 9, 232) for k in range(5))   # every image exactly once
    assert len(g0[0]) == 8 and len(g0[1]) == 7    # batches 0,2 -> rank 0 ; 1,3 -> rank 1 (generate.py:204)
