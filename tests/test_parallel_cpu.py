"""Host-side logic of the data-parallel training path on CPU (gloo, world_size 2): batch sharding, replica broadcast,
flat gradient buffer layout and the per-section gradient averaging that backward triggers (train.GradSync)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vit_grid_model_b200 import MetNet3, DataParallel
        from vit_grid_model_b200.parallel import shard_batch
        from vit_grid_model_b200.train import GradBuffer
        cfg = synth.CFG_SMALL128
        torch.manual_seed(100 + rank)                       # replicas start different ...
        model = MetNet3(**cfg.metnet3_kwargs(), dropout=0.0)
        ddp = DataParallel(model)                           # ... and are equalised by the rank-0 broadcast
        digest = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum()
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        same_weights = all(torch.equal(g, gathered[0]) for g in gathered)
        assert all(k.startswith("module.") for k in ddp.state_dict())
        # gradient buffer: every parameter has an aligned view, sections tile the buffer in backward order
        G = GradBuffer(model)
        assert set(G.views) == {n for n, _ in model.named_parameters()}
        assert all(v.storage_offset() % GradBuffer.ALIGN == 0 for v in G.views.values())
        assert G.bounds[0][0] == 0 and G.bounds[-1][1] == G.flat.numel()
        assert all(a[1] == b[0] for a, b in zip(G.bounds, G.bounds[1:]))
        assert next(iter(G.views)).startswith("classifier_pm25.") and list(G.views)[-1].startswith("condition_")
        # backward's per-section hook: rank r contributes (r+1) everywhere -> mean (world+1)/2
        G.flat.fill_(float(rank + 1))
        sync = model._grad_sync
        assert sync.world == world
        for i in range(len(G.bounds)):
            sync.section_done(G.section(i))
        sync.finish()
        averaged = bool(torch.allclose(G.flat, torch.full_like(G.flat, (world + 1) / 2)))
        lo, hi = shard_batch(7, rank, world)
        q.put((rank, same_weights, averaged, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert [r[3] for r in res] == [(0, 4), (4, 7)]


def test_shard_batch_covers_everything():
    from vit_grid_model_b200.parallel import shard_batch
    for B in (1, 5, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
