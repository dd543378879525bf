"""CPU, world_size 2 over gloo: the data-parallel plumbing of the KD step — equal batch shards, one flat all-reduce
of the student gradients, and "mean of per-rank means == global mean" — checked with the CPU oracle as the model."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import glow_oracle as O
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import allreduce_mean_, glow_cfg, randomise_zero_params, shard_batch
    cfg = glow_cfg([6], 2, 1, 8, is_1d=True, y_classes=0)
    torch.manual_seed(42)                                   # same weights on every rank
    model = create_glow_model(cfg)
    randomise_zero_params(model, 1)
    sd = {k: v.clone().requires_grad_(k in dict(model.named_parameters())) for k, v in model.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    g = torch.Generator().manual_seed(7)
    xg = torch.randn(16, 6, generator=g)                    # the GLOBAL batch, identical on both ranks
    x = shard_batch(xg, rank, world)
    _, nll = O.glow_forward(sd, cfg, x)
    nll.mean().backward()
    flat = torch.zeros(sum(p.numel() for p in params))
    grads = [p.grad for p in params]
    allreduce_mean_(grads, flat)
    if rank == 0:
        # single-process reference on the whole batch
        sd1 = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in sd.items()}
        _, nll1 = O.glow_forward(sd1, cfg, xg)
        nll1.mean().backward()
        ref = [v.grad for v in sd1.values() if v.requires_grad]
        err = max((a - b).abs().max().item() / (b.abs().max().item() + 1e-12) for a, b in zip(grads, ref))
        out.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-5


def test_shard_batch_is_an_equal_partition():
    from nf_distillation_b200.train import shard_batch
    x = torch.arange(24.0).view(12, 2)
    parts = [shard_batch(x, r, 4) for r in range(4)]
    assert all(p.shape[0] == 3 for p in parts) and torch.equal(torch.cat(parts), x)
