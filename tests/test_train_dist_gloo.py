"""Gradient bucket all-reduce of the training step (mqgan_b200.training.GradBucketReducer) with two gloo ranks
on the CPU: the averaged gradients equal the single-process gradient of the concatenated batch's mean loss,
including a parameter that receives no gradient (its bucket must still be reduced, not stall)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model_params():
    g = torch.Generator().manual_seed(0)
    return [torch.randn(7, 5, generator=g), torch.randn(5, generator=g), torch.randn(3, 7, generator=g),
            torch.randn(4, generator=g), torch.randn((), generator=g)]          # params[3] stays unused


def _loss(ps, x):
    h = torch.tanh(x @ ps[0].t() * ps[4] + 0.0) @ ps[2].t()
    return (h ** 2).mean() + (x @ ps[1]).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mqgan_b200.training import GradBucketReducer
    ps = [p.clone().requires_grad_(True) for p in _model_params()]
    red = GradBucketReducer(ps, bucket_bytes=64)                                 # several tiny buckets
    assert len(red.buckets) > 1
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 5, generator=g)[rank * 4:(rank + 1) * 4]
    for _ in range(2):                                                           # reuse across iterations
        red.zero()
        _loss(ps, x).backward()
        red.finish()
    if rank == 0:
        torch.save([p.grad.clone() for p in ps], out)
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    ps = [p.clone().requires_grad_(True) for p in _model_params()]
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 5, generator=g)
    (0.5 * (_loss(ps, x[:4]) + _loss(ps, x[4:]))).backward()
    for a, p in zip(got, ps):
        ref = torch.zeros_like(p) if p.grad is None else p.grad
        assert torch.allclose(a, ref, rtol=1e-5, atol=1e-7)
    assert float(got[3].abs().max()) == 0.0
