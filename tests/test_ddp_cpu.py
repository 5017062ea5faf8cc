"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: clip sharding + the single flat
gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sam2_video_training_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, rdv, use_bucket, q, steps=1):
    # file rendezvous: no TCP-store port to lose in a race between picking a free port and binding it
    os.environ["GLOO_SOCKET_IFNAME"] = "lo"
    dist.init_process_group("gloo", init_method="file://" + rdv, rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    if use_bucket:
        bucket = ddp.attach_grad_bucket(model)
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    clips = ddp.shard_clips(7, rank, world)
    x = torch.stack([torch.full((8,), float(c + 1)) for c in clips])
    for step in range(steps):
        if step > 0:
            # the PyTorch default detaches every .grad from the bucket; autograd then writes FRESH gradient tensors and
            # the all-reduce must pick those up (and re-install the views), not reduce the stale flat buffer
            model.zero_grad(set_to_none=True)
        loss = model(x).pow(2).sum() / 7.0 * world  # so that the mean over ranks == full-batch gradient
        loss.backward()
        if use_bucket and step == 0:
            h = ddp.allreduce_gradients_async(model, world)     # overlappable form; wait() orders it before the optimizer
            assert h is not None
            h.wait()
        else:
            ddp.allreduce_gradients(model, world)
        if use_bucket:
            assert bucket.owns_all()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    q.put((rank, clips, flat))
    dist.barrier()                      # nobody tears its pairs down while the peer is still in the collective
    dist.destroy_process_group()


class _Transient(Exception):
    pass


def _run_world2(use_bucket, tmp_path, attempt, steps=1):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    rdv = str(tmp_path / f"rdv_{attempt}")
    procs = [ctx.Process(target=_worker, args=(r, 2, rdv, use_bucket, q, steps)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    try:
        res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    except queue.Empty as e:            # a worker died before reporting (loop-back connection reset while the pairs connect)
        raise _Transient(str(e))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return res


@pytest.mark.parametrize("use_bucket,steps", [(True, 1), (False, 1), (True, 2)])
def test_grad_allreduce_world2(use_bucket, steps, tmp_path):
    """steps = 2: zero_grad(set_to_none=True) between the steps (the stale bucket must not be reduced)."""
    try:
        res = _run_world2(use_bucket, tmp_path, 0, steps)
    except _Transient:                  # ONLY a worker that never reported is retried; assertion failures are not
        res = _run_world2(use_bucket, tmp_path, 1, steps)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert torch.allclose(res[0][2], res[1][2])
    # single-process reference over all 7 clips
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    x = torch.stack([torch.full((8,), float(c + 1)) for c in range(7)])
    (model(x).pow(2).sum() / 7.0).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(res[0][2], ref, rtol=1e-5, atol=1e-6)


def test_shard_clips_partition():
    for world in (1, 2, 4, 8):
        allc = sorted(c for r in range(world) for c in ddp.shard_clips(13, r, world))
        assert allc == list(range(13))
