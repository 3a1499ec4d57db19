"""Two-rank NCCL parity of the data-parallel training step (needs >= 2 visible GPUs; skipped otherwise).

Replaces nn.DataParallel (depth_main.py:72): one process per GPU, per-rank BatchNorm statistics, gradients summed over
ranks (the mean is folded into the fused Adam kernel).  Each rank steps on its own shard; the test checks that

  * the exchanged gradient equals the sum of the two ranks' own gradients (each rank first steps a local,
    single-process Trainer from the same weights on its shard) -- to fp32 rounding in fp32 mode, to bf16 rounding with the
    compressed exchange of the bf16 mode (run with the deterministic BatchNorm reductions, B2POSE_BN_TOTALS=0: on this
    batch-2 fixture the order of the fp32 atomics alone moves the bf16 gradients by 10-20 %),
  * both ranks hold identical weights after the step (eager and CUDA-graph mode, plain and two-stage overlapped
    exchange),
  * (single GPU) the two-stage backward pass of the overlapped mode -- the autograd graph cut at the input of layer3 --
    yields the gradients of the plain backward pass, eagerly and replayed from its two CUDA graphs.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import __graft_entry__ as ge
    import pose_oracle as po
    return ge.load_package(), po


def _make(b2, po, dev, half):
    cfg = po.net_config(side_in=64, num_joints=17, depth_only=False)
    net = b2.partial_fusionnet.resnet18(cfg, False)
    net.load_state_dict(po.init_state("partial_fusionnet", "resnet18", cfg, seed=5))
    args = b2.train_args(model="resnet18", num_joints=17, side_in=64, depth_only=False, do_fusion=True, half_acc=half)
    return net.to(dev).train(), args


def _worker(rank, world, port, out, half, overlap, use_graph, steps):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), B2POSE_DDP_OVERLAP="1" if overlap else "0", B2POSE_BN_TOTALS="0")
    b2, po = _setup()
    from b2pose import parallel as P
    import torch.distributed as dist
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    # this rank's own gradient first: a local Trainer (no process group yet) on the same shard
    net, args = _make(b2, po, dev, half)
    local = b2.Trainer(args, net, dict(key_index=16), use_graph=False)
    local.train_step(tuple(t.to(dev) for t in po.synth_batch(2, 64, 17, seed=20 + rank)))
    g_local = local.flat.g.cpu()
    del local, net
    P.init_from_env("nccl")
    net, args = _make(b2, po, dev, half)
    tr = b2.Trainer(args, net, dict(key_index=16), use_graph=use_graph)
    assert tr.world == 2 and tr.overlap == overlap
    batch = tuple(t.to(dev) for t in po.synth_batch(2, 64, 17, seed=20 + rank))
    losses = [float(tr.train_step(batch)["loss"]) for _ in range(steps)]
    torch.cuda.synchronize()
    torch.save(dict(g=tr.flat.g.cpu(), w=tr.flat.w.cpu(), losses=losses, g_local=g_local), out % rank)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("half,overlap,use_graph", [(False, False, False), (True, False, False), (True, True, False),
                                                    (True, False, True), (True, True, True)])
def test_two_rank_step(tmp_path, half, overlap, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    steps = 5 if use_graph else 1          # graph mode: 3 eager warm-ups, capture, 1 replay
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, half, overlap, use_graph, steps), nprocs=2, join=True)
    a, b = torch.load(out % 0), torch.load(out % 1)
    assert torch.equal(a["w"], b["w"])                      # identical replicas after the step(s)
    assert torch.equal(a["g"], b["g"])                      # both hold the same reduced gradient
    assert all(l == l for l in a["losses"] + b["losses"])
    if steps > 1:
        return
    # the reduced gradient is the sum of the ranks' own gradients
    total = a["g_local"] + b["g_local"]
    err = float((a["g"] - total).norm() / total.norm())
    print("half", half, "overlap", overlap, "reduced-gradient L2 error vs sum of per-rank gradients: %.2e" % err)
    assert err < (1e-2 if half else 1e-4)


def _one_trainer(b2, dev, kind, two, use_graph, steps):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pose_oracle as po
    fused = kind == "partial_fusionnet"
    cfg = po.net_config(side_in=64, num_joints=17, depth_only=not fused)
    net = getattr(b2, kind).resnet18(cfg, False)
    net.load_state_dict(po.init_state(kind, "resnet18", cfg, seed=5))
    # learning rate 0: the weights stay put, so every step (eager warm-up or graph replay) computes the same gradient
    args = b2.train_args(model="resnet18", num_joints=17, side_in=64, depth_only=not fused, do_fusion=fused, half_acc=False,
                         learn_rate=0.0, weight_decay=0.0, warmup=0)
    tr = b2.Trainer(args, net.to(dev).train(), dict(key_index=16), use_graph=use_graph)
    tr._force_two = two
    batch = tuple(t.to(dev) for t in po.synth_batch(4, 64, 17, seed=20))
    for _ in range(steps):
        tr.train_step(batch)
    torch.cuda.synchronize()
    return tr.flat.g.clone()


@pytest.mark.parametrize("kind", ["partial_fusionnet", "partial_depthnet"])
def test_staged_backward_matches_plain(b2pose, dev, kind):
    """fp32 (deterministic up to the order of the weight-gradient atomics): every parameter's gradient, including the
    node that PRODUCES the boundary tensor, must come out of the staged pass once -- not zero, not twice."""
    plain = _one_trainer(b2pose, dev, kind, False, False, 1)
    staged = _one_trainer(b2pose, dev, kind, True, False, 1)
    err = float((staged - plain).norm() / plain.norm())
    print(kind, "staged vs plain backward, eager, step 1: %.2e" % err)
    assert err < 1e-5
    staged5 = _one_trainer(b2pose, dev, kind, True, True, 5)          # 3 eager warm-ups, capture, 1 replay of fb + fb2
    err = float((staged5 - plain).norm() / plain.norm())
    print(kind, "staged (replayed from its two CUDA graphs) vs plain eager: %.2e" % err)
    assert err < 1e-5
