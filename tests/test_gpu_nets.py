"""GPU parity of the five networks and of the full training step (forward, head, loss, backward,
clip-norm, Adam) against fixtures produced by running the reference itself
(oracle/make_golden.py), in fp32; bf16 mode against the same fixtures at the bf16 tolerance; and
CUDA-graph replay against the eager step."""
import numpy as np
import pytest
import torch

import pose_oracle as po
from conftest import rel_err

pytestmark = pytest.mark.gpu

NET_CASES = {
    "pdepth18_s65": ("partial_depthnet", "resnet18", 65, 2, 17, {}),
    "pdepth50_s64": ("partial_depthnet", "resnet50", 64, 2, 17, {}),
    "pfusion50_s64": ("partial_fusionnet", "resnet50", 64, 2, 17, {}),
    "pfusion18_s49_j25": ("partial_fusionnet", "resnet18", 49, 2, 25, {}),
    "fusion50_s64": ("fusionnet", "resnet50", 64, 2, 17, {}),
    "fusion18_skip": ("fusionnet", "resnet18", 64, 2, 17, dict(skip_relu=True, early_dist=True)),
    "depth50_rgb_s64": ("depthnet", "resnet50", 64, 2, 19, dict(depth_only=False)),
    "depth18_d_s33": ("depthnet", "resnet18", 33, 3, 17, {}),
    "legacy50_s64": ("resnet", "resnet50", 64, 2, 19, {}),
    "pdepth50_stride8": ("partial_depthnet", "resnet50", 64, 2, 17, dict(stride=8)),
}


def build(b2pose, dev, kind, model, cfg):
    mod = getattr(b2pose, kind)
    net = getattr(mod, model)(cfg) if kind == "resnet" else getattr(mod, model)(cfg, False)
    net.load_state_dict(po.init_state(kind, model, cfg, seed=11))
    return net.to(dev)


def targs(b2pose, kind, model, cfg, **kw):
    return b2pose.train_args(model=model, num_joints=cfg.num_joints, side_in=cfg.side_in, stride=cfg.stride,
                             depth_only=cfg.depth_only, do_fusion=kind in ("fusionnet", "partial_fusionnet"), **kw)


def forward(net, kind, cfg, batch):
    if kind in ("fusionnet", "partial_fusionnet"):
        return net(batch[0], batch[1])
    if kind == "resnet":
        return net(batch[0]), None
    return net(batch[1] if (kind == "partial_depthnet" or cfg.depth_only) else batch[0])


@pytest.mark.parametrize("tag", sorted(NET_CASES))
def test_net_and_step_fp32(b2pose, dev, golden_dir, tag):
    g = np.load(golden_dir + "/nets.npz")
    kind, model, side, N, J, extra = NET_CASES[tag]
    cfg = po.net_config(side_in=side, num_joints=J, **extra)
    net = build(b2pose, dev, kind, model, cfg)
    batch = tuple(t.to(dev) for t in po.synth_batch(N, side, J, seed=3, invalid_frac=0.25))

    net.eval()
    with torch.no_grad():
        z_eval, _ = forward(net, kind, cfg, batch)
    assert tuple(z_eval.shape) == g[f"{tag}_z_eval"].shape
    assert rel_err(z_eval, g[f"{tag}_z_eval"]) < 1e-4

    net.train()
    z, last = forward(net, kind, cfg, batch)
    assert rel_err(z, g[f"{tag}_z"]) < 1e-4
    if last is not None:
        assert rel_err(last[:, :8], g[f"{tag}_last_slice"]) < 1e-4
        assert abs(float(last.mean()) - g[f"{tag}_last_mean"][0]) < 1e-4 * max(1.0, abs(g[f"{tag}_last_mean"][1]))

    # the forward above updated the BN running stats once; restart from the seed state for the steps
    net = build(b2pose, dev, kind, model, cfg)
    net.train()
    trainer = b2pose.Trainer(targs(b2pose, kind, model, cfg), net, dict(key_index=J - 1), use_graph=False)
    losses, gns = [], []
    for it in range(2):
        out = trainer.train_step(batch)
        losses.append(float(out["loss"]))
        gns.append(float(out["grad_sumsq"].sqrt()))
        if it == 0:
            spec = out["spec_cam"].cpu().numpy()
            assert np.abs(spec - g[f"{tag}_spec"]).max() < 0.1                     # mm
            true_cam, valid = batch[2].cpu().numpy(), batch[3].cpu().numpy()
            assert abs(po.mpjpe(spec, true_cam, valid) - po.mpjpe(g[f"{tag}_spec"], true_cam, valid)) < 0.1
            grads = {n: p.grad for n, p in net.named_parameters()}
            for key in g.files:
                if key.startswith(f"{tag}_gslice_"):
                    name = key[len(tag) + 8:]
                    # clip_grad_norm_ scaled the reference grads in place before they were recorded
                    coef = min(1.0, 5.0 / (g[f"{tag}_gradnorm"][0] + 1e-6))
                    want_norm = float(g[f"{tag}_gnorm_{name}"])
                    got = grads[name].detach().contiguous().reshape(-1)[:64] * coef
                    scale = max(want_norm, 1e-12)
                    assert float((got.cpu() - torch.tensor(g[key])).abs().max()) / scale < 3e-2, name
                    assert abs(float(grads[name].norm()) * coef - want_norm) / scale < 3e-2, name
    # step 1 is a pure function of the inputs; step 2 starts from Adam-updated weights, and Adam's
    # first update is ~lr*sign(g), i.e. rounding noise decides the direction for near-zero
    # gradients -- so the second step is only reproducible to a looser bound (any two BLAS differ so).
    np.testing.assert_allclose(losses[0], g[f"{tag}_loss"][0], rtol=1e-4)
    np.testing.assert_allclose(gns[0], g[f"{tag}_gradnorm"][0], rtol=5e-3)
    np.testing.assert_allclose(losses[1], g[f"{tag}_loss"][1], rtol=1e-2)
    np.testing.assert_allclose(gns[1], g[f"{tag}_gradnorm"][1], rtol=1e-1)
    sd = net.state_dict()
    np.testing.assert_allclose(sd["bn1.running_mean"].cpu().numpy(), g[f"{tag}_bn1_running_mean"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(sd["bn1.running_var"].cpu().numpy(), g[f"{tag}_bn1_running_var"], rtol=2e-3)
    # two Adam steps move every weight by <= ~2*lr = 1e-4, in a direction that is rounding noise for
    # near-zero gradients: the bulk must agree tightly, stragglers by at most the two-step travel.
    got = sd["conv1.weight"].cpu().contiguous().reshape(-1)[:64].numpy()
    diff = np.abs(got - g[f"{tag}_conv1_after"])
    assert np.median(diff) < 1e-5 and diff.max() < 2.1e-4, (np.median(diff), diff.max())
    assert int(sd["bn1.num_batches_tracked"]) == 2
    assert sum(p.numel() for p in net.parameters()) == int(g[f"{tag}_nparams"])


@pytest.mark.parametrize("tag", ["pdepth50_s64", "pfusion50_s64", "fusion50_s64", "pdepth18_s65"])
def test_net_bf16(b2pose, dev, golden_dir, tag):
    """bf16 tensor-core mode (`half_acc`): outputs within 2e-2 relative of the fp32 reference."""
    g = np.load(golden_dir + "/nets.npz")
    kind, model, side, N, J, extra = NET_CASES[tag]
    cfg = po.net_config(side_in=side, num_joints=J, **extra)
    net = build(b2pose, dev, kind, model, cfg).half()
    batch = tuple(t.to(dev) for t in po.synth_batch(N, side, J, seed=3, invalid_frac=0.25))
    net.eval()
    with torch.no_grad():
        z_eval, _ = forward(net, kind, cfg, batch)
    assert z_eval.dtype == torch.bfloat16
    assert rel_err(z_eval, g[f"{tag}_z_eval"]) < 2e-2
    # Training step: the fp32 fixture is not a meaningful reference for bf16 training-mode numbers (training-mode
    # BatchNorm amplifies rounding ~200x through ResNet-50, tests/golden/bf16_sensitivity.npz), so the step is compared
    # with the oracle under the same bf16 storage contract, on bf16-rounded filters and inputs; the per-layer bounds
    # (2e-2 forward, 3e-2 backward) are in test_gpu_bf16_blocks.py, the full-size step in test_gpu_bf16_step.py.
    sd = po.init_state(kind, model, cfg, seed=11)
    sd = {k: (v.bfloat16().float() if v.dim() == 4 else v) for k, v in sd.items()}
    hb = po.synth_batch(N, side, J, seed=3, invalid_frac=0.25)
    hb = (hb[0].bfloat16().float(), hb[1].bfloat16().float(), hb[2], hb[3])
    orc = po.StepOracle({k: v.clone() for k, v in sd.items()}, kind, model, cfg, key_index=J - 1, act_round=po.round_bf16)
    loss_o, gn_o, spec_o, _ = orc.step(hb)
    net = getattr(getattr(b2pose, kind), model)(cfg, False)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    trainer = b2pose.Trainer(targs(b2pose, kind, model, cfg, half_acc=True), net, dict(key_index=J - 1),
                             use_graph=False)
    out = trainer.train_step(tuple(t.to(dev) for t in hb))
    spec = out["spec_cam"].cpu().numpy()
    true_cam, valid = hb[2].numpy(), hb[3].numpy()
    dm = abs(po.mpjpe(spec, true_cam, valid) - po.mpjpe(spec_o.numpy(), true_cam, valid))
    print(tag, "bf16 loss", float(out["loss"]), "contract oracle", loss_o, "|dMPJPE| mm", dm)
    # (batch 2: 32 values per channel in layer4 -- on this fixture the ORACLE's own loss moves by 0.6 % between two CPUs
    #  through the summation order of its BLAS; the well-conditioned batch-8 fixture of test_gpu_bf16_step.py holds 2e-2.
    #  Measured over the round's kernel versions: loss within 0.01-1.7 %, |dMPJPE| <= 13 mm.  This is a plausibility check of
    #  the whole step on a chaotic fixture; the parity gates proper are the per-block tests named above)
    assert abs(float(out["loss"]) - loss_o) / loss_o < 5e-2
    assert dm < 40.0


def test_graph_replay_matches_eager(b2pose, dev):
    """The CUDA-graph step (captured after 3 eager warm-ups) continues the same trajectory as eager."""
    kind, model = "partial_fusionnet", "resnet18"
    cfg = po.net_config(side_in=64, num_joints=17)
    batch = tuple(t.to(dev) for t in po.synth_batch(2, 64, 17, seed=3))
    traj = {}
    for use_graph in (False, True):
        net = build(b2pose, dev, kind, model, cfg)
        net.train()
        tr = b2pose.Trainer(targs(b2pose, kind, model, cfg), net, dict(key_index=16), use_graph=use_graph)
        traj[use_graph] = [float(tr.train_step(batch)["loss"]) for _ in range(6)]
        traj[(use_graph, "w")] = net.state_dict()["layer3.0.conv1.weight"].clone()
        traj[(use_graph, "nbt")] = int(net.state_dict()["bn1.num_batches_tracked"])
    np.testing.assert_allclose(traj[True][:4], traj[False][:4], rtol=1e-3)      # 3 eager warm-ups + first replay
    np.testing.assert_allclose(traj[True], traj[False], rtol=5e-3)              # fp32 wgrad atomics reorder sums
    assert rel_err(traj[(True, "w")], traj[(False, "w")]) < 1e-3
    assert traj[(True, "nbt")] == traj[(False, "nbt")] == 6
    assert traj[False][-1] < traj[False][0]            # the loss goes down on a fixed batch


def test_step_vs_oracle_other_seed(b2pose, dev):
    """Independent of the fixtures: a fresh seed / shape checked against the CPU oracle's step."""
    kind, model = "partial_depthnet", "resnet18"
    cfg = po.net_config(side_in=97, num_joints=19)
    sd = po.init_state(kind, model, cfg, seed=21)
    batch = po.synth_batch(3, 97, 19, seed=8, invalid_frac=0.5)
    orc = po.StepOracle({k: v.clone() for k, v in sd.items()}, kind, model, cfg, key_index=18, criterion="L1")
    loss_o, gn_o, spec_o, z_o = orc.step(batch)
    net = getattr(b2pose, kind).resnet18(cfg, False)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    tr = b2pose.Trainer(b2pose.train_args(model=model, num_joints=19, side_in=97, criterion="L1"), net,
                        dict(key_index=18), use_graph=False)
    out = tr.train_step(tuple(t.to(dev) for t in batch))
    assert abs(float(out["loss"]) - loss_o) / loss_o < 1e-3
    assert abs(float(out["grad_sumsq"].sqrt()) - gn_o) / gn_o < 1e-2
    assert float((out["spec_cam"].cpu() - spec_o).abs().max()) < 0.1


def test_state_dict_and_api(b2pose, dev):
    cfg = po.net_config(side_in=64, num_joints=17)
    net = b2pose.partial_fusionnet.resnet50(cfg, False)
    sd = net.state_dict()
    assert list(sd.keys()) == list(po.param_shapes("partial_fusionnet", "resnet50", cfg).keys())
    for k, shp in po.param_shapes("partial_fusionnet", "resnet50", cfg).items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert isinstance(net.conv2, torch.nn.Conv2d) and isinstance(net.conv2, b2pose.PartialConv)
    assert isinstance(net.layer5[0].conv2, b2pose.PartialConv) and not isinstance(net.layer1[0].conv2, b2pose.PartialConv)
    with pytest.raises(TypeError):
        net.to(dev)(torch.randn(1, 3, 64, 64, device=dev))
    with pytest.raises(RuntimeError):                       # CPU input: no fallback
        b2pose.partial_depthnet.resnet18(cfg, False)(torch.randn(1, 1, 64, 64))
    # outputs are logical NCHW
    z, feat = net(torch.randn(2, 3, 64, 64, device=dev), torch.rand(2, 1, 64, 64, device=dev))
    assert tuple(z.shape) == (2, 272, 4, 4) and tuple(feat.shape) == (2, 2048, 4, 4)
    # round trip through a state dict saved from the flat-buffer trainer
    tr = b2pose.Trainer(b2pose.train_args(num_joints=17, side_in=64, do_fusion=True), net, dict(key_index=16),
                        use_graph=False)
    net2 = b2pose.partial_fusionnet.resnet50(cfg, False)
    net2.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    assert torch.equal(net2.layer6[0].conv2.weight, net.layer6[0].conv2.weight.cpu())


def test_epoch_loop_prefetch(b2pose, dev):
    """Trainer.train(epoch, loader) with pinned host batches (one-batch lookahead H2D on a side
    stream) follows the same trajectory as feeding device batches step by step."""
    kind, model = "partial_depthnet", "resnet18"
    cfg = po.net_config(side_in=64, num_joints=17)
    batches = [po.synth_batch(2, 64, 17, seed=10 + i) for i in range(4)]
    out = {}
    for mode in ("steps", "epoch"):
        net = build(b2pose, dev, kind, model, cfg)
        net.train()
        tr = b2pose.Trainer(targs(b2pose, kind, model, cfg), net, dict(key_index=16), use_graph=False)
        if mode == "steps":
            tr.adapt_learn_rate(1)
            losses = [float(tr.train_step(tuple(t.to(dev) for t in b))["loss"]) for b in batches]
            out[mode] = sum(l * 2 for l in losses) / 8
        else:
            pinned = [tuple(t.pin_memory() for t in b) for b in batches]
            out[mode] = tr.train(1, pinned)["cam_train_loss"]
            assert tr.lr == pytest.approx(5e-5 * 0.2)            # warm-up epoch (depth_train.py:621-625)
    assert out["epoch"] == pytest.approx(out["steps"], rel=2e-3)


def test_frozen_batchnorm_and_eval(b2pose, dev):
    """depthnet.freeze_batchnorm() (depthnet.py:158-161): BN layers run on their running statistics
    while the rest of the net trains; gradients follow the frozen-statistics formula."""
    kind, model = "depthnet", "resnet18"
    cfg = po.net_config(side_in=64, num_joints=17)
    sd = po.init_state(kind, model, cfg, seed=31)
    g = torch.Generator().manual_seed(5)
    for k in sd:                                     # non-trivial running statistics
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    batch = po.synth_batch(2, 64, 17, seed=6)
    ref = {k: v.clone() for k, v in sd.items()}
    for k in po.trainable_names(ref):
        ref[k].requires_grad_(True)
    z, _ = po.net_forward(ref, kind, model, cfg, batch[1], None, training=False)
    loss, spec = po.pose_loss(z, batch[2], batch[3], depth=16, num_joints=17, side_out=4, depth_range=1000.0,
                              key_index=16)
    loss.backward()
    net = b2pose.depthnet.resnet18(cfg, False)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    net.freeze_batchnorm()
    zg, _ = net(batch[1].to(dev))
    coords = b2pose.heatmap_coords(zg, 16, 17, 1000.0)
    lg, sg = b2pose.pose_loss(coords, batch[2].to(dev), batch[3].to(dev), 16)
    lg.backward()
    assert abs(float(lg) - float(loss)) / float(loss) < 1e-4
    assert float((sg.cpu() - spec.detach()).abs().max()) < 0.1
    for name in ("conv1.weight", "layer2.0.conv1.weight", "layer3.1.bn2.weight", "layer4.0.downsample.1.bias",
                 "regressor.weight"):
        got = dict(net.named_parameters())[name].grad
        assert rel_err(got, ref[name].grad) < 2e-3, name
    assert torch.equal(net.bn1.running_mean.cpu(), sd["bn1.running_mean"])      # untouched
    assert int(net.bn1.num_batches_tracked) == 0
    # Trainer.predict: eval-mode forward + head
    tr = b2pose.Trainer(targs(b2pose, kind, model, cfg), net, dict(key_index=16), use_graph=False)
    spec_eval, _ = tr.predict(batch)
    assert float((spec_eval.cpu() - spec.detach()).abs().max()) < 0.1
