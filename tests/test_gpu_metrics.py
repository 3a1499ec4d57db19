"""GPU parity of the evaluation metrics: utils.analyze / statistics / parse_epoch with the
back-rotation of the test loops, against fixtures produced by the reference's utils module; and the
Trainer's test loop against the oracle."""
import numpy as np
import pytest
import torch

import pose_oracle as po

pytestmark = pytest.mark.gpu
KEYS = ("solid", "close", "depth", "jitter", "switch", "fail", "score_pck", "score_auc", "cam_mean")


def test_analyze_and_epoch(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/metrics.npz")
    thresh = dict(zip(("solid", "close", "rough"), g["thresh"]))
    acc = b2pose.MetricAccumulator(g["mirror"], thresh, dev)
    per_batch = []
    for b in range(int(g["n_batches"])):
        args = [torch.tensor(g[f"b{b}_{k}"], device=dev) for k in ("spec", "true", "valid")]
        rot = torch.tensor(g[f"b{b}_rot"], device=dev)
        s = b2pose.analyze(*args, g["mirror"], thresh, back_rotate=rot)
        assert s["batch_size"] == int(g[f"b{b}_batch_size"])
        for k in KEYS:
            np.testing.assert_allclose(s[k], g[f"b{b}_{k}"], rtol=2e-6, atol=1e-9), (b, k)
        per_batch.append(s)
        acc.update(*args, rot)
    for ep in (acc.result(), b2pose.parse_epoch(per_batch)):
        for k in KEYS:
            np.testing.assert_allclose(ep[k], g[f"epoch_{k}"], rtol=2e-6, atol=1e-9)
    # without rotation / mirror the distances are unchanged; only the tangent (depth) class may move
    s0 = b2pose.analyze(*[torch.tensor(g[f"b0_{k}"], device=dev) for k in ("spec", "true", "valid")], None, thresh)
    np.testing.assert_allclose(s0["cam_mean"], g["b0_cam_mean"], rtol=2e-6)
    np.testing.assert_allclose(s0["score_pck"], g["b0_score_pck"], rtol=1e-12)


def test_trainer_test_loop(b2pose, dev):
    kind, model, side, J = "partial_depthnet", "resnet18", 64, 17
    cfg = po.net_config(side_in=side, num_joints=J)
    sd = po.init_state(kind, model, cfg, seed=11)
    net = getattr(b2pose, kind).resnet18(cfg, False)
    net.load_state_dict(sd)
    net = net.to(dev)
    thresh = dict(solid=50.0, close=100.0, rough=150.0)
    mirror = np.array([0, 2, 1, 4, 3, 5, 7, 6, 8, 10, 9, 12, 11, 14, 13, 16, 15])
    args = b2pose.train_args(model=model, num_joints=J, side_in=side, stride=16, thresh=thresh)
    tr = b2pose.Trainer(args, net, dict(key_index=J - 1, mirror=mirror), use_graph=False)
    rng = np.random.RandomState(0)
    loader, want_stats, want_loss, total = [], [], 0.0, 0
    for b, N in enumerate((3, 2)):
        color, depth, true_cam, true_val = po.synth_batch(N, side, J, seed=20 + b)
        q, _ = np.linalg.qr(rng.randn(N, 3, 3))
        rot = torch.tensor(q.astype(np.float32))
        loader.append((color, depth, true_cam, true_val, rot))
        with torch.no_grad():
            z, _ = po.net_forward(sd, kind, model, cfg, depth, None, training=False)
            loss, spec = po.pose_loss(z, true_cam, true_val, depth=cfg.depth, num_joints=J, side_out=(side - 1) // 16 + 1,
                                      depth_range=1000.0, key_index=J - 1, loss_div=10.0, criterion="SmoothL1")
        want_stats.append(po.analyze(spec.numpy(), true_cam.numpy(), true_val.numpy(), mirror, thresh, rot.numpy()))
        want_loss += float(loss) * N
        total += N
    rec = tr.test(1, loader)
    assert not net.training
    want = po.parse_epoch(want_stats)
    np.testing.assert_allclose(rec["test_loss"], want_loss / total, rtol=1e-4)
    np.testing.assert_allclose(rec["cam_mean"], want["cam_mean"], rtol=1e-4)
    for k in KEYS[:8]:
        assert abs(rec[k] - want[k]) <= 1.0 / 40 + 1e-9, k        # at most one joint changes class at a threshold
    with pytest.raises(RuntimeError, match="thresh"):
        tr.thresh = None
        tr.test(1, loader)
