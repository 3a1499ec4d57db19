"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/b2pose.h
declares; the ctypes table matches the header; host-side logic (network tables, schedules,
sharding, bucketing) agrees with the oracle / reference semantics.  No kernels are launched."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import pose_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b2pose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header(b2pose):
    syms = header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(b2pose._lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libb2pose.so does not export %s" % s
    assert sorted(b2pose._lib.SIGNATURES) == syms          # the binding covers exactly the header
    assert b2pose._lib.lib().b2_abi_version() == b2pose._lib.ABI_VERSION == 5


def test_header_arg_counts_match_binding(b2pose):
    src = open(os.path.join(ROOT, "include", "b2pose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in b2pose._lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_flag_constants_match_header(b2pose):
    """The binding's convolution flags are the header's enum values (a drifted flag would silently select another path)."""
    src = open(os.path.join(ROOT, "include", "b2pose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    header = {k: int(v) for k, v in re.findall(r"\bB2_(CONV_[A-Z_0-9]+)\s*=\s*(\d+)", src)}
    assert len(header) >= 9 and header["CONV_X_CONCAT"] == 256
    for name, value in header.items():
        assert getattr(b2pose._lib, name) == value, name
    assert int(re.search(r"#define\s+B2_ABI_VERSION\s+(\d+)", src).group(1)) == b2pose._lib.ABI_VERSION


def test_tensor_pair_is_two_device_pointers(b2pose):
    """B2_CONV_X_CONCAT operands travel as a host array of two pointers (here: CPU tensors, no launch)."""
    a, b = torch.zeros(2, 3, 4, 8), torch.ones(2, 3, 4, 8)
    pair = b2pose._lib.TensorPair(a, b)
    arr = b2pose._lib.ptr(pair)
    assert len(arr) == 2 and arr[0] == a.data_ptr() and arr[1] == b.data_ptr()
    with pytest.raises(AssertionError):
        b2pose._lib.TensorPair(a, torch.zeros(2, 3, 4, 16))


def test_argument_validation_without_gpu(b2pose):
    """Bad descriptors are rejected on the host before any launch."""
    L = b2pose._lib
    d = L.ConvDesc(1, 8, 8, 4, 4, 3, 3, 1, 1, 1, 7, 8, L.F32, 0)          # wrong Ho
    rc = L.lib().b2_pconv_fprop(ctypes.byref(d), 1, None, 1, None, 1, None, None, None, None, 0, None)
    assert rc == -1 and b"output size" in L.lib().b2_last_error()
    rc = L.lib().b2_head_fwd(None, 1, 1, 1, 1, 1, 0, 0, 1.0, None, None, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        L.call("b2_pose_loss", None, None, None, 1, 1, 0, 1.0, 0, None, None, None, None)


def test_no_cpu_fallback(b2pose):
    conv = b2pose.PartialConv(4, 4, kernel_size=3, padding=1, bias=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(torch.randn(1, 4, 5, 5), torch.ones(1, 1, 5, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        b2pose.heatmap_coords(torch.zeros(1, 32, 4, 4), 16, 2, 1000.0)
    net = b2pose.depthnet.resnet18(po.net_config(side_in=64), False)
    with pytest.raises(RuntimeError):
        b2pose.Trainer(b2pose.train_args(side_in=64, num_joints=17), net, dict(key_index=16))


@pytest.mark.parametrize("kind", ["depthnet", "partial_depthnet", "fusionnet", "partial_fusionnet", "resnet"])
@pytest.mark.parametrize("model", ["resnet18", "resnet50"])
def test_state_dict_layout_matches_reference(b2pose, kind, model):
    cfg = po.net_config(side_in=257, num_joints=19, joint_space=(kind == "resnet"))
    mod = getattr(b2pose, kind)
    net = getattr(mod, model)(cfg) if kind == "resnet" else getattr(mod, model)(cfg, False)
    want = po.param_shapes(kind, model, cfg)
    sd = net.state_dict()
    assert list(sd.keys()) == list(want.keys())
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    convs = [m for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
    assert all(m.weight.permute(0, 2, 3, 1).is_contiguous() for m in convs)      # KRSC memory
    n_partial = sum(isinstance(m, b2pose.PartialConv) for m in convs)
    if kind == "partial_depthnet" and model == "resnet50":
        assert n_partial == 22 and len(convs) == 54
    if not kind.startswith("partial"):
        assert n_partial == 0


def test_golden_key_order(b2pose, golden_dir):
    g = np.load(os.path.join(golden_dir, "shapes.npz"))
    cfg = po.net_config(side_in=256, num_joints=17)
    for kind in ("partial_depthnet", "partial_fusionnet", "fusionnet"):
        net = getattr(b2pose, kind).resnet50(cfg, False)
        assert list(net.state_dict().keys()) == [str(k) for k in g[f"{kind}_keys"]]
        assert sum(p.numel() for p in net.parameters()) == int(g[f"{kind}_nparams"])


def test_stage_strides_and_asserts(b2pose):
    for s in (4, 8, 16, 32):
        assert b2pose.nets.stage_strides(s) == po.stage_strides(s)
    assert b2pose.nets.stage_strides(16) == ((1, 2, 2, 1), (1, 1, 1, 2))
    with pytest.raises(AssertionError):
        b2pose.resnet.resnet18(po.net_config(stride=8))                 # resnet.py:126
    with pytest.raises(AssertionError):
        b2pose.partial_depthnet.resnet18(po.net_config(depth_only=False), False)   # partial_depthnet.py:164


def test_learn_rate_schedule(b2pose):
    class Fake:      # adapt_learn_rate only touches these fields
        pass
    tr = Fake()
    tr.warmup, tr.learn_rate, tr.learn_decay, tr.warmup_factor = 1, 5e-5, 0.2, 0.2
    for epoch in (1, 2, 15, 16, 20, 21, 25, 26, 40):
        got = b2pose.Trainer.adapt_learn_rate(tr, epoch)
        assert got == pytest.approx(po.learn_rate_at(epoch)), epoch


def test_pretrain_surgery(b2pose, tmp_path):
    """ImageNet-checkpoint surgery of the builders (fusionnet.py:243-297, partial_depthnet.py:232-257)."""
    cfg = po.net_config(side_in=64, num_joints=17)
    donor = po.init_state("depthnet", "resnet18", po.net_config(depth_only=False), seed=4)
    donor = {k: v for k, v in donor.items() if not k.startswith("regressor")}
    donor["fc.weight"] = torch.zeros(10, 512)
    path = str(tmp_path / "imagenet.pth")
    torch.save(donor, path)
    cfg.model_path, cfg.depth_host, cfg.host_path = path, False, None
    net = b2pose.partial_fusionnet.resnet18(cfg, True)
    sd = net.state_dict()
    assert torch.equal(sd["conv1.weight"], donor["conv1.weight"])
    assert torch.equal(sd["conv2.weight"], donor["conv1.weight"][:, :1])
    assert torch.equal(sd["layer5.1.conv2.weight"], donor["layer1.1.conv2.weight"])
    assert torch.equal(sd["layer6.0.downsample.0.weight"], donor["layer2.0.downsample.0.weight"])
    assert torch.equal(sd["bn2.running_var"], donor["bn1.running_var"])
    net = b2pose.partial_depthnet.resnet18(cfg, True)
    assert torch.equal(net.state_dict()["conv1.weight"], donor["conv1.weight"][:, :1])


def test_synthetic_batch_contract(b2pose):
    color, depth, cam, val = b2pose.synthetic_batch(3, 64, 17, None, seed=2, invalid_frac=0.25)
    assert color.shape == (3, 3, 64, 64) and depth.shape == (3, 1, 64, 64)
    assert cam.shape == (3, 17, 3) and val.shape == (3, 17) and val.dtype == torch.bool
    assert bool(val[:, 16].all())
    frac = float((depth == 0).float().mean())
    assert 0.2 < frac < 0.6 and float(depth[depth != 0].min()) >= 0.05


def test_bucket_and_shard_helpers(b2pose):
    P = b2pose.parallel if hasattr(b2pose, "parallel") else __import__("b2pose.parallel", fromlist=["x"])
    b = P.bucket_bounds(1000, 256)
    assert b[0] == (0, 256) and b[-1][1] == 1000 and all(lo % 64 == 0 for lo, _ in b)
    assert sum(hi - lo for lo, hi in b) == 1000
    assert P.bucket_bounds(0, 256) == []
    got = [P.shard_range(10, r, 4) for r in range(4)]
    assert got == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) prints ONE JSON line with the
    contract's keys; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-batch", "1", "--side", "64", "--model", "resnet18"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "partial_fusionnet.py"))
    # the reference's own modules when baseline/_ref is staged (baseline/stage_reference.py), else the oracle port
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_host_side_helpers_against_golden(b2pose, golden_dir):
    """Pure host logic of the widened surface: the float32 homography of the crop (cameralib.py:672-674) and
    parse_epoch (utils.py:224-231) reproduce the fixtures produced by the reference, no GPU involved."""
    import numpy as np
    g = np.load(os.path.join(golden_dir, "pipeline.npz"))
    for name in g["names"]:
        hom = b2pose.pipeline.homography((g[f"{name}_K_old"], g[f"{name}_R_old"]), (g[f"{name}_K_new"], g[f"{name}_R_new"]))
        assert hom.dtype == np.float32 and np.array_equal(hom, g[f"{name}_hom"])
    m = np.load(os.path.join(golden_dir, "metrics.npz"))
    keys = ("solid", "close", "depth", "jitter", "switch", "fail", "score_pck", "score_auc", "cam_mean", "batch_size")
    stats = [{k: float(m[f"b{b}_{k}"]) for k in keys} for b in range(int(m["n_batches"]))]
    ep = b2pose.parse_epoch(stats)
    for k in keys[:-1]:
        np.testing.assert_allclose(ep[k], float(m[f"epoch_{k}"]), rtol=1e-12)
    with pytest.raises(RuntimeError, match="CUDA"):
        b2pose.pipeline.crop_enhance_depth(torch.zeros(1, 8, 8), None, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        b2pose.mimic_loss(torch.zeros(1, 4, 2, 2), torch.zeros(1, 4, 2, 2), torch.ones(1, 1, 2, 2))
