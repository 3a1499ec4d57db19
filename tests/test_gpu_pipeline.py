"""GPU parity of the on-device input pipeline: homography crop with cv2.remap
semantics fused with ToTensor + Normalize (colour) and with to_depth + enhance_ntu / enhance_pku (depth),
against fixtures produced by the reference's cameralib / depth_datasets / utils code."""
import numpy as np
import pytest
import torch

import pose_oracle as po

pytestmark = pytest.mark.gpu


def _close(got, want, atol, rtol, outlier_frac, outlier_tol):
    """All but a tiny fraction of the pixels agree tightly; the rest (source coordinates that land within one
    float ulp of a 1/64-pixel rounding boundary, where numpy's sgemm and the kernel's multiply-adds may round
    apart) by at most one interpolation step."""
    err = np.abs(got - want)
    bad = err > (atol + rtol * np.abs(want))
    assert bad.mean() <= outlier_frac, bad.mean()
    assert err.max() <= outlier_tol, err.max()


def test_crop_normalize_rgb(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/pipeline.npz")
    P = b2pose.pipeline
    for name in g["names"]:
        hom = P.homography((g[f"{name}_K_old"], g[f"{name}_R_old"]), (g[f"{name}_K_new"], g[f"{name}_R_new"]))
        assert np.array_equal(hom, g[f"{name}_hom"])
        frames = torch.tensor(g[f"{name}_color"], device=dev)[None]
        side = g[f"{name}_color_crop"].shape[0]
        out = P.crop_normalize_rgb(frames, hom[None], side)
        assert tuple(out.shape) == (1, 3, side, side) and out.dtype == torch.float32
        # one uint8 step is 1/255/std = 0.0175
        _close(out[0].cpu().numpy(), g[f"{name}_color_out"], 1e-6, 1e-6, 2e-3, 0.04)
    # a batch: two frames of one size with their own homographies
    frames = torch.tensor(np.stack([g["centre_color"], g["corner_flip_color"]]), device=dev)
    out = P.crop_normalize_rgb(frames, np.stack([g["centre_hom"], g["corner_flip_hom"]]), 48)
    _close(out[1].cpu().numpy(), g["corner_flip_color_out"], 1e-6, 1e-6, 2e-3, 0.04)
    with pytest.raises(TypeError):
        P.crop_normalize_rgb(frames.float(), g["centre_hom"][None], 48)
    with pytest.raises(RuntimeError, match="CUDA"):
        P.crop_normalize_rgb(frames.cpu(), g["centre_hom"][None], 48)


def test_crop_enhance_depth(b2pose, dev, golden_dir):
    g = np.load(golden_dir + "/pipeline.npz")
    P = b2pose.pipeline
    for name in g["names"]:
        hom, side = g[f"{name}_hom"], g[f"{name}_depth_crop"].shape[0]
        frames = torch.tensor(g[f"{name}_depth"], device=dev)[None]
        crop = P.crop_enhance_depth(frames, hom[None], side, enhance=False)
        _close(crop[0, 0].cpu().numpy(), g[f"{name}_depth_crop"], 1e-9, 1e-6, 2e-3, 0.01)
        for key, kw in (("ntu_exp", dict(data_name="ntu", nexponent=True)), ("ntu_lin", dict(data_name="ntu", nexponent=False)),
                        ("pku_exp", dict(data_name="pku", nexponent=True)),
                        ("todepth_ntu_exp", dict(data_name="ntu", nexponent=True, to_depth_intrinsics=g[f"{name}_K"]))):
            out = P.crop_enhance_depth(frames, hom[None], side, **kw)
            assert tuple(out.shape) == (1, 1, side, side)
            _close(out[0].cpu().numpy(), g[f"{name}_{key}"], 1e-7, 2e-6, 3e-3, 1.0)
        # the stand-alone enhance_* on an already cropped image: exact crop in -> tight everywhere
        d = torch.tensor(g[f"{name}_depth_crop"], device=dev)
        np.testing.assert_allclose(P.enhance_ntu(d, True).cpu().numpy(), g[f"{name}_ntu_exp"], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(P.enhance_ntu(d, False).cpu().numpy(), g[f"{name}_ntu_lin"], rtol=1e-6)
        np.testing.assert_allclose(P.enhance_pku(d, True).cpu().numpy(), g[f"{name}_pku_exp"], rtol=2e-6, atol=1e-9)
        # invalid pixels stay exactly zero: they become the veil of the partial convolutions
        assert np.array_equal(P.enhance_ntu(d, True).cpu().numpy() == 0, g[f"{name}_ntu_exp"] == 0)


def test_pipeline_feeds_the_network(b2pose, dev):
    """Full-size property run: 64 frames 424x512 -> 256x256 crops -> enhance -> partial_depthnet veil; an
    identity homography reproduces the top-left window exactly, invalid pixels are exactly 0."""
    N, Hs, Ws, S = 64, 424, 512, 256
    g = torch.Generator(device="cpu").manual_seed(0)
    depth = (torch.rand(N, Hs, Ws, generator=g) * 0.08 + 0.002) * (torch.rand(N, Hs, Ws, generator=g) > 0.25)
    depth = depth.to(dev)
    eye = np.repeat(np.eye(3, dtype=np.float32)[None], N, 0)
    out = b2pose.pipeline.crop_enhance_depth(depth, eye, S, "ntu", True)
    want = po.enhance(depth[:2, :S, :S].cpu().numpy().reshape(-1, S), True, "ntu").reshape(2, S, S)
    np.testing.assert_allclose(out[:2, 0].cpu().numpy(), want, rtol=2e-6, atol=1e-9)
    veil = b2pose.ops.veil_from_depth(out.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(veil.bool(), depth[:, :S, :S] / (10.0 / 255.0) >= 0.1)
    frames = torch.randint(0, 256, (8, 270, 480, 3), dtype=torch.uint8, generator=g).to(dev)
    rgb = b2pose.pipeline.crop_normalize_rgb(frames, eye[:8], S)
    want = po.normalize_rgb(frames[3, :S, :S].cpu().numpy())
    np.testing.assert_allclose(rgb[3].cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-6)
