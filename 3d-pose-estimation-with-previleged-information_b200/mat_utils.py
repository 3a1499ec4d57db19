"""Drop-in for the 2-D heat-map head of the reference's ``mat_utils`` (mat_utils.py:31-55):
``to_heatmap`` = soft-max over the H*W pixels of every (sample, joint), ``decode`` = soft-argmax with
``linspace(0, 1, n) * map_range``.  Both are the D = 1 case of the volumetric head kernels (``csrc/head.cu``):
``linspace(0, 1, n) * r == linspace(0, 2, n) * (r / 2)`` exactly in fp32."""
from . import ops


def to_heatmap(ausgabe, num_joints, height, width):
    """[N, J, H, W] logits -> [N, J, H, W] fp32 soft-max heat-map (mat_utils.py:31-41)."""
    heat = ops.ToHeatmapFn.apply(ausgabe.reshape(-1, num_joints, height, width), 1, num_joints, height, width)
    return heat.reshape(-1, num_joints, height, width)


def decode(heatmap, map_range):
    """[N, J, H, W] heat-map -> [N, J, 2] = (x, y) in [0, map_range] (mat_utils.py:44-55)."""
    return ops.DecodeFn.apply(heatmap.unsqueeze(-1), float(map_range) / 2.0)[..., :2]


def heatmap_coords(ausgabe, num_joints, map_range):
    """decode(to_heatmap(.)) in one pass over the logits."""
    return ops.HeadFn.apply(ausgabe, 1, num_joints, float(map_range) / 2.0)[..., :2]
