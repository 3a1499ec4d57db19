"""nn.Module building blocks on the libb2pose kernels.

``PartialConv`` keeps the reference's operator interface (partial_conv.py:6-58):
``PartialConv(*Conv2d args, multi_channel=False, return_mask=True)(input, mask_in) ->
(output, updated_mask)`` on NCHW tensors, and is an ``nn.Conv2d`` subclass so the reference's
init loops (`isinstance(m, nn.Conv2d)`, partial_depthnet.py:187-189) and state_dict layout are
unchanged.  Filters are held in channels_last memory (= KRSC, what the kernels read); logical
shapes stay ``[K, C, R, S]``.
"""
import torch
from torch import nn

from . import ops


def _square(v, what):
    if isinstance(v, (tuple, list)):
        if len(v) != 2 or v[0] != v[1]:
            raise NotImplementedError("b2pose convolutions need equal %s in both dims, got %r" % (what, v))
        return int(v[0])
    return int(v)


class _B2ConvBase(nn.Conv2d):
    """Shared plumbing: geometry checks, KRSC parameter memory, bf16 shadow filter."""

    force_ffma = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.groups != 1 or self.padding_mode != "zeros" or isinstance(self.padding, str):
            raise NotImplementedError("b2pose convolutions support groups=1, zero padding only")
        self._s = _square(self.stride, "stride")
        self._p = _square(self.padding, "padding")
        self._d = _square(self.dilation, "dilation")
        self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)
        self._shadow = None           # bf16 KRSC copy maintained by the Trainer's fused Adam step
        self._grad_sink = None        # [K,C,R,S] view of the Trainer's flat gradient buffer
        self._bias_sink = None        # [K] view of the same buffer (layers with a bias)

    def shadow(self, dtype):
        s = self._shadow
        return s if (s is not None and s.dtype == dtype and dtype != self.weight.dtype) else None

    def _sinks(self):
        """(dw sink, db sink) inside a Trainer-managed model with gradients enabled, else None."""
        if self._grad_sink is None or not torch.is_grad_enabled() or (self.bias is not None and self._bias_sink is None):
            return None
        return (self._grad_sink, self._bias_sink)

    def _conv_cfg(self, partial, premasked=False):
        return (self._s, self._p, self._d, partial, premasked, self.force_ffma)

    @staticmethod
    def _to_nhwc(t):
        return t.permute(0, 2, 3, 1).contiguous()


class Conv2d(_B2ConvBase):
    """Plain convolution (nn.Conv2d.forward replacement), NCHW in / NCHW (channels_last memory) out."""

    def forward(self, input):
        assert len(input.shape) == 4
        y = self.forward_nhwc(self._to_nhwc(input))
        return y.permute(0, 3, 1, 2)

    def forward_nhwc(self, x):
        y, _ = ops.ConvFn.apply(x, None, self.weight, self.bias, self.shadow(x.dtype), self._conv_cfg(False),
                                self._sinks())
        return y


class PartialConv(_B2ConvBase):
    """Mask-renormalised convolution (partial_conv.py:6-58)."""

    def __init__(self, *args, **kwargs):
        self.multi_channel = kwargs.pop("multi_channel", False)
        self.return_mask = kwargs.pop("return_mask", True)
        super().__init__(*args, **kwargs)
        if self.multi_channel:
            raise NotImplementedError("multi_channel=True is not used by any reference network and is not built")
        self.slide_winsize = self.kernel_size[0] * self.kernel_size[1]      # partial_conv.py:28

    def forward(self, input, mask_in):
        assert len(input.shape) == 4                                          # partial_conv.py:33
        if mask_in.shape[1] != 1 or mask_in.shape[0] != input.shape[0] or mask_in.shape[2:] != input.shape[2:]:
            raise ValueError("mask_in must be [N, 1, H, W] matching the input")
        x = self._to_nhwc(input)
        m = mask_in.detach().reshape(mask_in.shape[0], mask_in.shape[2], mask_in.shape[3])
        y, mo = self.forward_nhwc(x, m.float())
        out = y.permute(0, 3, 1, 2)
        if self.bias is not None and input.dtype != torch.float32:
            out = out.float()          # the reference's bias path multiplies by the fp32 mask (:51)
        if self.return_mask:
            return out, mo.unsqueeze(1).to(mask_in.dtype)
        return out

    def forward_nhwc(self, x, mask, premasked=False):
        return ops.ConvFn.apply(x, mask, self.weight, self.bias, self.shadow(x.dtype),
                                self._conv_cfg(True, premasked), self._sinks())


PartialConv2d = PartialConv     # the name BASELINE.json uses


class BatchNorm2d(nn.BatchNorm2d):
    """Parameter container for the fused conv+BN node; `defer_count` lets the Trainer bump
    num_batches_tracked for all layers with one foreach op outside the captured graph."""

    defer_count = False
    _grad_sinks = None      # (gamma.grad, beta.grad) views of the Trainer's flat gradient buffer

    def tick(self):
        if self.training and self.track_running_stats and not self.defer_count:
            self.num_batches_tracked += 1

    def forward(self, input):
        raise NotImplementedError("b2pose BatchNorm2d is evaluated fused with its convolution (conv_bn)")


def conv_bn(x, veil, conv, bn, relu, residual=None, mask_output=False, premasked=False, dx_holder=None,
            res_holder=None, park_holder=None, x2=None):
    """conv -> bn (+residual) (+relu) (*veil) on NHWC tensors; returns (z, veil_out).
    x2: convolve the channel concatenation [x, x2] without materialising it (ops.ConvBNFn)."""
    partial = isinstance(conv, PartialConv)
    training = bn.training
    if bn.momentum is None:
        # nn.BatchNorm2d(momentum=None) is a cumulative moving average (factor 1/num_batches_tracked); no reference
        # net uses it and the fused kernels take one constant factor
        raise NotImplementedError("b2pose BatchNorm2d needs a numeric momentum (cumulative averaging is not built)")
    cfg = (conv._s, conv._p, conv._d, partial, premasked and partial, relu, mask_output and partial, training,
           bn.momentum, bn.eps, conv.force_ffma)
    sinks = None
    if conv._grad_sink is not None and bn._grad_sinks is not None and torch.is_grad_enabled():
        sinks = (conv._grad_sink,) + bn._grad_sinks
    z, vout = ops.ConvBNFn.apply(x, veil if partial else None, conv.weight, conv.shadow(x.dtype), bn.weight,
                                 bn.bias, bn.running_mean, bn.running_var, residual, cfg, sinks, dx_holder, res_holder,
                                 park_holder, x2)
    bn.tick()
    return z, (vout if partial else veil)
