"""Drop-in for the reference's ``partial_conv`` module (partial_conv.py:6): ``PartialConv``."""
from .layers import PartialConv, PartialConv2d  # noqa: F401
