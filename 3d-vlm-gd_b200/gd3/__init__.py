"""gd3 -- B200-native geometric-distillation hot path (host-side mirror of the reference's Python API).

Layers
  gd3._lib      ctypes binding of the C ABI in include/gd3.h (lib3dgd.so, hand-written sm_100a CUDA)
  gd3.ops       torch.autograd.Function wrappers: the fused, batched entry points
  gd3.compat    modules with the reference's exact function signatures
                (utils.losses / utils.functions / mast3r.fast_nn)
There is no CPU or PyTorch fallback: every compute entry point raises if lib3dgd.so is missing or
no CUDA device is present.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
__version__ = '0.1.0'
