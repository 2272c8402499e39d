"""vlg -- host side of the B200-native Video-LlamaGen decode path.

Python + PyTorch own tensors, weights and the process group; every per-token kernel lives in
``lib/libvlg.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/vlg.h``).  There is no CPU or
eager fallback: importing works anywhere, but any compute call raises unless the library is built and a
B200 is present.

Drop-in modules mirroring the reference's import paths are re-exported from the sibling packages
``autoregressive`` and ``tokenizer`` (put ``video-llamagen_b200/`` on ``sys.path`` ahead of the reference
checkout).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
