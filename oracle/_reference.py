"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference implementation.

Used by ``tests/golden/make_golden.py`` (fixture generation) and by the
"oracle vs live reference" tests, both of which only run in the authoring
container: the reference tree is not shipped to the GPU box, so every caller
must guard with :func:`available`.

The reference is imported, never copied.  Three shims are needed before its
modules import:
  * a no-op ``ipdb`` (live ``ipdb.set_trace()`` calls sit on the inference path:
    autoregressive/models/gpt.py:350, autoregressive/models/generate.py:121,172),
  * MagicMock stand-ins for optional third-party imports pulled in by
    ``gpt_video.py`` at module level (ftfy, bs4, decord, diffusers, matplotlib),
  * a dummy ``tokenizer.tokenizer_image.vae_model`` exposing ``VAE_models``.
"""
import os
import sys
import types
import importlib
from unittest import mock

_CANDIDATES = [os.environ.get("VLG_REFERENCE"), "/root/reference"]


def root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "autoregressive", "models", "gpt.py")):
            return c
    return None


def available():
    return root() is not None


class _RefPath:
    """Context manager: put the reference root first on sys.path and hide any
    same-named drop-in packages (``autoregressive``, ``tokenizer``, ``utils``)
    that the product installs, restoring everything on exit."""

    _PKGS = ("autoregressive", "tokenizer", "utils", "language", "dataset", "diffusion")

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items()
                      if k.split(".")[0] in self._PKGS}
        for k in self.saved:
            del sys.modules[k]
        # The reference's packages have no __init__.py (namespace packages), so a regular package of the same
        # name anywhere on sys.path -- the product's drop-in shims -- would win regardless of order: hide it.
        self.hidden = [p for p in sys.path if os.path.isfile(os.path.join(p, "autoregressive", "__init__.py"))]
        for p in self.hidden:
            sys.path.remove(p)
        sys.path.insert(0, root())
        return self

    def __exit__(self, *exc):
        sys.path.remove(root())
        sys.path[0:0] = self.hidden
        self.ref_modules = {k: v for k, v in sys.modules.items()
                            if k.split(".")[0] in self._PKGS}
        for k in self.ref_modules:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


def _shim_ipdb():
    if "ipdb" not in sys.modules:
        m = types.ModuleType("ipdb")
        m.set_trace = lambda *a, **k: None
        sys.modules["ipdb"] = m


_cache = {}


def load(video=False, diff=False):
    """Return a namespace with the reference's hot-path modules.

    .gpt / .generate / .vq_model always; .gpt_video when ``video=True``; .diffloss, .gpt_video_diff and
    .generate_video_diff (the diffusion-head sampler) when ``diff=True``.
    """
    key = (bool(video), bool(diff))
    if key in _cache:
        return _cache[key]
    if not available():
        raise RuntimeError("reference tree not present (authoring container only)")
    _shim_ipdb()
    ns = types.SimpleNamespace()
    with _RefPath():
        ns.gpt = importlib.import_module("autoregressive.models.gpt")
        ns.generate = importlib.import_module("autoregressive.models.generate")
        ns.vq_model = importlib.import_module("tokenizer.tokenizer_image.vq_model")
        if video:
            for name in ("ftfy", "bs4", "decord", "diffusers", "diffusers.utils",
                         "matplotlib", "matplotlib.pyplot"):
                sys.modules.setdefault(name, mock.MagicMock())
            dummy = types.ModuleType("tokenizer.tokenizer_image.vae_model")
            dummy.VAE_models = {}
            sys.modules["tokenizer.tokenizer_image.vae_model"] = dummy
            ns.gpt_video = importlib.import_module("autoregressive.models.gpt_video")
        if diff:
            for name in ("ftfy", "bs4", "decord", "diffusers", "diffusers.utils",
                         "matplotlib", "matplotlib.pyplot"):
                sys.modules.setdefault(name, mock.MagicMock())
            if "tokenizer.tokenizer_image.vae_model" not in sys.modules:
                dummy = types.ModuleType("tokenizer.tokenizer_image.vae_model")
                dummy.VAE_models = {}
                sys.modules["tokenizer.tokenizer_image.vae_model"] = dummy
            # gpt_video_diff.py:27 does a bare ``from diffloss import DiffLoss``: its own directory must be importable
            models_dir = os.path.join(root(), "autoregressive", "models")
            sys.path.insert(0, models_dir)
            try:
                ns.diffloss = importlib.import_module("diffloss")
                ns.gpt_video_diff = importlib.import_module("autoregressive.models.gpt_video_diff")
                ns.generate_video_diff = importlib.import_module("autoregressive.models.generate_video_diff")
            finally:
                sys.path.remove(models_dir)
                sys.modules.pop("diffloss", None)
    _cache[key] = ns
    return ns
