"""Import-path shim (see autoregressive/__init__.py)."""
import os

_ref = os.environ.get("VLG_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "tokenizer")):
    __path__.append(os.path.join(_ref, "tokenizer"))
