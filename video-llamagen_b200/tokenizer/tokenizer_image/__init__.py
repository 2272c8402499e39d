import os

_ref = os.environ.get("VLG_REFERENCE")
if _ref and os.path.isdir(os.path.join(_ref, "tokenizer", "tokenizer_image")):
    __path__.append(os.path.join(_ref, "tokenizer", "tokenizer_image"))
