"""``tokenizer.tokenizer_image.vq_model`` -> vlg drop-in (reference: tokenizer/tokenizer_image/vq_model.py)."""
from vlg.vq_model import ModelArgs, VectorQuantizer  # noqa: F401
