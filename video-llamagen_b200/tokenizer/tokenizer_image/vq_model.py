"""``tokenizer.tokenizer_image.vq_model`` -> vlg drop-in (reference: tokenizer/tokenizer_image/vq_model.py)."""
from vlg.vq_model import (ModelArgs, VectorQuantizer, VQModel, VQ_models, VQ_8, VQ_16, Encoder, Decoder,  # noqa: F401
                          ResnetBlock, AttnBlock, Upsample, Downsample)
