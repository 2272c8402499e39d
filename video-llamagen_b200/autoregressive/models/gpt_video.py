"""``autoregressive.models.gpt_video`` -> vlg drop-in (reference: autoregressive/models/gpt_video.py)."""
from vlg.gpt_video import GPT_models, ModelArgs, Transformer, precompute_freqs_cis_3d_video  # noqa: F401
