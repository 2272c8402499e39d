"""``autoregressive.models.gpt`` -> vlg drop-in (reference: autoregressive/models/gpt.py)."""
from vlg.gpt import *  # noqa: F401,F403
from vlg.gpt import GPT_models, ModelArgs, Transformer, RMSNorm, precompute_freqs_cis_2d, find_multiple  # noqa: F401
