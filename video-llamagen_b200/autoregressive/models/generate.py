"""``autoregressive.models.generate`` -> vlg drop-in (reference: autoregressive/models/generate.py)."""
from vlg.generate import (generate, sample, top_k_top_p_filtering, logits_to_probs, prefill,  # noqa: F401
                          decode_one_token, decode_n_tokens)
