"""``autoregressive.models.gpt_video_diff`` -> vlg drop-in (sample_t2v_1f_diff.py:29)."""
from vlg.gpt_video_diff import GPT_models, ModelArgs, Transformer  # noqa: F401
