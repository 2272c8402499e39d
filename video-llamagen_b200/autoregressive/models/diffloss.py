"""``autoregressive.models.diffloss`` -> vlg drop-in (gpt_video_diff.py:27 imports it as a bare ``diffloss`` module)."""
from vlg.diffloss import DiffLoss, SimpleMLPAdaLN  # noqa: F401
