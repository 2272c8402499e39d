"""``autoregressive.models.generate_video_diff`` -> vlg drop-in (sample_t2v_1f_diff.py:28)."""
from vlg.generate_video_diff import generate  # noqa: F401
