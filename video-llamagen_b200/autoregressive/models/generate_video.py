"""``autoregressive.models.generate_video`` -> vlg drop-in (the reference imports this module in sample_t2v_1f.py:30 but
does not ship it)."""
from vlg.generate_video import generate  # noqa: F401
