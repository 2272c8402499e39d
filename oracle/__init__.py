"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the reference hot path.

Nothing under ``oracle/`` is product code.  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and only as the checker / the timed CPU baseline --
never by ``video-llamagen_b200/`` (the product fails loudly when its CUDA
library is missing instead of falling back to anything here).

Parity pinning: the reference (BinZhu-ece/Video-LlamaGen) ships no tests, golden
vectors or fixtures for this path, so the restatement
is pinned against *outputs of the reference itself*, produced in the authoring
container by importing the unmodified reference from /root/reference
(``oracle/_reference.py``) and committed under ``tests/golden/`` by
``tests/golden/make_golden.py``.  ``tests/test_oracle_golden.py`` re-checks the
restatement against those fixtures on every run, and against the live reference
whenever /root/reference is present.
"""
