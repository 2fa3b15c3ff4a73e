"""TEST INFRASTRUCTURE ONLY — CPU oracle for the cross-modal-hash retrieval/encode path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and only as the checker / the timed CPU baseline — never as a fallback for the CUDA path.

Pinning status: the reference (kalenforn/clip-based-cross-modal-hash) ships **no tests, golden vectors
or known-answer fixtures** for this path.  The oracle is therefore pinned against the
reference *itself*, executed in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/common/calc_utils.py`` unmodified); the resulting vectors are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
