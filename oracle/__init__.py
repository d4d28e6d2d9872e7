"""CPU oracle for the DoubleMHA speaker-embedding extraction path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``doubleattentionspeakerverification_b200``)
never imports this package and raises if its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported from
``/root/reference/scripts`` in the build container by ``oracle/make_golden.py``
and committed under ``tests/golden/``.
"""
