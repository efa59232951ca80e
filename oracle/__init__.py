"""CPU oracle: a restatement of SurrealAI/surreal's actor -> replay -> learner hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``surreal_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do,
and there only as the checker / the timed CPU baseline -- never as the product path.

Each function cites the reference file:line it restates (paths relative to the reference root).
The reference is pure Python over stock ``torch``/``numpy``/``random``; so is this oracle, which
makes the two agree bit-for-bit on most fixtures.

Pinning: ``tests/test_oracle_golden.py`` checks every function here against the golden vectors
in ``tests/golden/*.npz``, which were produced by RUNNING THE REFERENCE ITSELF in the build
container (``tests/golden/make_golden.py``; third-party plumbing stubbed, arithmetic untouched).
Parity is therefore pinned, even though the reference's own tests hold no numeric fixtures
(SURVEY.md §4, §8c).
"""
from . import pd, filters, nets, gae, replay, windowing, aggregator  # noqa: F401
