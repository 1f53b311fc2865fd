"""CPU oracle for the Fresnel differentiable Gaussian-splatting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fresnel_b200/`` may import this
package: it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
is the CUDA library and fails loudly without it.

Parity status: PINNED against outputs of the reference module itself
(``/root/reference/scripts/models/differentiable_renderer.py``) executed in the
build container by ``oracle/make_golden.py``; the resulting vectors live in
``tests/golden/`` and ``tests/test_oracle_golden.py`` re-checks the oracle
against them on every CPU run.  The reference ships no golden vectors or
known-answer tests of its own for this path.
"""
