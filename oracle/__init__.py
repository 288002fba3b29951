"""CPU oracle for the NoisyNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``noisynet_b200``)
may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs, and there only as the checker / the CPU arm being timed.

Parity status: the reference (michaelklachko/NoisyNet) ships no tests and no
golden vectors, so parity is "unpinned" by the reference's own tests.  The
oracle is pinned instead against outputs of the unmodified reference modules
executed in the build container (``oracle/gen_golden.py`` -> ``tests/golden``).
"""
