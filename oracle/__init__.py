"""CPU oracle for the RRDBNet(+Z) -> CEM hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain fp32 restatement (numpy /
torch-CPU functional ops) of the reference algorithm and exists so that the CUDA
path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``explorable-super-resolution_old_b200/`` (the product) imports it,
and the product raises when its CUDA library is missing instead of falling back
to this code.

Parity pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, produced in the
build container by ``oracle/gen_golden.py`` (which imports the unmodified
reference from /root/reference/codes through a few import shims) and committed as
small fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` replays
them.
"""
