"""CPU oracle for the CCDM reverse-process hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ccdm_b200`` (the product) imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  See ``ccdm_oracle.c`` for how
parity is pinned (fixtures generated from the imported reference).
"""
