"""CPU oracle for the seq2squiggle predict hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``seq2squiggle_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the CPU arm, never as the thing shipped.

Parity pin: ``oracle/make_golden.py`` runs the *unmodified* reference modules
(``/root/reference/src/seq2squiggle/{layers,modules}.py``) in the build
container and stores their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
bit for bit (fp32, matmul precision "highest").
"""
