"""ORACLE -- test infrastructure, NOT product code.

CPU restatement of the reference prover's hot path (scroll-tech/spartan-parallel).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package. The product package
(``spartan_parallel_b200``) never does.

Layout:
  fq.c / polys.c / sumcheck.c   C restatement (field, tables, round loops)
  cbind.py                      ctypes binding over ``_build/liboracle.so``
  pyfield.py                    independent pure-python-int field (cross-check)
  merlin.py / ristretto.py      third-party algorithms restated (merlin 3.0, RFC 9496)
  protocol.py                   transcript-level restatement (sumcheck provers,
                                sigma protocols, R1CSProof prove/verify)
"""
