"""TEST INFRASTRUCTURE -- not part of the product.

CPU restatements of the reference's algorithms (and, under ``_ref/``, the reference's own CUDA kernels compiled from where they lie) used as
the checker of the parity tests.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` /
``ref_*`` legs may import anything from here; ``transoar_b200/`` never does (``tests/test_host_logic.py`` checks that)."""
