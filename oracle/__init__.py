"""CPU oracle for the DeformConv2d hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package ``jittor_dcn_b200``.  Allowed
importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.
"""
