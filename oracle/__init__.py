"""CPU oracle for the reduced-basis hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or as the timed CPU baseline.  The product (``hippyflow_b200``) never imports this package.
"""
