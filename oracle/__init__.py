"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's HEA hot path.

Nothing under ``quanonet_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the timed CPU baseline.
"""
