"""TEST INFRASTRUCTURE.  CPU (numpy) restatement of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package -- as the checker or the timed CPU
baseline, never as part of the product path (``pyqmc_b200`` does not import it).
"""
