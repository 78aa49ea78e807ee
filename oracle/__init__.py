"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the occlusions-4d encoder/decoder hot path.
Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / timed CPU baseline.
The product path (``occlusions-4d_b200/``) never imports this package and fails
loudly when its CUDA library is missing.
"""
