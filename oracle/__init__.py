"""CPU oracle for the AFGSA hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``pixel_heal_thyself_b200/`` may import this package: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs are allowed to use it, and there only as the checker
or the reported CPU baseline, never as the product path.
"""
