"""CPU oracle for the Faster R-CNN region path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this package.
The product (``minddet_b200``) never does.  The arithmetic lives in ``region_oracle.c`` (strict
fp32, compiled with ``-ffp-contract=off``); this module is the ctypes/numpy binding.
Conventions: ``oracle/CONVENTIONS.md``.  PARITY STATUS per function: header of region_oracle.c.
"""
from .cpu import *  # noqa: F401,F403
