"""CPU oracles for the dense optical-flow hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``opencv_contrib_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.  The product path is the CUDA
library (``libb200flow.so``) and fails loudly when that is missing.
"""
