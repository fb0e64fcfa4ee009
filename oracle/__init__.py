"""TEST INFRASTRUCTURE: the CPU oracle (see oracle.py).  Not importable from product code."""
