"""TEST INFRASTRUCTURE ONLY: the CPU oracle (see oracle/api.py, oracle/gk_oracle.c)."""
