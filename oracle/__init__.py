"""CPU oracle (test infrastructure only -- see oracle/kiltro_oracle.py header)."""
