"""CPU oracle package - TEST INFRASTRUCTURE ONLY (see oracle/pgtt_oracle.h). PARITY UNPINNED."""
