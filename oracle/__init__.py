"""CPU oracles for the Hamming-matching hot path.  TEST INFRASTRUCTURE ONLY (see hamming_oracle.py)."""
