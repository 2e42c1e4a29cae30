"""CPU oracle (test infrastructure only; see seekr_oracle.py and skr_oracle.c)."""
