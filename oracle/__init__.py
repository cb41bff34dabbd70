"""CPU oracle (test / baseline infrastructure only; see lexp_oracle.py header)."""
