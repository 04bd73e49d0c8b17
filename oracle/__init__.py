"""CPU oracle: test infrastructure only (see oracle/dawn_oracle.h)."""
