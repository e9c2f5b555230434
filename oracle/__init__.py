"""CPU oracle (test infrastructure only; see oracle/oracle.cc header)."""
