"""CPU checkers for the timestep path (test infrastructure only; see oracle_port.c)."""
