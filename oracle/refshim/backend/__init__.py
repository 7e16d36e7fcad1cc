"""NumPy implementation of the reference's `probdiffeq.backend` interface (test infrastructure, see oracle/refshim)."""
