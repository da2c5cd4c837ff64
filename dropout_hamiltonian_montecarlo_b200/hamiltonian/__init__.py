"""Host-side mirror of the reference's ``hamiltonian`` package for the hot path."""
