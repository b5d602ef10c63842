"""Import shim (gym_blocks/util.py imports mpi4py lazily inside mpi_fork only)."""
