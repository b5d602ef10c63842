"""Import shim for the un-vendored `baselines` package (TEST INFRASTRUCTURE)."""
from . import logger  # noqa: F401
