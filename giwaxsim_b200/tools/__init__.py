"""Reference-facing modules: same module and function names as the reference's
`tools` package for the hot path (voxelgrids, detector, comparison, utilities)."""
from . import utilities, voxelgrids, detector, comparison  # noqa: F401
