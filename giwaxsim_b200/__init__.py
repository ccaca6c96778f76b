"""giwaxsim_b200 -- B200 (sm_100a) implementation of GIWAXSim's reciprocal-space
hot path behind the reference's own Python interface.

    from giwaxsim_b200.tools.comparison import voxelgridmaker_fitting, detectormaker_fitting

The CUDA library (libgiwaxs_b200.so, built in-tree by `python -m
giwaxsim_b200.build`) is loaded lazily on first use; there is no CPU fallback.
"""
__version__ = "0.1.0"
