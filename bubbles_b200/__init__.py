"""bubbles_b200 -- B200-native SPH / PCISPH time-step engine behind the Bubbles solver API.

The product is the CUDA library bubbles_b200/lib/libbbx.so (C ABI in include/bbx.h); this package
is the thin host-side view used by tests and bench.py.  Importing it does not need a GPU, creating
an Engine does (there is no CPU fallback).
"""
from . import _lib  # noqa: F401
from ._lib import *  # noqa: F401,F403  (enum values)
from .engine import (BbxError, ColliderSetBuilder3, Engine, MakeBox, MakeGrid, MakeMesh, MakeSDFShape, MakeSphere,  # noqa: F401
                     Translate, UtilBuildGridForDomain, identity, sdf_grid_layout)
from .slab import LocalSlabGroup, NcclSlab, plan_slabs, plan_step, plane_histogram, slab_capacity  # noqa: F401,E402
