"""Drop-in replacements for the reference's model/*.py classes (same ctor / forward signatures and
state_dict keys), executing on the sm_100a kernels in libpvcr_b200.so."""
from .S2VTAttModel import S2VTAttModel  # noqa: F401
from .S2VTModel import S2VTModel  # noqa: F401
from .RationaleNet import Generator, RationaleNet  # noqa: F401
from .SpatialNet import SpatialNet  # noqa: F401
