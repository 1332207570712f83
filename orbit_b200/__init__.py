"""Import alias: the product package lives in ``orbit-dataset_b200/`` (not a valid Python
identifier), this module makes it importable as ``orbit_b200``."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "orbit-dataset_b200"))

from ._package import *  # noqa: F401,F403,E402
from ._package import __all__  # noqa: E402
