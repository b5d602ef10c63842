from . import seeding  # noqa: F401
from .ezpickle import EzPickle  # noqa: F401
