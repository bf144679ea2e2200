from .box import Box  # noqa: F401
from .discrete import Discrete  # noqa: F401
from . import box, discrete  # noqa: F401
