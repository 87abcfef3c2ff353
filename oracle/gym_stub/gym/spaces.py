from . import _Box as Box, _MultiDiscrete as MultiDiscrete  # noqa: F401
