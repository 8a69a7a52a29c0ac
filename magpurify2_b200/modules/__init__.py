from . import composition, coverage  # noqa: F401
