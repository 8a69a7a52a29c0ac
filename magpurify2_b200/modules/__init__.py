from . import composition, coverage, filter  # noqa: F401
