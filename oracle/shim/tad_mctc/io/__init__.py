from . import checks, read  # noqa: F401
