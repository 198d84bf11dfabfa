from . import container, geometry, property  # noqa: F401
