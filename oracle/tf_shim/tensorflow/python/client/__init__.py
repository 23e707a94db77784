from . import device_lib  # noqa: F401
