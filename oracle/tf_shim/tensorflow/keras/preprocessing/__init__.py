from . import sequence, text  # noqa: F401
