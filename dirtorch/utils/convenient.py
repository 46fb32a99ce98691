from dirb200.pipeline import mkdir  # noqa: F401
