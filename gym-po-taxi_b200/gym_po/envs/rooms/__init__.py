from .rooms import RoomsEnv  # noqa: F401
