from .crooms import CRoomsEnv  # noqa: F401
from .rooms import RoomsEnv  # noqa: F401
from .msrooms import MultistoryFourRoomsEnv  # noqa: F401
