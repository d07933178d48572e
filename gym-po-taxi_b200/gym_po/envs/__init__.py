from .extended_taxi import (TaxiVecEnv, HansenTaxiVecEnv, ExtendedHansenTaxiVecEnv, EXTENDED_TAXI_MAP,  # noqa: F401
                            ExtendedTaxiVecEnv, TAXI_MAP)
from .rooms import RoomsEnv, CRoomsEnv, MultistoryFourRoomsEnv  # noqa: F401
from .tag import TagVecEnv  # noqa: F401
from .car_flag import CarVecEnv, DiscreteActionCarVecEnv  # noqa: F401

__all__ = ["RoomsEnv", "CRoomsEnv", "MultistoryFourRoomsEnv", "TagVecEnv", "CarVecEnv", "DiscreteActionCarVecEnv", "TaxiVecEnv", "HansenTaxiVecEnv", "ExtendedHansenTaxiVecEnv", "ExtendedTaxiVecEnv", "EXTENDED_TAXI_MAP",
           "TAXI_MAP"]
