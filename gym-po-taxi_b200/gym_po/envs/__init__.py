from .extended_taxi import (TaxiVecEnv, HansenTaxiVecEnv, ExtendedHansenTaxiVecEnv, EXTENDED_TAXI_MAP,  # noqa: F401
                            ExtendedTaxiVecEnv, TAXI_MAP)
from .rooms import RoomsEnv  # noqa: F401

__all__ = ["RoomsEnv", "TaxiVecEnv", "HansenTaxiVecEnv", "ExtendedHansenTaxiVecEnv", "ExtendedTaxiVecEnv", "EXTENDED_TAXI_MAP",
           "TAXI_MAP"]
