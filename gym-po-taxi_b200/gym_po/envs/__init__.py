from .extended_taxi import (TaxiVecEnv, HansenTaxiVecEnv, ExtendedHansenTaxiVecEnv, EXTENDED_TAXI_MAP,  # noqa: F401
                            ExtendedTaxiVecEnv, TAXI_MAP)

__all__ = ["TaxiVecEnv", "HansenTaxiVecEnv", "ExtendedHansenTaxiVecEnv", "ExtendedTaxiVecEnv", "EXTENDED_TAXI_MAP",
           "TAXI_MAP"]
