from .base_model import BaseModel, Fetch, OutOfRangeError, Session
from .cbfssm import CBFSSM
from .cbfssmhalf import CBFSSMHALF

__all__ = ["BaseModel", "CBFSSM", "CBFSSMHALF", "Fetch", "OutOfRangeError", "Session"]
