from .base_model import BaseModel, Fetch, OutOfRangeError, Session
from .cbfssm import CBFSSM

__all__ = ["BaseModel", "CBFSSM", "Fetch", "OutOfRangeError", "Session"]
