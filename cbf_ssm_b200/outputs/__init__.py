from .outputs import Outputs

__all__ = ["Outputs"]
