from .base_ds import BaseDS
from .synthetic import (RoboMoveSynthetic, SarcosSynthetic, SpringNonlinearSynthetic,
                        VoliroShapedSynthetic)

__all__ = ["BaseDS", "RoboMoveSynthetic", "SarcosSynthetic", "SpringNonlinearSynthetic",
           "VoliroShapedSynthetic"]
