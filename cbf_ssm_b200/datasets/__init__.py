from .base_ds import BaseDS
from .mat_ds import (DSManager, DSManagerDS, RoboMove, RoboMoveSimple, SpringNonlinear, create_robomove,
                     create_spring_nonlinear)
from .synthetic import (RoboMoveSynthetic, SarcosSynthetic, SpringNonlinearSynthetic,
                        VoliroShapedSynthetic)

__all__ = ["BaseDS", "DSManager", "DSManagerDS", "RoboMove", "RoboMoveSimple", "SpringNonlinear",
           "create_robomove", "create_spring_nonlinear", "RoboMoveSynthetic", "SarcosSynthetic",
           "SpringNonlinearSynthetic", "VoliroShapedSynthetic"]
