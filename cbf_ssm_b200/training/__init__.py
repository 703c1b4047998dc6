from .trainer import Trainer

__all__ = ["Trainer"]
from .tf_checkpoint import (export_reference_checkpoint, import_reference_checkpoint, read_tf_checkpoint,  # noqa: F401
                            reference_variable_names, write_tf_checkpoint)
