from .outputs import Outputs

from .output_summary import OutputSummary

__all__ = ["Outputs", "OutputSummary"]
