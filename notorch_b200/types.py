"""Type aliases shared with the reference (``notorch/types.py:57``)."""
from typing import Literal

Reduction = Literal["mean", "sum", "min", "max"]
