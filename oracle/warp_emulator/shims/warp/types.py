from typing import Any, TypeVar
Array = Any; Cols = TypeVar("Cols"); Rows = TypeVar("Rows"); Scalar = TypeVar("Scalar"); Matrix = Any; Vector = Any
