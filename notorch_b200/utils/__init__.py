from .utils import UpdateMixin

__all__ = ["UpdateMixin"]
