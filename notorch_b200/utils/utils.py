"""``UpdateMixin`` — same contract as the reference's ``notorch/utils/utils.py:34-40``: ``update``
returns a *shallow copy* with some attributes replaced, so index tensors (and anything cached on the
instance, e.g. the int32 CSR bundle) are shared between the input and output graphs."""
from __future__ import annotations

import copy as _copy
from typing import Self


class UpdateMixin:
    def update(self, in_place: bool = False, **kwargs) -> Self:
        target = self if in_place else _copy.copy(self)
        for name, value in kwargs.items():
            setattr(target, name, value)
        return target
