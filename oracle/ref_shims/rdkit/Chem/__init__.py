"""TEST INFRASTRUCTURE ONLY — stub: the reference's hot path only needs the *name* ``Mol``
(``notorch/types.py:5``, reached from ``notorch/nn/gnn/chemprop.py:10``)."""


class Mol:  # pragma: no cover - never instantiated
    pass


class Atom:  # pragma: no cover
    pass


class Bond:  # pragma: no cover
    pass
