"""Minimal ``Manifold`` base: the registry that lets ``{"type": "Spline"}`` JSON round-trip
(reference ``bspy/manifold.py:22, 295-301``) and the metadata slot.  The CSG interface of the
reference's Manifold (intersect, cached_intersect, ...) is outside the evaluation path and is
not provided."""


class Manifold:
    minSeparation = 0.0001
    factory = {}

    def __init__(self, metadata=None):
        self.metadata = dict(metadata or {})

    @staticmethod
    def register(cls):
        """Class decorator: make ``cls`` constructible from ``Manifold.factory[cls.__name__]``."""
        Manifold.factory[cls.__name__] = cls
        return cls

    @staticmethod
    def from_dict(dictionary):
        return Manifold.factory[dictionary.get("type", "Spline")].from_dict(dictionary)

    def copy(self):
        raise NotImplementedError

    def domain_dimension(self):
        return 0

    def range_dimension(self):
        return 0

    def evaluate(self, domainPoint):
        raise NotImplementedError

    def normal(self, domainPoint, normalize=True, indices=None):
        raise NotImplementedError

    def tangent_space(self, domainPoint):
        raise NotImplementedError

    def to_dict(self):
        raise NotImplementedError
