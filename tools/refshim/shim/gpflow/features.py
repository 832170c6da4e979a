from .params import Parameter
from .dispatch import _registry, _Dispatcher


class InducingFeature:
    pass


class InducingPointsBase(InducingFeature):
    def __init__(self, Z):
        self.Z = Parameter(Z)

    def __len__(self):
        return self.Z.shape[0]


Kuu = _registry.setdefault("Kuu", _Dispatcher("Kuu"))
Kuf = _registry.setdefault("Kuf", _Dispatcher("Kuf"))
