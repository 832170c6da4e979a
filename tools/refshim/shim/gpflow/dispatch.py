"""First-match isinstance dispatch with a global namespace, like gpflow.dispatch (multipledispatch) for our needs."""
_registry = {}


class _Dispatcher:
    def __init__(self, name):
        self.name, self.impls = name, []

    def __call__(self, *args, **kw):
        for types, fn in self.impls:
            if len(types) == len(args) and all(isinstance(a, t) for a, t in zip(args, types)):
                return fn(*args, **kw)
        raise NotImplementedError("no %s for %r" % (self.name, tuple(type(a) for a in args)))


def dispatch(*types):
    def deco(fn):
        d = _registry.setdefault(fn.__name__, _Dispatcher(fn.__name__))
        d.impls.insert(0, (types, fn))
        return d
    return deco
