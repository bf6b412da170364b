"""Minimal ``glom`` stand-in (test infrastructure only) covering the specs
used by /root/reference/src/krotov/convergence.py:109-310: attribute/key path
tuples and ``T[index]``."""


class GlomError(Exception):
    pass


class _T:
    def __init__(self, ops=()):
        self._ops = ops

    def __getitem__(self, item):
        return _T(self._ops + (('item', item),))

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return _T(self._ops + (('attr', name),))


T = _T()


def glom(target, spec, **kwargs):
    if isinstance(spec, tuple):
        for s in spec:
            target = glom(target, s)
        return target
    if isinstance(spec, _T):
        for kind, arg in spec._ops:
            target = target[arg] if kind == 'item' else getattr(target, arg)
        return target
    if isinstance(spec, str):
        for part in spec.split('.'):
            try:
                target = getattr(target, part)
            except AttributeError:
                target = target[part]
        return target
    if callable(spec):
        return spec(target)
    raise GlomError("unsupported spec %r" % (spec,))
