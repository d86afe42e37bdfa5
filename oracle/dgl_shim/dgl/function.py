"""Builtin message / reduce function tags (`dgl.function`), only the ones the reference names."""


class _Binary:
    def __init__(self, name, lhs, rhs, out):
        self.name, self.lhs, self.rhs, self.out = name, lhs, rhs, out


class _Reduce:
    def __init__(self, name, msg, out):
        self.name, self.msg, self.out = name, msg, out


def u_add_v(lhs, rhs, out):
    return _Binary('u_add_v', lhs, rhs, out)


def u_mul_e(lhs, rhs, out):
    return _Binary('u_mul_e', lhs, rhs, out)


def copy_u(u, out):
    return _Binary('copy_u', u, None, out)


def sum(msg, out):  # noqa: A001  (dgl.function.sum shadows the builtin on purpose)
    return _Reduce('sum', msg, out)
