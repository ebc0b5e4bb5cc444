"""TEST INFRASTRUCTURE ONLY -- ``dgl.function`` builtins used by the reference
(models/RGCN.py:4, 101: ``fn.sum(msg='msg', out='h')``)."""


class _SumReducer(object):
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _SumReducer(msg, out)
