"""Inert stand-in for matplotlib.

TEST INFRASTRUCTURE ONLY.  The reference package imports matplotlib at import time
(plotting helpers), but matplotlib is not installed in the build container.  This stub
lets ``oracle/make_golden.py`` import the unmodified reference from /root/reference to
generate golden vectors.  It is never imported by the product package.
"""


class _Any:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        return _Any()

    def __call__(self, *args, **kwargs):
        return _Any()

    def __or__(self, other):
        return self

    def __ror__(self, other):
        return self


from . import axis, figure, patches, pyplot, ticker  # noqa: E402,F401


def __getattr__(name):  # defined last: `from . import x` probes getattr first
    return _Any()
