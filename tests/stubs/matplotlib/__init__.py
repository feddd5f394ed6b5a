"""No-op stand-in for matplotlib (absent from this image) so the reference's UNCHANGED driver scripts can be
imported by tests/test_unchanged_drivers.py; their plotting functions are never called by the tests."""


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def use(*a, **k):
    return None


def __getattr__(name):
    return _Anything()
