"""`from jaxtyping import Array, Float` (functionals.py:2): annotation-only names."""


class _Sub:
    def __class_getitem__(cls, item):
        return cls


class Array(_Sub):
    pass


class Float(_Sub):
    pass
