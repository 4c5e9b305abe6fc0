"""Host-side float32 vector helpers (reference graphics/vector.py host functions). The
device twins of the reference (``d_*``) have no counterpart here: they are inlined in the
CUDA tracer (csrc/rf_tracer.cuh)."""

import numpy

V2F = tuple[numpy.float32, numpy.float32]
V3F = tuple[numpy.float32, numpy.float32, numpy.float32]


def v2f(x: float = 0.0, y: float = 0.0) -> V2F:
    """reference vector.py:15-25"""

    return (numpy.float32(x), numpy.float32(y))


def v3f(x: float = 0.0, y: float = 0.0, z: float = 0.0) -> V3F:
    """reference vector.py:42-53"""

    return (numpy.float32(x), numpy.float32(y), numpy.float32(z))


def add_v3f(summands: tuple[V3F, ...]) -> V3F:
    """reference vector.py:102-113 (numpy.sum over axis 0: left-to-right float32 adds)"""

    s = numpy.sum(summands, axis=0)
    return (s[0], s[1], s[2])


def sub_v3f(a: V3F, b: V3F) -> V3F:
    """reference vector.py:150-162"""

    r = numpy.subtract(a, b)
    return (r[0], r[1], r[2])


def smul_v3f(v: V3F, s: float) -> V3F:
    """reference vector.py:193-205 (a Python-float ``s`` is weak: float32 multiply)"""

    r = numpy.multiply(v, s)
    return (r[0], r[1], r[2])


def cross_v3f(a: V3F, b: V3F) -> V3F:
    """reference vector.py:268-279"""

    c = numpy.cross(numpy.asarray(a), numpy.asarray(b))
    return (c[0], c[1], c[2])


def length_v3f(vector: V3F) -> float:
    """reference vector.py:317-327"""

    return float(numpy.linalg.norm(numpy.asarray(vector)))


def norm_v3f(vector: V3F) -> V3F:
    """reference vector.py:342-351"""

    return smul_v3f(vector, 1.0 / length_v3f(vector))
