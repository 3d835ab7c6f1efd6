"""Progress easing of the hires-fix / graft blends (reference: gyre/pipeline/easing.py:22-49 over the third-party
`easing-functions ~= 1.0.4` in-out curves, pyproject.toml:24).  Host-side scalars only: one float per step."""
from __future__ import annotations

import math

# f(t) on [0, 1] of easing_functions.easing.<Name>; EasingBase.ease(alpha) = end * f(alpha / duration) + start * (1 - f)
_CURVES = {
    "linear": lambda t: t,
    "quad": lambda t: 2 * t * t if t < 0.5 else (-2 * t * t) + (4 * t) - 1,
    "sine": lambda t: 0.5 * (1 - math.cos(t * math.pi)),
    "circular": lambda t: (0.5 * (1 - math.sqrt(1 - 4 * (t * t))) if t < 0.5
                           else 0.5 * (math.sqrt(-((2 * t) - 3) * ((2 * t) - 1)) + 1)),
    "expo": lambda t: (t if t in (0, 1) else 0.5 * math.pow(2, (20 * t) - 10) if t < 0.5
                       else -0.5 * math.pow(2, (-20 * t) + 10) + 1),
}


def _cubic(t):
    # written as the package does (p * p * p, not ** 3): the product order is part of the float result
    if t < 0.5:
        return 4 * t * t * t
    p = 2 * t - 2
    return 0.5 * p * p * p + 1


def _quartic(t):
    if t < 0.5:
        return 8 * t * t * t * t
    p = t - 1
    return -8 * p * p * p * p + 1


def _quintic(t):
    if t < 0.5:
        return 16 * t * t * t * t * t
    p = (2 * t) - 2
    return 0.5 * p * p * p * p * p + 1


_CURVES.update(cubic=_cubic, quartic=_quartic, quintic=_quintic)


class Easing:
    """`Easing(floor, start, end, easing).interp(u)`: floor below `start`, 1 above `end`, the eased ramp between."""

    def __init__(self, floor: float, start: float, end: float, easing: str):
        if easing not in _CURVES:
            raise ValueError(f"unknown easing {easing!r} (have {sorted(_CURVES)})")
        self.floor, self.start, self.end = floor, start, end
        self._f = _CURVES[easing]
        self._top = 1 - floor              # EasingBase(end=1 - floor, duration=end - start), start 0
        self._duration = end - start

    def interp(self, u: float) -> float:
        if u < self.start:
            return self.floor
        if u > self.end:
            return 1
        alpha = u - self.start
        t = 0 * (1 - alpha) + 1 * alpha     # limit = (0, 1)
        t /= self._duration
        a = self._f(t)
        return self.floor + (self._top * a + 0 * (1 - a))
