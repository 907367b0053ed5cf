"""Simulated ROS clock: ``Time``, ``Duration``, ``Rate`` and a process-wide clock
that tests and the adapter harness advance by hand.

Semantics needed by the reference (``prkt_core_v2.py:40,158,165,174``,
``prkt_ros.py:70``, ``test_prkt_ros2.py:53-55``): ``Time - Time -> Duration``,
``Time + Duration -> Time``, ``Duration.to_sec()``, integer ``Duration.secs``,
``Duration.from_sec``, ``Time.now().to_sec()``.  Times are kept as integer
nanoseconds, like genpy, so that ``last_update + dt`` is exact.
"""
from __future__ import annotations


def _split(nsecs_total: int):
    secs, nsecs = divmod(int(nsecs_total), 1000000000)
    return secs, nsecs


class _TVal(object):
    __slots__ = ("_ns",)

    def __init__(self, secs=0, nsecs=0):
        if isinstance(secs, float):
            whole = int(secs)
            nsecs = int(nsecs) + int(round((secs - whole) * 1e9))
            secs = whole
        self._ns = int(secs) * 1000000000 + int(nsecs)

    @property
    def secs(self):
        return _split(self._ns)[0]

    @property
    def nsecs(self):
        return _split(self._ns)[1]

    def to_sec(self):
        secs, nsecs = _split(self._ns)
        return float(secs) + float(nsecs) / 1e9

    def to_nsec(self):
        return self._ns

    def is_zero(self):
        return self._ns == 0

    @classmethod
    def from_sec(cls, float_secs):
        secs = int(float_secs)
        nsecs = int((float_secs - secs) * 1000000000)
        return cls(secs, nsecs)

    def __hash__(self):
        return hash((type(self).__name__, self._ns))

    def __eq__(self, other):
        return type(other) is type(self) and other._ns == self._ns

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._ns < other._ns

    def __le__(self, other):
        return self._ns <= other._ns

    def __gt__(self, other):
        return self._ns > other._ns

    def __ge__(self, other):
        return self._ns >= other._ns

    def __repr__(self):
        return "%s[%d]" % (type(self).__name__, self._ns)

    def __deepcopy__(self, memo):
        out = type(self)()
        out._ns = self._ns
        return out


class Duration(_TVal):
    __slots__ = ()

    def __add__(self, other):
        if isinstance(other, Time):
            return other + self
        if isinstance(other, Duration):
            return Duration(0, self._ns + other._ns)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, Duration):
            return Duration(0, self._ns - other._ns)
        return NotImplemented

    def __neg__(self):
        return Duration(0, -self._ns)

    def __mul__(self, k):
        return Duration(0, int(self._ns * k))


class Time(_TVal):
    __slots__ = ()

    @staticmethod
    def now():
        return clock.now()

    def __add__(self, other):
        if isinstance(other, Duration):
            return Time(0, self._ns + other._ns)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Time):
            return Duration(0, self._ns - other._ns)
        if isinstance(other, Duration):
            return Time(0, self._ns - other._ns)
        return NotImplemented


class SimClock(object):
    """Manually advanced clock; ``rospy.Time.now()`` of the fakes reads it."""

    def __init__(self):
        self._ns = 0
        self.sleep_advances = True

    def now(self):
        return Time(0, self._ns)

    def set(self, secs):
        self._ns = int(round(float(secs) * 1e9))

    def set_nsec(self, nsecs):
        self._ns = int(nsecs)

    def advance(self, secs):
        self._ns += int(round(float(secs) * 1e9))

    def advance_nsec(self, nsecs):
        self._ns += int(nsecs)


clock = SimClock()


class Rate(object):
    """``rospy.Rate`` on the simulated clock: ``sleep()`` advances the clock."""

    def __init__(self, hz):
        self.period_ns = int(round(1e9 / float(hz)))

    def sleep(self):
        if clock.sleep_advances:
            clock.advance_nsec(self.period_ns)
