"""The slice of ``rospy`` the reference node and core call."""
from __future__ import annotations

from .rostime import Duration, Rate, Time, clock

__all__ = ["Time", "Duration", "Rate", "Publisher", "Subscriber", "init_node", "loginfo",
           "logwarn", "logerr", "logdebug", "is_shutdown", "signal_shutdown", "get_time",
           "ROSInterruptException", "topics", "reset"]

_state = {"shutdown": False, "node": None}
topics = {}          # topic name -> list of subscriber callbacks
published = {}       # topic name -> count of messages published


class ROSInterruptException(Exception):
    pass


def init_node(name, *args, **kwargs):
    _state["node"] = name


def is_shutdown():
    return _state["shutdown"]


def signal_shutdown(reason=""):
    _state["shutdown"] = True


def reset():
    _state["shutdown"] = False
    _state["node"] = None
    topics.clear()
    published.clear()


def get_time():
    return clock.now().to_sec()


def loginfo(*args, **kwargs):
    pass


logwarn = logerr = logdebug = loginfo


class Publisher(object):
    """In-process pub/sub: ``publish`` calls every subscriber callback synchronously."""

    def __init__(self, name, data_class=None, queue_size=None, **kwargs):
        self.name = name
        self.data_class = data_class

    def publish(self, msg):
        published[self.name] = published.get(self.name, 0) + 1
        for cb in list(topics.get(self.name, ())):
            cb(msg)


class Subscriber(object):
    def __init__(self, name, data_class=None, callback=None, **kwargs):
        self.name = name
        self.data_class = data_class
        self.callback = callback
        if callback is not None:
            topics.setdefault(name, []).append(callback)

    def unregister(self):
        cbs = topics.get(self.name, [])
        if self.callback in cbs:
            cbs.remove(self.callback)
