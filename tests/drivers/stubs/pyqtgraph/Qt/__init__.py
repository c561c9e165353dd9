"""QtCore / QtGui / QtWidgets stand-ins: QTimer callbacks are run back to back by QApplication.exec() until one of them
leaves through exit() (the reference scripts stop that way at the end of the input, show_results_from_model.py:141-142)."""
import types

_timers = []
MAX_TICKS = 1000000


class _Signal(object):
    def __init__(self):
        self.slots = []

    def connect(self, fn):
        self.slots.append(fn)


class QTimer(object):
    def __init__(self, *a, **k):
        self.timeout = _Signal()
        self.period = None

    def start(self, period=0):
        self.period = period
        _timers.append(self)

    def stop(self):
        if self in _timers:
            _timers.remove(self)


class QApplication(object):
    _instance = None

    def __init__(self, argv=None):
        QApplication._instance = self

    @staticmethod
    def instance():
        return QApplication._instance

    def exec(self):
        for _ in range(MAX_TICKS):
            if not _timers:
                return 0
            for t in list(_timers):
                for fn in t.timeout.slots:
                    fn()                      # exceptions propagate: a real Qt slot that raises aborts the viewer too
        return 0

    exec_ = exec


QtCore = types.SimpleNamespace(QTimer=QTimer)
QtGui = types.SimpleNamespace()
QtWidgets = types.SimpleNamespace(QApplication=QApplication)
