"""pyqtgraph.opengl stand-ins: a scene that remembers its items."""
import numpy as np

from .. import _RECORD


class _Item(object):
    def __init__(self, **kw):
        self.kw = kw

    def setColor(self, *a): pass
    def setSize(self, *a, **k): pass
    def rotate(self, *a): pass
    def translate(self, *a): pass
    def setGLOptions(self, *a): pass


class GLGridItem(_Item):
    pass


class GLScatterPlotItem(_Item):
    def __init__(self, **kw):
        _Item.__init__(self, **kw)
        _RECORD.append(('points', np.array(kw.get('pos'), dtype=np.float64).tolist()))

    def setData(self, **kw):
        self.kw.update(kw)
        _RECORD.append(('points', np.array(kw.get('pos'), dtype=np.float64).tolist()))


class GLLinePlotItem(_Item):
    def __init__(self, **kw):
        _Item.__init__(self, **kw)
        _RECORD.append(('line', np.array(kw.get('pos'), dtype=np.float64).tolist()))


class GLViewWidget(object):
    def __init__(self, *a, **k):
        self.opts = {}
        self.items = []

    def setBackgroundColor(self, *a): pass
    def setWindowTitle(self, *a): pass
    def setGeometry(self, *a): pass
    def show(self): pass

    def addItem(self, it):
        self.items.append(it)

    def removeItem(self, it):
        self.items.remove(it)
