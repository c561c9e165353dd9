"""Headless stand-in for pyqtgraph (absent from this image): enough of the API for the reference's
test/show_results_from_model.py / show_results_from_triangulation.py to run their Qt timer loop without a display.
Everything handed to the scene is appended to _RECORD so a test can compare what would have been drawn."""
_RECORD = []

_COLORS = {'r': (1., 0., 0., 1.), 'g': (0., 1., 0., 1.), 'b': (0., 0., 1., 1.), 'm': (1., 0., 1., 1.), 'c': (0., 1., 1., 1.),
           'y': (1., 1., 0., 1.), 'd': (.6, .6, .6, 1.), 'k': (0., 0., 0., 1.), 'w': (1., 1., 1., 1.)}


def glColor(c):
    return _COLORS.get(c, (0., 0., 0., 1.)) if isinstance(c, str) else tuple(c)
