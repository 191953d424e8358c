class TriAnalyzer:  # imported by _triinterpolate.py:10, used only by the (dead) cubic interpolator
    def __init__(self, triangulation):
        self._triangulation = triangulation
