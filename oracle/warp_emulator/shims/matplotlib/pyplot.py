"""Inert stand-in for matplotlib.pyplot (the reference's compare() plots its sampled lists)."""


class _Ax:
    def __getattr__(self, n): return lambda *a, **k: None


def subplots(n=1, *a, **k): return _Ax(), tuple(_Ax() for _ in range(n))
def show(*a, **k): pass
PLOTS = []          # every plt.plot(x, y, ..., label=...) call, so that drivers which only plot (drag) can be read back


def plot(*a, **k): PLOTS.append((a, k))
def figure(*a, **k): return _Ax()


def __getattr__(name):                                   # xlabel, ylabel, legend, title, ...: inert
    return lambda *a, **k: None
