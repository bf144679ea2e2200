def _noop(*a, **k):
    raise RuntimeError("matplotlib stub: plotting is out of scope for the golden-vector generator")


subplots = plot = savefig = close = _noop
