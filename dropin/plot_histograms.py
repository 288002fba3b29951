"""Import-only stub: the reference's plot_histograms.py needs matplotlib (absent) and is diagnostics,
out of scope.  noisynet.py imports these names at module level."""


def _unavailable(*a, **k):
    raise NotImplementedError("plot_histograms diagnostics are out of scope of noisynet_b200")


get_layers = plot = plot_layers = plot_grid = _unavailable
