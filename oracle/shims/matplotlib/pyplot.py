"""TEST INFRASTRUCTURE ONLY. matplotlib.pyplot stub: the reference only plots behind
verbosity flags (gpet.py:666-764); every attribute access raises so a plot call is loud."""


def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name} called: plotting is stubbed out in the oracle harness")
