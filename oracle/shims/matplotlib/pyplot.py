def __getattr__(name):
    def _stub(*a, **k):
        raise RuntimeError(f"matplotlib.pyplot.{name} is a stub in the oracle harness")
    return _stub
