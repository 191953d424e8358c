def __getattr__(name):  # plotting is out of scope; any use is an error
    raise AttributeError(f"matplotlib.pyplot.{name} is not available in the oracle stub")
