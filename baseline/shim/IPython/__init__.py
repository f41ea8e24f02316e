"""Empty stand-in: the reference imports IPython.display.clear_output for verbose='plot' only."""
