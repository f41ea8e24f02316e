"""Empty stand-in: the reference imports matplotlib.pyplot for its plotting helpers only."""
