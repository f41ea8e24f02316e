"""Empty stand-in: the reference's metrics import skimage.measure (surface areas); unused by the solvers."""
