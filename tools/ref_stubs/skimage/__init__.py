"""Stand-in package: skimage.measure is only used by the reference mesh export (utils.py:165)."""
