"""Empty stand-in: ``cggp/covertree.py`` imports ``matplotlib.pyplot`` and never uses it."""
