"""Import stand-in for matplotlib (not installed here).  The reference pulls it in at import time via
encoder/RateControl/lookup.py:10 -> metrics/plot_rd_curves.py; nothing on the hot path calls it."""
