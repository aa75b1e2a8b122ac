# Imported automatically by the spawn workers of tools/measure_reference.py (PYTHONPATH):
# the unmodified reference does `from numba import jitclass` (game/hex.py:6), a name that
# modern Numba moved to numba.experimental.
try:
    import numba
    import numba.experimental
    numba.jitclass = numba.experimental.jitclass
except Exception:       # not a numba environment: nothing to shim
    pass
