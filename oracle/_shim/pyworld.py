"""Stand-in so that `voice100.vocoder` imports in the build container (pyworld is a C extension that is not
installed here).  Only the pure-numpy helpers of that module (create_mc2sp_matrix, freqt) are used by
oracle/gen_golden.py; any attempt to run WORLD analysis/synthesis fails loudly."""


def __getattr__(name):
    raise RuntimeError(f"pyworld.{name} is not available in this container (stand-in module)")
