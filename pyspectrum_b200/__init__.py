"""pyspectrum_b200 -- B200 (sm_100a) implementation of pySpectrum's FFT-based estimator hot path (periodic box and survey geometry).

    from pyspectrum_b200 import pyspectrum as pySpec      # Pk_periodic, Pk_periodic_rsd, Bk_periodic, B0_survey, ...
    from pyspectrum_b200 import estimator                 # f2py-shaped drop-in for `import estimator`
    from pyspectrum_b200 import util                      # ijl_order, radecz_to_cartesian, applyRSD (pyspectrum/util.py)
    from pyspectrum_b200 import multigpu                  # one catalogue sharded over the GPUs of a node
"""
from .pyspectrum import dat_dir  # noqa: F401

__all__ = ['pyspectrum', 'estimator', 'util', 'multigpu', 'dat_dir']
