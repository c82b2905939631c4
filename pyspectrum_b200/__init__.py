"""pyspectrum_b200 -- B200 (sm_100a) implementation of pySpectrum's periodic-box estimator hot path.

    from pyspectrum_b200 import pyspectrum as pySpec      # Pk_periodic, Pk_periodic_rsd, Bk_periodic, ...
    from pyspectrum_b200 import estimator                 # f2py-shaped drop-in for `import estimator`
"""
from .pyspectrum import dat_dir  # noqa: F401

__all__ = ['pyspectrum', 'estimator', 'dat_dir']
