"""projectd_core_b200 -- B200-native batched vehicle-physics core (drop-in for the Car::step hot path of
wongfei/projectd-core).

The compute path is the CUDA library ``libpd_b200.so`` (kernels for sm_100a + the C ABI declared in
``include/pd_batch.h``).  This package is the thin host-side mirror of the reference's Python-facing surface:

* :class:`Batch` -- ctypes binding of the C ABI (one batch = N simulators with one car each);
* :mod:`projectd_core_b200.pyprojectd` -- the reference's ``PyProjectD`` function names
  (``createSimulator``, ``loadTrack``, ``addCar``, ``setCarControls``, ``stepSimulator``, ``getCarState`` ...,
  src/PyProjectD/PyProjectD.cpp:515-641) on top of it, plus batched variants;
* :class:`projectd_core_b200.env.BatchedProjectDEnv` -- the vectorised ``ProjectDEnv``
  (pyprojectd/projectd_env.py) with observations handed out through DLPack as torch CUDA tensors.

There is no CPU fallback: importing works anywhere, but creating a batch without the CUDA library or without a
GPU raises.
"""
from .binding import Batch, PdError, lib_path, load_library, STATE_WORDS, OBS_DIM  # noqa: F401

from . import dist  # noqa: F401,E402

__all__ = ["Batch", "PdError", "lib_path", "load_library", "STATE_WORDS", "OBS_DIM", "dist"]
