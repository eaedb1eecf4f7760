"""microfc_b200 -- B200-native (sm_100a) right-hand side + TVD-RK time stepping for MicroFC.

Only what the hot path needs lives here:

* ``csrc/``       hand-written FP64 CUDA kernels + the C ABI (``libmfc_b200.so``)
* ``abi``         ctypes binding of ``include/mfc_b200.h``
* ``simulation``  host driver that plays the reference's ``p_main`` time loop over the ABI
* ``case`` / ``cases`` / ``pre_process`` / ``domain``  case files, initial conditions,
  domain decomposition and grid metrics (host side of the path, as in the reference)
"""
__version__ = "0.1.0"
