"""B200-native batched LIPM-MPC.  ``csrc/`` holds the CUDA kernels and the C ABI,
``mpc_bipedal/`` the host-side mirror of the reference package.  Put this directory on
``sys.path`` and ``import mpc_bipedal`` (the directory name itself is not an identifier)."""
