"""sailfish_b200: the collide-and-propagate hot path of sailfish-team/sailfish on NVIDIA B200.

``csrc/`` holds the CUDA kernels and the C ABI (``libslf_b200.so``, built by ``python -m sailfish_b200.build``);
``backend_b200`` is the ctypes binding with the reference's backend interface; ``plugin`` / ``run`` plug it
into an unmodified reference checkout; the remaining modules are the self-contained host side (controller,
subdomain runner, geometry) used where no reference checkout exists."""
