"""shermbot-navigation_b200: B200-native batched EKF-SLAM + scan circle-detection engine.

Drop-in for the hot path of ``sziselman/Shermbot-Navigation``'s ``nuslam`` / ``rigid2d`` libraries
(see DESIGN.md): hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/nuslam_b200.h``; this package is the host-side mirror of the reference's library API.
"""
__version__ = "0.1.0"
