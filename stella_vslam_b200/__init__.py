"""B200-native implementation of stella_vslam's per-frame hot path (ORB extract -> Hamming match -> local BA).

The compute lives in hand-written sm_100a kernels behind the C ABI of include/b200vslam.h (libb200vslam.so, built
in-tree by __graft_entry__.build()).  This package is the host-side mirror of the reference's class surfaces.
"""
__version__ = "0.1.0"
