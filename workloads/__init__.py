"""Synthetic inputs (seeded image streams, descriptor sets, BA windows) shared by tests/, bench.py, tools/ and smoke().
Not part of the product package: stella_vslam_b200/ holds only the hot path and its host-side interface."""
