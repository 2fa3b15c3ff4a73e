"""B200-native hot path of kalenforn/clip-based-cross-modal-hash (encode -> bit-packed codes ->
Hamming / top-k / mAP retrieval).  See DESIGN.md.  The CUDA library is loaded lazily by ``_lib``;
importing the package itself (and ``synth``) does not need a GPU.
"""
__version__ = "0.1.0"
