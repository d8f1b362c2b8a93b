"""rawcooked_b200 — B200-native FFV1 (+FLAC) encode path behind RAWcooked.

The product is the C-ABI shared library `libb200enc.so` (include/b200enc.h) built from
rawcooked_b200/csrc/ by `__graft_entry__.build()`; this package is the thin host-side mirror used by the
tests, bench.py and Python callers. There is no CPU fallback: importing the encoder classes without the
built library, or encoding without a CUDA device, raises."""
__version__ = "0.1.0"
