"""magpurify2_b200 -- B200-native (sm_100a) composition / coverage scoring paths of MAGpurify2.

The package mirrors the reference's Python surface for these two paths only:

    magpurify2_b200._composition.get_tnf / get_gc     <- magpurify2._composition  (src/composition.rs)
    magpurify2_b200._coverage.get_bam_coverages       <- magpurify2._coverage     (src/coverage.rs)
    magpurify2_b200.tools / core / modules / cli      <- the callers either side of those paths

All compute goes through libmp2.so (include/mp2.h).  There is no CPU fallback: importing works
anywhere, but every compute call raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
