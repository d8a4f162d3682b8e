#!/usr/bin/env python
"""`python eesampler.py ...` -- the reference's DeeDiff / AdaDiff early-exit sampling CLI (eesampler.py:114-209: same
flags, same output files incl. the two .pt logs) on the B200-native path.  A thin shim over duodiff_b200.eesampler."""
from duodiff_b200.eesampler import dump_samples, dump_statistics, get_args, get_samples, main  # noqa: F401

if __name__ == "__main__":
    main()
