#!/usr/bin/env python
"""`python sampler.py ...` -- the reference's DuoDiff sampling CLI (sampler.py:192-352: same flags, same output files)
on the B200-native path.  A thin shim over duodiff_b200.sampler so that the reference's command lines (README.md:104-
118) work unchanged from this repository's root; `get_samples`, the post-processing rule objects and the dump helpers
are re-exported under the reference's names."""
from duodiff_b200.sampler import (dump_samples, dump_statistics, get_args, get_samples, main,  # noqa: F401
                                  predict_noise_postprocessing, predict_original_postprocessing,
                                  predict_previous_postprocessing)

if __name__ == "__main__":
    main()
