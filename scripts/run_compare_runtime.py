#!/usr/bin/env python3
"""Runtime vs horizon, strict vs non-strict (reference scripts/run_compare_runtime.py), batched.

The reference times ONE tick (x-solve + y-solve) of one walk for N = 10..300 with x = 0 and +-0.1 bounds
(3 warm-ups, 10 timed calls, mean; its Herdt leg crashes -- SURVEY 3.5).  Here one tick is timed for a
batch of independent axis problems with non-trivial inputs (BASELINE config 3), CUDA events, median.

    python scripts/run_compare_runtime.py --horizons 100 150 300 600 --batch 16384
"""

import os
import sys

from _common import REPO

sys.path.insert(0, os.path.join(REPO, "tools"))
import horizon_scaling  # noqa: E402

if __name__ == "__main__":
    horizon_scaling.main()
