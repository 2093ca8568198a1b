#!/usr/bin/env python
"""tests/golden/make_wrench_table.py -- the wrench series the reference replays in mode 2 of applyBodyWrench
(bluerov2_dobmpc/src/bluerov2_dob.cpp:818-874: config/forcex.txt, forcey.txt, forcez.txt, torquez.txt, one value per line, all four
read at the same counter) stacked column-wise into tests/golden/wrench_table.npz ([496, 4] = fx, fy, fz, tz).  Run where
/root/reference exists; the fixture travels, the reference does not."""
import os
import numpy as np

SRC = "/root/reference/bluerov2_dobmpc/config"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    cols = [np.loadtxt(os.path.join(SRC, f)) for f in ("forcex.txt", "forcey.txt", "forcez.txt", "torquez.txt")]
    assert len({c.shape for c in cols}) == 1
    table = np.stack(cols, axis=1)
    np.savez_compressed(os.path.join(HERE, "wrench_table.npz"), table=table)
    print(table.shape, table.mean(0), table.std(0))
