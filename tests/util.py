"""Shared helpers for the test-suite (oracle side)."""
import json
import math
import os

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEFAULT_RANGES = ((0, 0.1), (0, 0.1), (-0.25, 0.25), (0.001, 0.1))
CFG1 = "8|5;9|113;10|89"
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
CFG3 = ";".join(
    "%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) + 300 * math.exp(-0.5 * ((a - 36) / 5) ** 2)))
    for a in range(10, 50))


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def cfg_from_case(noise_coef, kw, **extra):
    rg = kw.get("ranges", DEFAULT_RANGES)
    return orc.make_cfg(noise_coef=noise_coef, random_start=kw.get("random", False),
                        up_preference=kw.get("up_preference", False), floor_opt=kw.get("floor", 5),
                        n_cycles=kw.get("n_cycles", 24), capture_cycle=kw.get("capture_cycle", 13),
                        start_copies=kw.get("start_copies", 300000), down=rg[0], up=rg[1], capture=rg[2],
                        efficiency=rg[3], **extra)
