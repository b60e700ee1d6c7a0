"""Shared replay helper: drive any env with the vector calling convention through a golden file."""
import ast
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64 = ("lat", "lon", "heading", "speed", "new_heading", "new_speed", "cannon_remain", "cannon_burst",
       "cannon_max")
R64 = ("r_lat", "r_lon", "r_heading", "r_new_heading", "r_speed")
I32 = ("missile_remain", "rocket_max", "missile_wait", "alive", "has_missile", "opp_to_attack")
RI32 = ("r_alive", "r_target", "r_age")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "lowlevel_*.npz")))


def load(path):
    g = dict(np.load(path))
    seed, arena, level, mode = (int(x) for x in g["meta"])
    kw = ast.literal_eval(str(g["kw"]))
    return g, seed, arena, level, ("fight" if mode == 0 else "escape"), kw


def rel_close(a, b, rtol, atol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))
