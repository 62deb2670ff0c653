"""numpy float32 restatement of Utility::fast_atan2f (include/Utility.h:236-304), test helper."""
import re
import os

import numpy as np

_tbl = None


def table():
    global _tbl
    if _tbl is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "airspy_fmradion_b200", "csrc",
                         "fmr_tables_generated.inc")
        s = open(p).read()
        m = re.search(r"k_fast_atan_table\[(\d+)\] = \{(.*?)\};", s, re.S)
        _tbl = np.array([float(x) for x in m.group(2).replace("\n", " ").split(",") if x.strip()], dtype=np.float32)
        assert len(_tbl) == int(m.group(1)) == 257
    return _tbl


def fast_atan2f_np(y, x):
    f = np.float32
    t = table()
    ya, xa = f(abs(y)), f(abs(x))
    if not (ya > 0 or xa > 0):
        return f(0)
    z = f(ya / xa) if ya < xa else f(xa / ya)
    if float(z) < 0.003921569:
        base = z
    else:
        alpha = f(z * f(255))
        idx = int(alpha) & 0xff
        alpha = f(alpha - f(idx))
        base = f(t[idx] + f(f(t[idx + 1] - t[idx]) * alpha))
    if xa > ya:
        if x >= 0:
            return base if y >= 0 else f(-base)
        a = f(3.14159265358979323846)
        return f(a - base) if y >= 0 else f(base - a)
    if y >= 0:
        a = f(1.57079632679489661923)
        return f(a - base) if x >= 0 else f(a + base)
    a = f(-1.57079632679489661923)
    return f(a + base) if x >= 0 else f(a - base)
