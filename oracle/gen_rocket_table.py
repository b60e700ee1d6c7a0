"""oracle/gen_rocket_table.py -- TEST INFRASTRUCTURE.

Prints the 11 float64 values of the reference's rocket speed profile
(warsim/simulator/rocket_unit.py:16-21: scipy interp1d(kind='quadratic') through
(0,500),(10,2000),(20,1400),(30,600) knots) at life_time = 0..10 s, as C hex-float
literals.  These are baked into oracle/hhmarl_oracle.c (ROCKET_SPEED) and the CUDA kernels.
"""
import numpy as np
from scipy.interpolate import interp1d

sp = interp1d(np.array([0, 10, 20, 30]), np.array([500, 2000, 1400, 600]), kind="quadratic",
              assume_sorted=True, bounds_error=False, fill_value=(500, 600))
for t in range(11):
    v = float(sp(t))
    print(t, repr(v), v.hex())
