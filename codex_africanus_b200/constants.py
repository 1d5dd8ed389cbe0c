"""Physical constants, bit-identical to africanus/constants/consts.py:6-9."""
import math

c = 2.99792458e8
two_pi_over_c = 2 * math.pi / c
minus_two_pi_over_c = -two_pi_over_c
