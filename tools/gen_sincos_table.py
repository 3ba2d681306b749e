"""Generates code-robchar_b200/csrc/rc_sincos_table.inc: sin(j pi/32), cos(j pi/32) for j = 0..63, correctly
rounded to double (50-digit decimal Taylor series), used by rc_sincos_tab (rc_math.cuh)."""
import os
from decimal import Decimal, getcontext

getcontext().prec = 50
PI = Decimal("3.14159265358979323846264338327950288419716939937510")


def dsin(x):
    term = s = x
    n = 1
    while abs(term) > Decimal(10) ** -45:
        term = -term * x * x / ((2 * n) * (2 * n + 1)); s += term; n += 1
    return s


def dcos(x):
    term = s = Decimal(1)
    n = 1
    while abs(term) > Decimal(10) ** -45:
        term = -term * x * x / ((2 * n - 1) * (2 * n)); s += term; n += 1
    return s


lines = ["// sin(j pi/32), cos(j pi/32), j = 0..63, correctly rounded doubles (tools/gen_sincos_table.py)"]
for j in range(64):
    a = PI * j / 32
    s, c = float(dsin(a)), float(dcos(a))
    s = 0.0 if abs(s) < 1e-30 else s
    c = 0.0 if abs(c) < 1e-30 else c
    lines.append("    {%s, %s}," % (float.hex(s), float.hex(c)))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "code-robchar_b200", "csrc", "rc_sincos_table.inc")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", os.path.normpath(out))
