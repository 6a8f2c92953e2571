"""Table and constants of log10_fast (crisp_b200/csrc/dmath.cuh): rc = fl(1/c), -log10(rc) per top-seven-mantissa-bits cell,
log10(2) split hi/lo (hi exact times an 11-bit exponent), Taylor coefficients of log10(1+r).  Prints C initialisers."""
from decimal import Decimal, getcontext
from fractions import Fraction
import struct
getcontext().prec = 60
def todouble(d): return float(d)   # Decimal -> nearest double (Python rounds correctly)
rows=[]
for i in range(128):
    if i < 53:
        c = Fraction(1) + Fraction(2*i+1, 256) if i != 0 else Fraction(1)
    else:
        c = (Fraction(1) + Fraction(2*i+1, 256)) / 2 if i != 127 else Fraction(1)
    rc = float(Fraction(1)/c)          # correctly rounded
    rcd = Decimal(Fraction(rc).numerator) / Decimal(Fraction(rc).denominator)
    lc = -(rcd.log10())
    rows.append((rc, float(lc)))
l2 = Decimal(2).log10()
hi = float(l2)
b = struct.unpack('<Q', struct.pack('<d', hi))[0] & ~((1<<12)-1)
hi = struct.unpack('<d', struct.pack('<Q', b))[0]
lo = float(l2 - Decimal(Fraction(hi).numerator)/Decimal(Fraction(hi).denominator))
le = Decimal(1)/Decimal(10).ln()
K = [float(le/Decimal(k) * (1 if k%2 else -1)) for k in range(1,10)]
import sys
if True:
    f = sys.stdout
    f.write('static const double kLog10Tab[256] = {\n')
    for rc,lc in rows: f.write(f'\t{rc!r}, {lc!r},\n')
    f.write('};\n')
    f.write(f'static const double kL2hi = {hi!r}, kL2lo = {lo!r};\n')
    f.write('static const double kLogK[9] = {' + ', '.join(repr(k) for k in K) + '};\n')
