"""Prototype of the Ryu shortest-digits core (as it will run on the device) in exact Python integers, checked against
CPython repr (float64) and numpy Dragon4 (float32). The device kernel is a transliteration of d2d() below; the
128-bit tables are produced by gen_tables() (same definitions as the published Ryu generator)."""
import struct, random, sys
import numpy as np

POW5_INV_BITCOUNT = 125
POW5_BITCOUNT = 125

def pow5bits(e): return ((e * 1217359) >> 19) + 1
def log10Pow2(e): return (e * 78913) >> 18
def log10Pow5(e): return (e * 732923) >> 20

def gen_tables():
    inv = []
    for i in range(342):
        p = 5 ** i; j = pow5bits(i) - 1 + POW5_INV_BITCOUNT
        inv.append((1 << j) // p + 1)
    spl = []
    for i in range(326):
        p = 5 ** i; bl = pow5bits(i)
        assert bl == max(1, p.bit_length()), (i, bl, p.bit_length())
        sh = bl - POW5_BITCOUNT
        spl.append(p >> sh if sh >= 0 else p << -sh)
    return inv, spl

INV, SPL = gen_tables()

def mul_shift(m, mul, j):
    return (m * mul) >> j

def pow5factor(v):
    c = 0
    while v % 5 == 0 and v:
        v //= 5; c += 1
    return c

def d2d(m_ieee, e_ieee, mbits, bias):
    """returns (digits:int, exp10:int) with value = digits * 10^exp10"""
    if e_ieee == 0:
        e2 = 1 - bias - mbits - 2; m2 = m_ieee
    else:
        e2 = e_ieee - bias - mbits - 2; m2 = (1 << mbits) | m_ieee
    accept = (m2 & 1) == 0
    mv = 4 * m2
    mmShift = 1 if (m_ieee != 0 or e_ieee <= 1) else 0
    vmTZ = False; vrTZ = False
    if e2 >= 0:
        q = log10Pow2(e2) - (1 if e2 > 3 else 0)
        e10 = q
        k = POW5_INV_BITCOUNT + pow5bits(q) - 1
        i = -e2 + q + k
        vr = mul_shift(4 * m2, INV[q], i); vp = mul_shift(4 * m2 + 2, INV[q], i); vm = mul_shift(4 * m2 - 1 - mmShift, INV[q], i)
        if q <= 21:
            if mv % 5 == 0: vrTZ = pow5factor(mv) >= q
            elif accept: vmTZ = pow5factor(mv - 1 - mmShift) >= q
            else: vp -= 1 if pow5factor(mv + 2) >= q else 0
    else:
        q = log10Pow5(-e2) - (1 if -e2 > 1 else 0)
        e10 = q + e2
        i = -e2 - q
        k = pow5bits(i) - POW5_BITCOUNT
        j = q - k
        vr = mul_shift(4 * m2, SPL[i], j); vp = mul_shift(4 * m2 + 2, SPL[i], j); vm = mul_shift(4 * m2 - 1 - mmShift, SPL[i], j)
        if q <= 1:
            vrTZ = True
            if accept: vmTZ = mmShift == 1
            else: vp -= 1
        elif q < 63:
            vrTZ = (mv & ((1 << q) - 1)) == 0
    removed = 0; last = 0
    if vmTZ or vrTZ:
        while vp // 10 > vm // 10:
            vmTZ &= vm % 10 == 0; vrTZ &= last == 0
            last = vr % 10; vr //= 10; vp //= 10; vm //= 10; removed += 1
        if vmTZ:
            while vm % 10 == 0:
                vrTZ &= last == 0
                last = vr % 10; vr //= 10; vp //= 10; vm //= 10; removed += 1
        if vrTZ and last == 5 and vr % 2 == 0: last = 4
        out = vr + (1 if ((vr == vm and (not accept or not vmTZ)) or last >= 5) else 0)
    else:
        roundUp = False
        while vp // 10 > vm // 10:
            roundUp = vr % 10 >= 5
            vr //= 10; vp //= 10; vm //= 10; removed += 1
        out = vr + (1 if (vr == vm or roundUp) else 0)
    return out, e10 + removed

def digits_of(x):
    """(digit string without trailing zeros, decimal point position) like the oracle's Digits"""
    r = repr(x)
    mant, _, ex = r.partition("e")
    ex = int(ex) if ex else 0
    if "." in mant: ip, fp = mant.split(".")
    else: ip, fp = mant, ""
    ds = (ip + fp).lstrip("0"); lead = len((ip + fp)) - len((ip + fp).lstrip("0"))
    dp = len(ip) - lead + ex
    ds = ds.rstrip("0") or "0"
    return ds, dp

if __name__ == "__main__":
    rnd = random.Random(1)
    bad = 0
    vals = [5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1.0, 0.1, 0.3, 1e21, 1e22, 1e23, 9007199254740993.0, 123456.7, 4.35, 0.5, 2.0**-1074, 2.0**1023, 1e-5, 5e-5, 1.5e300, 7.0e-310]
    vals += [struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64) & 0x7FEFFFFFFFFFFFFF))[0] for _ in range(300000)]
    vals += [rnd.random() * 10 ** rnd.randint(-30, 30) for _ in range(100000)]
    vals += [float(rnd.randint(1, 10**rnd.randint(1, 17))) for _ in range(50000)]
    for v in vals:
        if v == 0 or v != v or v in (float("inf"),): continue
        bits = struct.unpack("<Q", struct.pack("<d", v))[0]
        out, e10 = d2d(bits & ((1 << 52) - 1), (bits >> 52) & 0x7FF, 52, 1023)
        s = str(out); ds = s.rstrip("0") or "0"; dp = len(s) + e10
        if (ds, dp) != digits_of(v):
            bad += 1
            if bad < 10: print("MISMATCH f64", v, (ds, dp), digits_of(v))
    print("float64 checked", len(vals), "bad", bad)
    bad = 0
    fb = np.random.default_rng(2).integers(1, 0x7F7FFFFF, 300000, dtype=np.uint32)
    for b in fb.tolist() + [1, 0x00800000, 0x7F7FFFFF, 0x3F800000, 0x00400000]:
        f = np.uint32(b).view(np.float32) if isinstance(b, int) else b
        f = np.array([b], dtype=np.uint32).view(np.float32)[0]
        out, e10 = d2d(b & ((1 << 23) - 1), (b >> 23) & 0xFF, 23, 127)
        s = str(out); ds = s.rstrip("0") or "0"; dp = len(s) + e10
        want = np.format_float_scientific(f, unique=True, trim="-")
        wm, _, we = want.partition("e"); wds = wm.replace(".", "").rstrip("0") or "0"; wdp = int(we) + 1
        if (ds, dp) != (wds, wdp):
            bad += 1
            if bad < 10: print("MISMATCH f32", f, (ds, dp), (wds, wdp))
    print("float32 checked", len(fb), "bad", bad)
