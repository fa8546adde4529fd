"""Pure-Python transcription of include/simian_spec.h (RNG, portable log, trace
hash) and of the PHOLD handler in include/simian_models.h.

TEST INFRASTRUCTURE ONLY.  Used (a) by oracle/pin_simianpie.py as the model
code driven by the UNMODIFIED reference engine SimianPie/simian.py, and (b) by
tests to cross-check the C++ header bit-for-bit.  CPython floats are IEEE
binary64 with one rounding per operation (no FMA contraction), which is what
the header requires.
"""
import struct

M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
C1 = 0xD1342543DE82EF95
C2 = 0xA0761D6478BD642F
NULL_ENTITY = 0xFFFFFFFF


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def bits_f64(u):
    return struct.unpack("<d", struct.pack("<Q", u & M64))[0]


def mix64(x):
    x &= M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def rng_key(seed, entity, commit_idx):
    k = mix64(seed ^ (((entity + 1) * C1) & M64))
    return mix64((k + commit_idx * C2) & M64)


def rng_draw(key, i):
    return mix64((key + (i + 1) * GOLD) & M64)


def u01(x):
    return float(x >> 11) * 1.1102230246251565e-16


def u01_open(x):
    return (float(x >> 12) + 0.5) * 2.220446049250313e-16


def log(x):
    ln2_hi = 6.93147180369123816490e-01
    ln2_lo = 1.90821492927058770002e-10
    L1, L2, L3, L4, L5, L6, L7 = (
        6.666666666666735130e-01, 3.999999999940941908e-01,
        2.857142874366239149e-01, 2.222219843214978396e-01,
        1.818357216161805012e-01, 1.531383769920937332e-01,
        1.479819860511658591e-01)
    ix = f64_bits(x)
    k = (ix >> 52) - 1023
    man = ix & 0x000FFFFFFFFFFFFF
    if man >= 0x6A09E667F3BCD:
        k += 1
        ix = man | 0x3FE0000000000000
    else:
        ix = man | 0x3FF0000000000000
    m = bits_f64(ix)
    f = m - 1.0
    dk = float(k)
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (L2 + w * (L4 + w * L6))
    t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)))
    R = t2 + t1
    hfsq = (0.5 * f) * f
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f)


def exponential(draw, lam):
    e = -log(u01_open(draw))
    return e if lam == 1.0 else e / lam


def trace_step(h, time, handler, src, data):
    x = mix64((f64_bits(time) + GOLD * ((handler << 32) | src)) & M64)
    x = mix64(x ^ data)
    h = ((h ^ x) * GOLD) & M64
    return h ^ (h >> 29)


def digest_term(entity, count, hsh):
    return (mix64(hsh ^ mix64((entity << 1) | 1)) +
            mix64((count + C1 * (entity + 1)) & M64)) & M64


PHOLD_UNIFORM, PHOLD_ROSS_REMOTE, PHOLD_OTHER_GROUP = 0, 1, 2


def phold_generate(seed, me, commit_idx, n_entities, p_remote, lam, lookahead,
                   mode, groups, first_id=0):
    """-> (target num, offset); transcription of simian_phold_generate."""
    key = rng_key(seed, first_id + me, commit_idx)
    target = me
    if mode == PHOLD_UNIFORM:
        target = int(u01(rng_draw(key, 1)) * float(n_entities))
    elif u01(rng_draw(key, 0)) < p_remote:
        if mode == PHOLD_ROSS_REMOTE:
            target = int(u01(rng_draw(key, 1)) * float(n_entities))
        else:
            g = groups
            per = n_entities // g
            k = int(u01(rng_draw(key, 1)) * float(per * (g - 1)))
            hop = 1 + k % (g - 1)
            row = k // (g - 1)
            target = row * g + (me % g + hop) % g
    offset = exponential(rng_draw(key, 2), lam) + lookahead
    return target, offset
