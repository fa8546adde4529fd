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


# ---- LANL PDES benchmark handlers: transcription of include/simian_models.h --
# (simian_lanl_construct / _send / _receive / _finish).  `ctx` supplies
#   now, draw(i) -> u64, req_service(offset, service_name, data, rx_num_or_None)
# `st` is any object with the fields of simian_lanl_state_t plus `list`.

CTOR_COMMIT = M64
LANL_TIME_BINS = 16


def sfloor(x):
    """simian_floor: Math.floor for |x| < 2^52 via truncation"""
    t = float(int(x))
    return t - 1.0 if t > x else t


def sceil(x):
    return -sfloor(-x)


def powi(b, n):
    """simian_powi: b**n as a fixed sequence of f64 multiplications"""
    r = 1.0
    while n:
        if n & 1:
            r = r * b
        b = b * b
        n >>= 1
    return r


class LanlParams:
    def __init__(self, n_ent, s_ent, q_avg, p_receive, m_ent, ops_ent, ops_sigma,
                 cache_friendliness, end_time, min_delay, time_bins,
                 p_send=0.0, p_list=0.0, invert=False, stats_on_path=False):
        import math
        self.p_send, self.p_list, self.invert = p_send, p_list, invert
        self.stats_on_path = stats_on_path
        self.n_ent, self.s_ent, self.q_avg, self.p_receive = n_ent, s_ent, q_avg, p_receive
        self.m_ent, self.ops_ent, self.ops_sigma = m_ent, ops_ent, ops_sigma
        self.cache_friendliness, self.end_time, self.min_delay = cache_friendliness, end_time, min_delay
        self.time_bins = time_bins
        self.r_min = math.pow(1.0 - p_receive, n_ent)


class _Di:
    """the running draw index `di` of a handler invocation"""
    def __init__(self):
        self.i = 0

    def next(self):
        i = self.i
        self.i += 1
        return i


def lanl_send_body(p, st, ctx, di):
    now, end_time = ctx.now, p.end_time
    st.send_count += 1
    st.q_size -= 1
    b = int(sfloor(now / (end_time + 0.0001) * float(p.time_bins)))
    if 0 <= b < LANL_TIME_BINS:
        st.time_sends[b] += 1
    while (float(st.q_size) < st.q_target) and not (st.last_scheduled > end_time):
        own_delay = exponential(ctx.draw(di.next()), 1.0 / st.local_intersend_delay)
        st.last_scheduled = st.last_scheduled + own_delay
        if st.last_scheduled < end_time:
            st.q_size += 1
            ctx.req_service(st.last_scheduled - now, "SendHandler", ctx.draw(di.next()), None)
    if p.p_receive == 1.0:
        dest = 0.0
    elif p.p_receive == 0:
        dest = sfloor(float(p.n_ent) * u01(ctx.draw(di.next())))
    else:
        U = (1.0 - p.r_min) * u01(ctx.draw(di.next())) + p.r_min
        if not (U > 0.0):
            U = 2.2250738585072014e-308
        dest = sfloor(sceil(log(U) / log(1.0 - p.p_receive))) - 1
        if dest < 0:
            dest = 0.0
        if dest > float(p.n_ent - 1):
            dest = float(p.n_ent - 1)
    new_seed = ctx.draw(di.next())
    if now + p.min_delay < end_time:
        ctx.req_service(p.min_delay, "ReceiveHandler", new_seed, int(dest))


def lanl_send(p, st, ctx):
    lanl_send_body(p, st, ctx, _Di())


def lanl_gauss(ctx, di, mu, sigma):
    import math
    while True:
        x1 = 2.0 * u01(ctx.draw(di.next())) - 1.0
        x2 = 2.0 * u01(ctx.draw(di.next())) - 1.0
        w = (x1 * x1) + (x2 * x2)
        if not (w >= 1.0 or w == 0.0):
            break
    w = math.sqrt((-2.0 * log(w)) / w)
    return ((x1 * w) * sigma) + mu


def lanl_receive(p, st, ctx):
    di = _Di()
    r_ops = sfloor(lanl_gauss(ctx, di, st.ops, st.ops * p.ops_sigma))
    if not (r_ops > 1.0):
        r_ops = 1.0
    r_active, r_ops_per_element = 1.0, 0.0
    if st.ops > 0.0:  # ops == 0 (p_list tail): no list work (simian_models.h)
        r_active = sfloor(float(st.active_elements) * (r_ops / st.ops))
        if r_active > float(st.list_size):
            r_active = float(st.list_size)
        if not (r_active > 1.0):
            r_active = 1.0
        r_ops_per_element = sfloor(r_ops / r_active)
    st.receive_count += 1
    if r_ops > st.ops_max:
        st.ops_max = r_ops
    if r_ops < st.ops_min:
        st.ops_min = r_ops
    st.ops_mean = (st.ops_mean * float(st.receive_count - 1) + r_ops) / float(st.receive_count)
    na, nj = int(r_active), int(r_ops_per_element)
    st.value_sink += lanl_weighted_sum(ctx, st.list, na, nj, di.i)
    st.ops_count += na * nj


def lanl_weighted_sum(ctx, lst, na, nj, di0):
    """simian_lanl_weighted_sum (include/simian_models.h): 32 partial sums in
    increasing term order, then the butterfly p[l] += p[l ^ o], o = 16 .. 1"""
    p = [0.0] * 32
    for k in range(na * nj):
        p[k & 31] += lst[k // nj] * u01(ctx.draw(di0 + k))
    o = 16
    while o:
        p = [p[l] + p[l ^ o] for l in range(32)]
        o >>= 1
    return p[0]


def lanl_construct(p, st, ctx):
    di = _Di()
    num = st.num
    prob = 1.0 / float(p.n_ent)
    if p.p_send != 0:
        prob = p.p_send * powi(1.0 - p.p_send, (p.n_ent - num) if p.invert else num)
    target_global_sends = float(p.n_ent) * p.s_ent
    target_sends = sfloor(prob * target_global_sends)
    st.local_intersend_delay = (p.end_time / target_sends) if target_sends > 0 else 10 * p.end_time
    prob = 1.0 / float(p.n_ent)
    if p.p_list != 0:
        prob = p.p_list * powi(1.0 - p.p_list, num)
    st.list_size = int(sfloor(prob * float(p.n_ent) * p.m_ent))
    st.ops = sfloor(prob * float(p.n_ent) * p.ops_ent)
    st.active_elements = int(sfloor(p.cache_friendliness * float(st.list_size)))
    st.list = [u01(ctx.draw(di.next())) for _ in range(st.list_size)]
    st.q_target = (p.q_avg / p.s_ent) * target_sends
    st.q_size = 1
    st.last_scheduled = ctx.now
    st.send_count = st.receive_count = st.ops_count = 0
    st.ops_max = 0.0
    st.ops_min = 1e100
    st.ops_mean = 0.0
    st.value_sink = 0.0
    st.time_sends = [0] * LANL_TIME_BINS
    st.stats_received = st.gsend_count = st.greceive_count = st.gops_count = 0
    st.gops_max = 0.0
    st.gops_min = 1e100
    st.gops_mean = 0.0
    st.gtime_sends = [0] * LANL_TIME_BINS
    ctx.req_service(p.end_time - ctx.now, "FinishHandler", 0, None)
    lanl_send_body(p, st, ctx, di)


# ---- hello.js handlers: transcription of simian_hello_* (simian_models.h) ------

def hello_generate(seed, gid, commit_idx, count):
    """-> (target class 0=Alice/1=Bob, target id, data f64)"""
    key = rng_key(seed, gid, commit_idx)
    target = int(u01(rng_draw(key, 0)) * 2.0)
    target_id = int(u01(rng_draw(key, 1)) * float(count))
    data = u01(rng_draw(key, 2)) * 100.0
    return target, target_id, data


# ---- payloads beyond 8 bytes: transcription of simian_blob_* (simian_models.h) --

BLOB_MODEL_MAX_WORDS = 24


def blob_send(seed, gid, commit_idx, n_entities, max_words, lookahead):
    """-> (target, words, delay to the receiver, own delay)"""
    key = rng_key(seed, gid, commit_idx)
    target = int(u01(rng_draw(key, 0)) * float(n_entities))
    mw = min(max_words, BLOB_MODEL_MAX_WORDS)
    nw = rng_draw(key, 1) % (mw + 1)
    words = [rng_draw(key, 4 + k) for k in range(nw)]
    return (target, words, exponential(rng_draw(key, 2), 1.0) + lookahead,
            exponential(rng_draw(key, 3), 1.0) + lookahead)


def blob_fold(words):
    fold = mix64((len(words) + 1) & M64)
    for w in words:
        fold = mix64(fold ^ w)
    return fold


def lanl_finish_msg(p, st):
    """simian_lanl_finish: the statistics message as its 64-bit words
    [num, send_count, receive_count, ops_min, ops_mean, ops_max, time_sends.., opsCount]"""
    bins = min(p.time_bins, LANL_TIME_BINS)
    return ([st.num, st.send_count, st.receive_count, f64_bits(st.ops_min),
             f64_bits(st.ops_mean), f64_bits(st.ops_max)] +
            [int(x) for x in st.time_sends[:bins]] + [st.ops_count])


def lanl_output(p, st, words):
    """simian_lanl_output: the accumulation of OutputHandler
    (pdes_lanl_benchmarkV8.js:321-330) on entity 0"""
    bins = min(p.time_bins, LANL_TIME_BINS)
    if len(words) != 7 + bins:
        return
    sends, recvs = words[1], words[2]
    ops_min, ops_mean, ops_max = bits_f64(words[3]), bits_f64(words[4]), bits_f64(words[5])
    st.stats_received += 1
    den = float(st.greceive_count + recvs)
    if not (den > 1.0):
        den = 1.0
    st.gops_mean = (float(recvs) * ops_mean + st.gops_mean * float(st.greceive_count)) / den
    st.gsend_count += sends
    st.greceive_count += recvs
    st.gops_count += words[6 + bins]
    if ops_min < st.gops_min:
        st.gops_min = ops_min
    if ops_max > st.gops_max:
        st.gops_max = ops_max
    for i in range(bins):
        st.gtime_sends[i] += words[6 + i]
