"""Seeded synthetic AIS stimulus (SURVEY.md 8d "Synthetic input definition").

GMSK (BT 0.4, h 0.5, 9600 Bd) bursts built from AIS payloads with HDLC framing
(24-bit 0101 preamble, 0x7E, LSB-first bytes, CRC-16/X.25 FCS, bit stuffing,
0x7E), NRZI (toggle on 0), placed at -25 kHz (channel A) / +25 kHz (channel B)
plus a small carrier offset, over complex AWGN.  Used by tests/ and bench.py
for input data only -- no part of the demodulator lives here.

The framing mirrors what the reference's bit decoder accepts
(reference Source/Marine/AIS.h:91-181, AIS.cpp:55-64) and the NMEA armouring
of Source/Marine/Message.cpp:569-686 (re-implemented independently below so
tests have a third opinion besides the oracle and the CUDA path).
"""
import numpy as np

SIXBIT = "0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVW`abcdefghijklmnopqrstuvw"
SAMPLE_A = "15MgK45P3@G?fl0E`JbR0OwT0@MS"  # reference python/tests/test_decode.py:12
SAMPLE_B = "15NPOOPP00o?b=bE`UNv4?w428D;"  # a second valid type-1 payload
CI_LINE = "13u?etPv2;0n:dDPwUM1U1Cb069D"   # type-1 payload shape used in the reference CI sample


def payload_to_bits(payload, fill=0):
    """6-bit de-armour an NMEA payload string -> list of bits (MSB first per letter)."""
    bits = []
    for c in payload:
        v = ord(c) - 48
        if v > 40:
            v -= 8
        bits.extend([(v >> (5 - k)) & 1 for k in range(6)])
    if fill:
        bits = bits[:-fill]
    return np.array(bits, dtype=np.uint8)


def bits_to_payload(bits):
    """bits (message order, MSB-first fields) -> (payload string, fill)."""
    n = len(bits)
    nl = (n + 5) // 6
    pad = np.zeros(nl * 6, dtype=np.uint8)
    pad[:n] = bits
    out = []
    for i in range(nl):
        v = 0
        for k in range(6):
            v = (v << 1) | int(pad[i * 6 + k])
        out.append(SIXBIT[v])
    return "".join(out), nl * 6 - n


def nmea_sentences(bits, channel, seq_start=0, own=False):
    """Independent !AIVDM builder (format of reference Message.cpp:569-631).

    Returns (list of sentences, next seq id)."""
    payload, fill = bits_to_payload(bits)
    nl = len(payload)
    ns = 1 if nl == 0 else (nl + 55) // 56
    seq = ""
    nxt = seq_start
    if ns > 1:
        seq = str(seq_start)
        nxt = (seq_start + 1) % 10
    out = []
    for s in range(ns):
        part = payload[s * 56:(s + 1) * 56]
        f = fill if s == ns - 1 else 0
        body = "AIVD%s,%d,%d,%s,%s,%s,%d" % ("O" if own else "M", ns, s + 1, seq, channel, part, f)
        c = 0
        for ch in body:
            c ^= ord(ch)
        out.append("!%s*%02X" % (body, c))
    return out, nxt


def crc16_x25(bits):
    crc = 0xFFFF
    for b in bits:
        if (int(b) ^ crc) & 1:
            crc = (crc >> 1) ^ 0x8408
        else:
            crc >>= 1
    return crc ^ 0xFFFF


def frame_bits(msg_bits, preamble=24, tail=8):
    """Message bits (MSB-first fields) -> transmitted HDLC bit sequence (before NRZI)."""
    msg_bits = np.asarray(msg_bits, dtype=np.uint8)
    nbytes = (len(msg_bits) + 7) // 8
    padded = np.zeros(nbytes * 8, dtype=np.uint8)
    padded[:len(msg_bits)] = msg_bits
    # each byte is sent LSB first
    tx = padded.reshape(nbytes, 8)[:, ::-1].reshape(-1)
    tx = tx[:len(tx)]  # whole bytes on air
    # NB: AIS messages are byte multiples on air for the standard types used here
    crc = crc16_x25(tx)
    fcs = np.array([(crc >> k) & 1 for k in range(16)], dtype=np.uint8)
    data = np.concatenate([tx, fcs])
    stuffed = []
    ones = 0
    for b in data:
        stuffed.append(int(b))
        if b:
            ones += 1
            if ones == 5:
                stuffed.append(0)
                ones = 0
        else:
            ones = 0
    flag = [0, 1, 1, 1, 1, 1, 1, 0]
    pre = [(k & 1) for k in range(preamble)]
    return np.array(pre + flag + stuffed + flag + [0] * tail, dtype=np.uint8)


def nrzi(bits, start=1):
    lvl = start
    out = np.empty(len(bits), dtype=np.float64)
    for i, b in enumerate(bits):
        if b == 0:
            lvl = -lvl
        out[i] = lvl
    return out


def gmsk_baseband(tx_bits, fs, timing_frac=0.0, bt=0.4, baud=9600.0, span=4):
    """NRZI + Gaussian-filtered MSK at sample rate fs; returns unit-amplitude complex64 burst."""
    lv = nrzi(tx_bits)
    sps = fs / baud
    n = int(np.ceil((len(lv) + span) * sps)) + 2
    t = (np.arange(n) + timing_frac) / sps  # symbol-time of each output sample
    idx = np.floor(t).astype(np.int64)
    idx = np.clip(idx, 0, len(lv) - 1)
    rect = lv[idx]
    rect[t >= len(lv)] = 0.0
    # gaussian pulse shaping filter sampled at fs
    sigma = np.sqrt(np.log(2.0)) / (2 * np.pi * bt)  # in symbol periods
    half = int(np.ceil(span * sps / 2))
    k = np.arange(-half, half + 1) / sps
    g = np.exp(-0.5 * (k / sigma) ** 2)
    g /= g.sum()
    f = np.convolve(rect, g, mode="same")
    phase = np.cumsum(f) * (np.pi / 2.0) / sps
    return np.exp(1j * phase)


def random_message_bits(rng, msg_type=None, nbits=168):
    """A structurally valid random single-sentence message (types 1/2/3/18, 168 bits)."""
    if msg_type is None:
        msg_type = int(rng.choice([1, 2, 3, 18]))
    bits = rng.integers(0, 2, nbits).astype(np.uint8)
    def put(start, ln, val):
        for k in range(ln):
            bits[start + k] = (val >> (ln - 1 - k)) & 1
    put(0, 6, msg_type)
    put(6, 2, int(rng.integers(0, 4)))
    put(8, 30, int(rng.integers(200000000, 780000000)))
    return bits


def type5_like_bits(rng):
    """424-bit type-5 shaped message (2 NMEA sentences) for the multi-sentence test."""
    bits = random_message_bits(rng, msg_type=1, nbits=424)
    for k in range(6):
        bits[k] = (5 >> (5 - k)) & 1
    return bits


class Burst:
    __slots__ = ("start", "channel", "bits", "amp", "foffs", "timing")

    def __init__(self, start, channel, bits, amp=0.3, foffs=0.0, timing=0.0):
        self.start = int(start)
        self.channel = channel
        self.bits = np.asarray(bits, dtype=np.uint8)
        self.amp = float(amp)
        self.foffs = float(foffs)
        self.timing = float(timing)


def render_stream(fs, n_samples, bursts, noise_sigma=0.02, seed=0):
    """complex64[n_samples]: AWGN + the given bursts mixed to -/+25 kHz (A/B)."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples)) * noise_sigma
    for b in bursts:
        bb = gmsk_baseband(frame_bits(b.bits), fs, timing_frac=b.timing)
        fc = (-25000.0 if b.channel == "A" else 25000.0) + b.foffs
        n0 = b.start
        n1 = min(n_samples, n0 + len(bb))
        if n1 <= n0:
            continue
        k = np.arange(n0, n1)
        ph0 = rng.uniform(0, 2 * np.pi)
        x[n0:n1] += b.amp * bb[:n1 - n0] * np.exp(1j * (2 * np.pi * fc * k / fs + ph0))
    return x.astype(np.complex64)


def burst_len_samples(nbits_msg, fs):
    # preamble 24 + flag 8 + data + fcs 16 + worst-case stuffing + flag 8 + tail
    return int((24 + 8 + nbits_msg + 16 + nbits_msg // 5 + 8 + 16) * fs / 9600.0) + 64


def random_stream(fs, n_samples, stream_id, bursts_per_sec=(2, 8), noise_sigma=0.02, base_seed=0xA15CA7,
                  multi_sentence=False):
    """SURVEY.md 8d generator: returns (complex64 samples, list of Burst) for one stream."""
    rng = np.random.default_rng(base_seed + stream_id)
    dur = n_samples / fs
    bursts = []
    for ch in "AB":
        k = int(rng.integers(bursts_per_sec[0], bursts_per_sec[1] + 1) * dur + 0.999)
        t = 0
        for _ in range(k):
            bits = type5_like_bits(rng) if (multi_sentence and rng.random() < 0.3) else random_message_bits(rng)
            ln = burst_len_samples(len(bits), fs)
            gap = int(rng.integers(ln // 8, max(ln // 8 + 1, (n_samples // max(k, 1)) - ln)))
            start = t + gap
            if start + ln >= n_samples:
                break
            bursts.append(Burst(start, ch, bits, amp=rng.uniform(0.05, 0.6), foffs=rng.uniform(-800, 800),
                                timing=rng.uniform(0, 1)))
            t = start + ln
    x = render_stream(fs, n_samples, bursts, noise_sigma=noise_sigma, seed=base_seed * 7 + stream_id)
    return x, bursts


def to_cu8(x):
    v = np.empty(2 * len(x), dtype=np.float32)
    v[0::2] = x.real
    v[1::2] = x.imag
    return np.clip(np.round(127.0 * v + 128.0), 0, 255).astype(np.uint8)


def fuzz_stream(fs, n_samples, seed, noise_sigma=0.02):
    """Decoder stress stimulus: densely packed bursts of random length (40..1064 bits) whose payloads are biased towards long
    runs of ones (heavy bit stuffing), with random -- also invalid -- type fields, amplitudes from below the noise to strong
    (some frames fail their CRC, noise-only stretches raise false start flags), short gaps, occasional collisions on a
    channel and truncated preambles.  Returns (complex64 samples, number of bursts).  Truth is whatever the oracle decodes."""
    rng = np.random.default_rng(0xF0221 + seed)
    bursts = []
    for ch in "AB":
        t = int(rng.integers(0, fs // 50))
        while True:
            nbits = 8 * int(rng.integers(5, 134))  # 40 .. 1064
            r = rng.random()
            ty = -1
            if r < 0.6:  # a (type, length) pair Decoder::cannotBeValid / Message::validate let through (AIS.cpp:111-142, Message.cpp:398-413)
                fixed = [(1, 168), (2, 168), (3, 168), (4, 168), (18, 168), (5, 424), (19, 312), (21, 360), (24, 160), (27, 96), (9, 168), (11, 168)]
                if rng.random() < 0.5:
                    ty, nbits = fixed[int(rng.integers(0, len(fixed)))]
                else:
                    ty = int(rng.choice([6, 8, 12, 14, 17, 26]))
                    nbits = 8 * int(rng.integers(12, 134))
            elif r < 0.7:  # type 0 or > 28: cannotBeValid fires at bit 30
                ty = int(rng.choice([0, 29, 40, 63]))
            p1 = float(rng.choice([0.5, 0.5, 0.7, 0.85, 0.95]))
            bits = (rng.random(nbits) < p1).astype(np.uint8)
            if ty >= 0:
                for k in range(6):
                    bits[k] = (ty >> (5 - k)) & 1
                if nbits >= 38 and rng.random() < 0.9:
                    mm = int(rng.integers(1, 999999999))
                    for k in range(30):
                        bits[8 + k] = (mm >> (29 - k)) & 1
            ln = burst_len_samples(nbits, fs)
            if t + ln >= n_samples:
                break
            amp = float(rng.choice([0.012, 0.02, 0.03, 0.05, 0.1, 0.2, 0.3, 0.4]))
            b = Burst(t, ch, bits, amp=amp, foffs=rng.uniform(-600, 600), timing=rng.uniform(0, 1))
            bursts.append(b)
            g = rng.random()
            if g < 0.15:
                t += int(ln * rng.uniform(0.3, 0.9))  # the next burst collides with the tail of this one
            elif g < 0.6:
                t += ln - int(rng.integers(0, 24) * fs / 9600.0)  # back to back, tails may touch the next preamble
            else:
                t += ln + int(rng.integers(0, fs // 10))  # a stretch of noise
    x = render_stream(fs, n_samples, bursts, noise_sigma=noise_sigma, seed=0xF0221 * 3 + seed)
    return x, len(bursts)
