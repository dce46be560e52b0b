"""FFV1 (version 3, Golomb-Rice coder, RGB + alpha, 8 bit) as OpenCV's FFmpeg backend writes it: range coder, configuration
record, slice headers / footers, sample decoder and encoder in plain Python.

TEST INFRASTRUCTURE ONLY -- the checker for a future GPU codec; nothing in the
package imports it.  The algorithm is not in the reference tree: every result video of the reference goes through
`cv2.VideoWriter_fourcc(*"FFV1")` (stereo_rerender.py:420-442,941; depth_frames_helper.py:125-161; 3d_view_depthfile.py:
118-127), i.e. libavcodec's FFV1 encoder inside the unpinned `opencv-python` wheel (OpenCV 4.13.0 with avcodec 62.11.100
here).  This file restates the published algorithm (RFC 9043 / libavcodec ffv1enc.c, ffv1dec.c, rangecoder.c, golomb.h)
and is pinned against that library itself (tests/test_ffv1_oracle.py): the configuration record and whole frame packets
(key and non-key frames) come out BYTE-IDENTICAL to libavcodec's, OpenCV-written packets decode to the original frames,
and all-key-frame streams with up to hundreds of slices written here are decoded bit-exactly by OpenCV.

Stream facts established with it: version 3, micro_version 4, coder 0 (Golomb-Rice), colourspace RGB with the JPEG2000
RCT, 9-bit samples, an alpha plane (OpenCV feeds BGRA, alpha = 255), 2 x 2 slices, quant-table set 0 (3 inputs, 666
contexts), per-slice CRC-32 (polynomial 0x04C11DB7, MSB first, initial value 0), key frame every 12 frames (non-key
frames carry the adaptive VLC states over).  Frames are BGRA uint8 (H, W, 4).  Slow: use small frames.
"""
from __future__ import annotations

import numpy as np

def build_rac_states(factor=int(0.05 * (1 << 32)), max_p=256 - 8):
    one = 1 << 32
    one_state = [0] * 256; zero_state = [0] * 256
    last_p8 = 0; p = one // 2
    for i in range(128):
        p8 = (256 * p + one // 2) >> 32
        if p8 <= last_p8: p8 = last_p8 + 1
        if last_p8 and last_p8 < 256 and p8 <= max_p: one_state[last_p8] = p8
        p += ((one - p) * factor + one // 2) >> 32
        last_p8 = p8
    for i in range(256 - max_p, max_p + 1):
        if one_state[i]: continue
        p = (i * one + 128) >> 8
        p += ((one - p) * factor + one // 2) >> 32
        p8 = (256 * p + one // 2) >> 32
        if p8 <= i: p8 = i + 1
        if p8 > max_p: p8 = max_p
        one_state[i] = p8
    for i in range(1, 255):
        zero_state[i] = 256 - one_state[256 - i]
    return zero_state, one_state

class RangeDecoder:
    def __init__(self, buf, end=None):
        self.buf = buf; self.pos = 0; self.end = len(buf) if end is None else end
        self.zero_state, self.one_state = build_rac_states()
        self.range = 0xFF00
        self.low = (buf[0] << 8) | buf[1]; self.pos = 2
        self.overread = 0
        if self.low >= 0xFF00:
            self.low = 0xFF00; self.end = self.pos
    def refill(self):
        if self.range < 0x100:
            self.range <<= 8; self.low <<= 8
            if self.pos < self.end:
                self.low += self.buf[self.pos]; self.pos += 1
            else:
                self.overread += 1
    def get_rac(self, state, idx):
        range1 = (self.range * state[idx]) >> 8
        self.range -= range1
        if self.low < self.range:
            state[idx] = self.zero_state[state[idx]]; self.refill(); return 0
        self.low -= self.range; state[idx] = self.one_state[state[idx]]; self.range = range1; self.refill(); return 1
    def get_symbol(self, state, is_signed):
        if self.get_rac(state, 0): return 0
        e = 0
        while self.get_rac(state, 1 + min(e, 9)):
            e += 1
            if e > 31: raise ValueError("bad symbol")
        a = 1
        for i in range(e - 1, -1, -1):
            a += a + self.get_rac(state, 22 + min(i, 9))
        neg = is_signed and self.get_rac(state, 11 + min(e, 10))
        return -a if neg else a

def read_quant_table(c, scale):
    q = [0] * 256; state = [128] * 32; i = 0; v = 0
    while i < 128:
        ln = c.get_symbol(state, 0) + 1
        if ln > 128 - i: raise ValueError("bad quant table")
        for _ in range(ln):
            q[i] = scale * v; i += 1
        v += 1
    for k in range(1, 128): q[256 - k] = -q[k]
    q[128] = -q[127]
    return q, 2 * v - 1

def read_quant_tables(c):
    tables = []; ctx = 1
    for _ in range(5):
        q, n = read_quant_table(c, ctx); tables.append(q); ctx *= n
    return tables, (ctx + 1) // 2

def parse_config(extra):
    c = RangeDecoder(extra)
    st = [128] * 32
    cfg = {}
    cfg["version"] = c.get_symbol(st, 0)
    if cfg["version"] > 2:
        c.end -= 4
        cfg["micro_version"] = c.get_symbol(st, 0)
    cfg["ac"] = c.get_symbol(st, 0)
    if cfg["ac"] == 2:
        cfg["state_transition_delta"] = [c.get_symbol(st, 1) for _ in range(1, 256)]
    cfg["colorspace"] = c.get_symbol(st, 0)
    cfg["bits"] = c.get_symbol(st, 0)
    cfg["chroma_planes"] = c.get_rac(st, 0)
    cfg["chroma_h_shift"] = c.get_symbol(st, 0)
    cfg["chroma_v_shift"] = c.get_symbol(st, 0)
    cfg["transparency"] = c.get_rac(st, 0)
    cfg["num_h_slices"] = 1 + c.get_symbol(st, 0)
    cfg["num_v_slices"] = 1 + c.get_symbol(st, 0)
    cfg["quant_table_count"] = c.get_symbol(st, 0)
    cfg["quant_tables"] = []; cfg["context_count"] = []
    for _ in range(cfg["quant_table_count"]):
        t, n = read_quant_tables(c); cfg["quant_tables"].append(t); cfg["context_count"].append(n)
    cfg["states_coded"] = []
    for i in range(cfg["quant_table_count"]):
        cfg["states_coded"].append(c.get_rac(st, 0))
        if cfg["states_coded"][-1]: raise NotImplementedError("initial states")
    if cfg["version"] > 2:
        cfg["ec"] = c.get_symbol(st, 0)
        if cfg.get("micro_version", 0) > 2: cfg["intra"] = c.get_symbol(st, 0)
    return cfg


LOG2_RUN = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24]


class BitReader:
    def __init__(self, buf, pos_bytes):
        self.buf = buf; self.pos = pos_bytes * 8
    def get1(self):
        b = (self.buf[self.pos >> 3] >> (7 - (self.pos & 7))) & 1; self.pos += 1; return b
    def get(self, n):
        v = 0
        for _ in range(n): v = (v << 1) | self.get1()
        return v

class VlcState:
    __slots__ = ("error_sum", "drift", "bias", "count")
    def __init__(self): self.error_sum = 4; self.drift = 0; self.bias = 0; self.count = 1

def fold(v, bits):
    m = 1 << bits; v &= m - 1
    return v - m if v >= (m >> 1) else v

def get_ur_golomb(gb, k, limit, esc_len):
    """golomb.h get_ur_golomb (the FFV1 flavour): fewer than `limit` leading zeros -> zeros << k | k suffix bits;
    otherwise exactly `limit` zero bits, then esc_len bits holding value - limit + 1."""
    z = 0
    while z < limit:
        if gb.get1(): break
        z += 1
    else:
        return gb.get(esc_len) + limit - 1
    return (z << k) + (gb.get(k) if k else 0)

def get_vlc_symbol(gb, st, bits):
    i = st.count; k = 0
    while i < st.error_sum: k += 1; i += i
    v = get_ur_golomb(gb, k, 12, bits)
    v = (v >> 1) ^ -(v & 1)
    if (2 * st.drift + st.count) < 0: v = ~v          # v ^= (2*drift+count) >> 31
    ret = fold(v + st.bias, bits)
    # update
    drift = st.drift; count = st.count
    st.error_sum += abs(v); drift += v
    if count == 128: count >>= 1; drift >>= 1; st.error_sum >>= 1
    count += 1
    if drift <= -count:
        st.bias = max(st.bias - 1, -128); drift = max(drift + count, -count + 1)
    elif drift > 0:
        st.bias = min(st.bias + 1, 127); drift = min(drift - count, 0)
    st.drift = drift; st.count = count
    return ret

def mid_pred(a, b, c):
    return sorted((a, b, c))[1]

def decode_line(gb, w, cur, last, qt, states, bits, run_index):
    """cur / last: lists with 3 cells of left border and 3 of right border (index offset 3)."""
    run_count = 0; run_mode = 0
    O = 3
    x = 0
    while x < w:
        LT = last[O + x - 1]; T = last[O + x]; RT = last[O + x + 1]; L = cur[O + x - 1]
        context = qt[0][(L - LT) & 0xFF] + qt[1][(LT - T) & 0xFF] + qt[2][(T - RT) & 0xFF]
        sign = context < 0
        if sign: context = -context
        if context == 0 and run_mode == 0: run_mode = 1
        if run_mode:
            if run_count == 0 and run_mode == 1:
                if gb.get1():
                    run_count = 1 << LOG2_RUN[run_index]
                    if x + run_count <= w: run_index += 1
                else:
                    run_count = gb.get(LOG2_RUN[run_index]) if LOG2_RUN[run_index] else 0
                    if run_index: run_index -= 1
                    run_mode = 2
            run_count -= 1
            if run_count < 0:
                run_mode = 0; run_count = 0
                diff = get_vlc_symbol(gb, states[context], bits)
                if diff >= 0: diff += 1
            else:
                diff = 0
        else:
            diff = get_vlc_symbol(gb, states[context], bits)
        if sign: diff = -diff
        pred = mid_pred(L, L + T - LT, T)
        cur[O + x] = (pred + diff) & ((1 << bits) - 1)
        x += 1
    return run_index


def crc32_mpeg(data, crc=0):
    """AV_CRC_32_IEEE as libavutil computes it for FFV1: polynomial 0x04C11DB7, MSB first, no reflection, no final xor."""
    for b in data:
        crc ^= b << 24
        for _ in range(8):
            crc = ((crc << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if crc & 0x80000000 else (crc << 1) & 0xFFFFFFFF
    return crc

def slice_ranges(packet, n_slices, ec):
    trailer = 3 + 5 * (1 if ec else 0)
    p = len(packet); out = []
    for _ in range(n_slices):
        size = int.from_bytes(packet[p - trailer:p - trailer + 3], "big") + trailer
        out.append((p - size, p)); p -= size
    assert p == 0
    return out[::-1]

class SliceState:
    def __init__(self, cfg):
        self.states = None  # per plane-context list of VlcState lists
    def reset(self, cfg, qidx):
        self.states = [[VlcState() for _ in range(cfg["context_count"][qidx[pc]])] for pc in range(len(qidx))]

def decode_frame(packet, cfg, W, H, slice_states):
    nh, nv = cfg["num_h_slices"], cfg["num_v_slices"]
    ranges = slice_ranges(packet, nh * nv, cfg["ec"])
    trailer = 3 + 5 * (1 if cfg["ec"] else 0)
    out = np.zeros((H, W, 4), np.uint8)
    key = None
    for si, (a, b) in enumerate(ranges):
        buf = packet[a:b]
        if cfg["ec"]: assert crc32_mpeg(buf) == 0, "slice CRC"
        c = RangeDecoder(buf)
        if si == 0:
            ks = [128]; key = c.get_rac(ks, 0)
        st = [128] * 32
        sx = c.get_symbol(st, 0); sy = c.get_symbol(st, 0); sw = c.get_symbol(st, 0) + 1; sh = c.get_symbol(st, 0) + 1
        qidx = [c.get_symbol(st, 0) for _ in range(2 + cfg["transparency"])]
        c.get_symbol(st, 0); c.get_symbol(st, 0); c.get_symbol(st, 0)   # picture structure, sar num, sar den
        x0 = sx * W // nh; x1 = (sx + sw) * W // nh; y0 = sy * H // nv; y1 = (sy + sh) * H // nv
        c.get_rac([129], 0)
        gb = BitReader(buf, c.pos - 1)
        ss = slice_states[si]
        if key or ss.states is None: ss.reset(cfg, qidx)
        w, h = x1 - x0, y1 - y0
        nplanes = 3 + cfg["transparency"]
        bufs = [[[0] * (w + 6), [0] * (w + 6)] for _ in range(nplanes)]
        run_index = 0
        for y in range(h):
            for pl in range(nplanes):
                bufs[pl][0], bufs[pl][1] = bufs[pl][1], bufs[pl][0]
                last, cur = bufs[pl][0], bufs[pl][1]
                cur[2] = last[3]; last[3 + w] = last[3 + w - 1]
                pc = (pl + 1) // 2
                run_index = decode_line(gb, w, cur, last, cfg["quant_tables"][qidx[pc]], ss.states[pc], 9, run_index)
            g = np.array(bufs[0][1][3:3 + w]); bb = np.array(bufs[1][1][3:3 + w]) - 256; r = np.array(bufs[2][1][3:3 + w]) - 256
            a_ = np.array(bufs[3][1][3:3 + w]) if nplanes > 3 else np.full(w, 255)
            g = g - ((bb + r) >> 2); bb = bb + g; r = r + g
            out[y0 + y, x0:x1] = np.stack([bb & 255, g & 255, r & 255, a_ & 255], axis=-1)
    return out, key


class RangeEncoder:
    def __init__(self):
        self.zero_state, self.one_state = build_rac_states()
        self.low = 0; self.range = 0xFF00; self.outstanding_count = 0; self.outstanding_byte = -1
        self.out = bytearray()
    def renorm(self):
        if self.outstanding_byte < 0:
            self.outstanding_byte = self.low >> 8
        elif self.low <= 0xFF00:
            self.out.append(self.outstanding_byte)
            self.out.extend(b"\xff" * self.outstanding_count); self.outstanding_count = 0
            self.outstanding_byte = self.low >> 8
        elif self.low >= 0x10000:
            self.out.append(self.outstanding_byte + 1)
            self.out.extend(b"\x00" * self.outstanding_count); self.outstanding_count = 0
            self.outstanding_byte = (self.low >> 8) & 0xFF
        else:
            self.outstanding_count += 1
        self.low = (self.low & 0xFF) << 8
        self.range <<= 8
    def put_rac(self, state, idx, bit):
        range1 = (self.range * state[idx]) >> 8
        if not bit:
            self.range -= range1; state[idx] = self.zero_state[state[idx]]
        else:
            self.low += self.range - range1; self.range = range1; state[idx] = self.one_state[state[idx]]
        while self.range < 0x100: self.renorm()
    def put_symbol(self, state, v, is_signed):
        if v:
            a = abs(v); e = a.bit_length() - 1
            self.put_rac(state, 0, 0)
            if e <= 9:
                for i in range(e): self.put_rac(state, 1 + i, 1)
                self.put_rac(state, 1 + e, 0)
                for i in range(e - 1, -1, -1): self.put_rac(state, 22 + i, (a >> i) & 1)
                if is_signed: self.put_rac(state, 11 + e, v < 0)
            else:
                for i in range(e): self.put_rac(state, 1 + min(i, 9), 1)
                self.put_rac(state, 1 + 9, 0)
                for i in range(e - 1, -1, -1): self.put_rac(state, 22 + min(i, 9), (a >> i) & 1)
                if is_signed: self.put_rac(state, 11 + 10, v < 0)
        else:
            self.put_rac(state, 0, 1)
    def terminate(self, sentinel=True):
        if sentinel: self.put_rac([129], 0, 0)
        self.range = 0xFF; self.low += 0xFF; self.renorm(); self.range = 0xFF; self.renorm()
        return bytes(self.out)

class BitWriter:
    def __init__(self): self.acc = 0; self.n = 0; self.out = bytearray()
    def put(self, nbits, v):
        if nbits == 0: return
        self.acc = (self.acc << nbits) | (v & ((1 << nbits) - 1)); self.n += nbits
        while self.n >= 8:
            self.n -= 8; self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1
    def flush(self):
        if self.n: self.out.append((self.acc << (8 - self.n)) & 0xFF); self.n = 0; self.acc = 0
        return bytes(self.out)

def set_ur_golomb(bw, i, k, limit, esc_len):
    e = i >> k
    if e < limit: bw.put(e + k + 1, (1 << k) + (i & ((1 << k) - 1)))
    else: bw.put(limit + esc_len, i - limit + 1)

def put_vlc_symbol(bw, st, v, bits):
    v = fold(v - st.bias, bits)
    i = st.count; k = 0
    while i < st.error_sum: k += 1; i += i
    code = ~v if (2 * st.drift + st.count) < 0 else v
    u = -2 * code - 1
    if u < 0: u = ~u
    set_ur_golomb(bw, u, k, 12, bits)
    drift = st.drift; count = st.count
    st.error_sum += abs(v); drift += v
    if count == 128: count >>= 1; drift >>= 1; st.error_sum >>= 1
    count += 1
    if drift <= -count: st.bias = max(st.bias - 1, -128); drift = max(drift + count, -count + 1)
    elif drift > 0: st.bias = min(st.bias + 1, 127); drift = min(drift - count, 0)
    st.drift = drift; st.count = count

def encode_line(bw, w, cur, last, qt, states, bits, run_index):
    run_count = 0; run_mode = 0
    for x in range(w):
        LT = last[3 + x - 1]; T = last[3 + x]; RT = last[3 + x + 1]; L = cur[3 + x - 1]
        ctx = qt[0][(L - LT) & 255] + qt[1][(LT - T) & 255] + qt[2][(T - RT) & 255]
        diff = cur[3 + x] - mid_pred(L, L + T - LT, T)
        if ctx < 0: ctx = -ctx; diff = -diff
        diff = fold(diff, bits)
        if ctx == 0: run_mode = 1
        if run_mode:
            if diff:
                while run_count >= 1 << LOG2_RUN[run_index]:
                    run_count -= 1 << LOG2_RUN[run_index]; run_index += 1; bw.put(1, 1)
                bw.put(1 + LOG2_RUN[run_index], run_count)
                if run_index: run_index -= 1
                run_count = 0; run_mode = 0
                if diff > 0: diff -= 1
            else:
                run_count += 1
        if run_mode == 0: put_vlc_symbol(bw, states[ctx], diff, bits)
    if run_mode:
        while run_count >= 1 << LOG2_RUN[run_index]:
            run_count -= 1 << LOG2_RUN[run_index]; run_index += 1; bw.put(1, 1)
        if run_count: bw.put(1, 1)
    return run_index

def encode_frame(bgra, cfg, key, slice_states):
    H, W = bgra.shape[:2]
    nh, nv = cfg["num_h_slices"], cfg["num_v_slices"]
    packet = bytearray()
    si = 0
    for sy in range(nv):
        for sx in range(nh):
            x0 = sx * W // nh; x1 = (sx + 1) * W // nh; y0 = sy * H // nv; y1 = (sy + 1) * H // nv
            rc = RangeEncoder()
            if si == 0: rc.put_rac([128], 0, 1 if key else 0)
            st = [128] * 32
            for v in (sx, sy, 0, 0): rc.put_symbol(st, v, 0)
            qidx = [0] * (2 + cfg["transparency"])
            for q in qidx: rc.put_symbol(st, q, 0)
            rc.put_symbol(st, 3, 0); rc.put_symbol(st, 0, 0); rc.put_symbol(st, 1, 0)   # progressive, sar 0/1 as OpenCV's files have it
            head = rc.terminate(True)
            ss = slice_states[si]
            if key or ss.states is None: ss.reset(cfg, qidx)
            w, h = x1 - x0, y1 - y0
            nplanes = 3 + cfg["transparency"]
            bufs = [[[0] * (w + 6), [0] * (w + 6)] for _ in range(nplanes)]
            bw = BitWriter(); run_index = 0
            for y in range(h):
                row = bgra[y0 + y, x0:x1].astype(int)
                b = row[:, 0] - row[:, 1]; r = row[:, 2] - row[:, 1]; g = row[:, 1] + ((b + r) >> 2); b += 256; r += 256
                for pl, vals in enumerate((g, b, r, row[:, 3])[:nplanes]):
                    bufs[pl][0], bufs[pl][1] = bufs[pl][1], bufs[pl][0]
                    last, cur = bufs[pl][0], bufs[pl][1]
                    cur[3:3 + w] = [int(v) for v in vals]
                    cur[2] = last[3]; last[3 + w] = last[3 + w - 1]
                    pc = (pl + 1) // 2
                    run_index = encode_line(bw, w, cur, last, cfg["quant_tables"][qidx[pc]], ss.states[pc], 9, run_index)
            body = head + bw.flush()
            foot = len(body).to_bytes(3, "big")
            if cfg["ec"]:
                foot += b"\x00"
                foot += crc32_mpeg(body + foot).to_bytes(4, "big")
            packet += body + foot
            si += 1
    return bytes(packet)


def write_quant_table(rc, q):
    """Mirror of read_quant_table: run lengths of equal values over q[0..127]."""
    state = [128] * 32
    i = 0
    while i < 128:
        j = i
        while j < 128 and q[j] == q[i]: j += 1
        rc.put_symbol(state, j - i - 1, 0)
        i = j

def write_config(cfg, num_h_slices, num_v_slices):
    rc = RangeEncoder(); st = [128] * 32
    rc.put_symbol(st, cfg["version"], 0)
    rc.put_symbol(st, cfg["micro_version"], 0)
    rc.put_symbol(st, cfg["ac"], 0)
    rc.put_symbol(st, cfg["colorspace"], 0)
    rc.put_symbol(st, cfg["bits"], 0)
    rc.put_rac(st, 0, cfg["chroma_planes"])
    rc.put_symbol(st, cfg["chroma_h_shift"], 0)
    rc.put_symbol(st, cfg["chroma_v_shift"], 0)
    rc.put_rac(st, 0, cfg["transparency"])
    rc.put_symbol(st, num_h_slices - 1, 0)
    rc.put_symbol(st, num_v_slices - 1, 0)
    rc.put_symbol(st, cfg["quant_table_count"], 0)
    for tables in cfg["quant_tables"]:
        scale = 1
        for q in tables:
            write_quant_table(rc, [v // scale if scale else v for v in q[:128]])
            n = 2 * len(set(q[:128])) - 1
            scale *= n
    for _ in range(cfg["quant_table_count"]): rc.put_rac(st, 0, 0)
    rc.put_symbol(st, cfg["ec"], 0)
    rc.put_symbol(st, cfg["intra"], 0)
    body = rc.terminate(False)
    return body + crc32_mpeg(body).to_bytes(4, "big")


