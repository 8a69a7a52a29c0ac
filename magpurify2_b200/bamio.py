"""Minimal BGZF/BAM WRITER for synthetic alignments (tests and the file-based bench slice).
Not on the product path: the product only READS BAM (csrc/mp2_bam.cu)."""
import struct
import zlib

import numpy as np

_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
CIGAR_OPS = "MIDNSHP=X"


def _bgzf_block(data, level=1):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
            + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bgzf(fout, payload, level=1, block=0xFF00):
    mv = memoryview(payload)
    for o in range(0, len(mv), block):
        fout.write(_bgzf_block(bytes(mv[o:o + block]), level))
    fout.write(_EOF)


def header_bytes(names, lengths, sort_order="coordinate", extra_text=""):
    text = f"@HD\tVN:1.6\tSO:{sort_order}\n" + "".join(
        f"@SQ\tSN:{n}\tLN:{int(l)}\n" for n, l in zip(names, lengths)) + extra_text
    tb = text.encode()
    out = [b"BAM\x01", struct.pack("<i", len(tb)), tb, struct.pack("<i", len(names))]
    for n, l in zip(names, lengths):
        nb = n.encode() + b"\x00"
        out += [struct.pack("<i", len(nb)), nb, struct.pack("<i", int(l))]
    return b"".join(out)


def record_bytes(tid, pos, cigar, flag=0, nm=0, name="r", seq_len=None, mapq=60, nm_type="C", extra_aux=b""):
    """cigar: list of (op_char, length).  nm=None omits the NM tag."""
    cig = [(CIGAR_OPS.index(op), ln) for op, ln in cigar]
    if seq_len is None:
        seq_len = sum(ln for op, ln in cig if op in (0, 1, 4, 7, 8))
    nb = name.encode() + b"\x00"
    ref_len = sum(ln for op, ln in cig if op in (0, 2, 3, 7, 8))
    end = pos + ref_len
    bin_ = 4680 + (pos >> 14) if pos >= 0 and (pos >> 14) == ((end - 1) >> 14) else 0
    aux = extra_aux
    if nm is not None:
        fmt = {"C": "<B", "c": "<b", "S": "<H", "s": "<h", "I": "<I", "i": "<i"}[nm_type]
        aux = aux + b"NM" + nm_type.encode() + struct.pack(fmt, nm)
    body = (struct.pack("<iiBBHHHIiii", tid, pos, len(nb), mapq, bin_, len(cig), flag, seq_len, -1, -1, 0)
            + nb + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cig)
            + b"\x11" * ((seq_len + 1) // 2) + b"\xff" * seq_len + aux)
    return struct.pack("<i", len(body)) + body


def write_bam(path, names, lengths, records, sort_order="coordinate", level=1):
    """records: iterable of dicts / tuples accepted by record_bytes(**r), already in BAM order."""
    payload = bytearray(header_bytes(names, lengths, sort_order))
    for r in records:
        payload += record_bytes(**r)
    with open(path, "wb") as f:
        write_bgzf(f, payload, level)


def write_bam_fixed(path, names, lengths, tid, pos, nm, flag=None, read_len=150, level=1):
    """Vectorised writer for the headline synthetic shape: every record is `read_len`M with an
    NM:C tag; tid/pos/nm (and optionally flag) are NumPy arrays in BAM order."""
    n = len(tid)
    name = b"r\x00"
    l_seq = read_len
    body_len = 32 + len(name) + 4 + (l_seq + 1) // 2 + l_seq + 4
    rec = np.zeros((n, 4 + body_len), dtype=np.uint8)
    v = rec.view()
    def put(col, arr, dt):
        a = np.ascontiguousarray(arr, dtype=dt).view(np.uint8).reshape(n, -1)
        v[:, col:col + a.shape[1]] = a
    put(0, np.full(n, body_len), "<i4")
    put(4, tid, "<i4")
    put(8, pos, "<i4")
    v[:, 12] = len(name)
    v[:, 13] = 60
    put(14, np.zeros(n), "<u2")          # bin (unused by readers that scan sequentially)
    put(16, np.ones(n), "<u2")           # n_cigar_op
    put(18, np.zeros(n) if flag is None else flag, "<u2")
    put(20, np.full(n, l_seq), "<i4")
    put(24, np.full(n, -1), "<i4")
    put(28, np.full(n, -1), "<i4")
    put(32, np.zeros(n), "<i4")
    o = 36
    v[:, o:o + len(name)] = np.frombuffer(name, np.uint8)
    o += len(name)
    put(o, np.full(n, (read_len << 4) | 0), "<u4")
    o += 4
    v[:, o:o + (l_seq + 1) // 2] = 0x11
    o += (l_seq + 1) // 2
    v[:, o:o + l_seq] = 0xFF
    o += l_seq
    v[:, o:o + 3] = np.frombuffer(b"NMC", np.uint8)
    v[:, o + 3] = np.minimum(np.asarray(nm), 255).astype(np.uint8)
    with open(path, "wb") as f:
        payload = header_bytes(names, lengths) + rec.tobytes()
        write_bgzf(f, payload, level)
