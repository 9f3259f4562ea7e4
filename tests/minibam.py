"""A tiny, independent BAM + FASTA writer for handcrafted pileup cases (tests only)."""
import struct
import zlib

_OPS = "MIDNSHP=X"
_NT = {"=": 0, "A": 1, "C": 2, "G": 4, "T": 8, "N": 15}


def _bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(comp) + 25)
    return head + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def _cigar(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append(int(num) << 4 | _OPS.index(ch))
            num = ""
    return out


def _tag(name, value):
    if isinstance(value, str):
        return name.encode() + b"Z" + value.encode() + b"\0"
    if 0 <= value < 256:
        return name.encode() + b"C" + struct.pack("<B", value)
    return name.encode() + b"i" + struct.pack("<i", value)


def write(bam_path, fasta_path, contigs, reads, read_groups=()):
    """contigs: [(name, sequence)]; reads: dicts with tid,pos,cigar,seq,qual and optional
    flag,mapq,tags (dict, insertion ordered), name.  Reads must already be coordinate sorted."""
    with open(fasta_path, "w") as f:
        for name, seq in contigs:
            f.write(">%s\n%s\n" % (name, seq))
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, len(s)) for n, s in contigs)
    text += "".join("@RG\tID:%s\tLB:%s\n" % (g, g) for g in read_groups)
    parts = [b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs))]   # joined once: many reads
    for name, seq in contigs:
        parts.append(struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", len(seq)))
    for i, r in enumerate(reads):
        name = r.get("name", "read%d" % i).encode() + b"\0"
        cig = _cigar(r["cigar"])
        seq = r["seq"]
        packed = bytearray((len(seq) + 1) // 2)
        for j, ch in enumerate(seq):
            packed[j >> 1] |= _NT[ch] << (4 if j % 2 == 0 else 0)
        body = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r.get("mapq", 42), 4680, len(cig), r.get("flag", 0),
                           len(seq), -1, -1, 0)
        body += name + b"".join(struct.pack("<I", c) for c in cig) + bytes(packed) + bytes(r["qual"])
        for k, v in r.get("tags", {}).items():
            body += _tag(k, v)
        parts.append(struct.pack("<i", len(body)) + body)
    raw = b"".join(parts)
    with open(bam_path, "wb") as f:
        for p in range(0, len(raw), 0xFF00):
            f.write(_bgzf_block(raw[p:p + 0xFF00]))
        f.write(_bgzf_block(b""))
