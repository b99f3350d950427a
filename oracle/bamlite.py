"""Minimal BAM reader / writer in pure Python (zlib only).  TEST INFRASTRUCTURE ONLY.

Used by the samtools stand-in (oracle/standin_samtools.py) to serve real BAM files to the
UNMODIFIED reference, by the tests as the independent restatement of what `samtools view -L`,
`view -f4`, `view -N` and `fastq` select and print, and to write synthetic BAMs.  Format: SAM/BAM
specification v1 section 4 (BGZF: RFC 1952 members with a BC extra subfield)."""
import struct
import zlib

SEQ_CODE = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"
_COMP = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn=", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn=")


def revcomp(s):
    return s.translate(_COMP)[::-1]


def bgzf_inflate(data):
    out = []
    while data:
        d = zlib.decompressobj(31)
        out.append(d.decompress(data))
        data = d.unused_data
    return b"".join(out)


def _bgzf_block(raw, level=6):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = c.compress(raw) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
            + comp + struct.pack("<II", zlib.crc32(raw) & 0xffffffff, len(raw)))


def bgzf_deflate(raw, block=0xff00, level=6):
    out = [_bgzf_block(raw[i:i + block], level) for i in range(0, len(raw), block)]
    out.append(_bgzf_block(b""))  # EOF marker
    return b"".join(out)


def read_bam(path):
    """-> (header_text, [(ref_name, ref_len)], [record dict])"""
    b = bgzf_inflate(open(path, "rb").read())
    assert b[:4] == b"BAM\x01", "not a BAM file"
    l_text, = struct.unpack_from("<I", b, 4)
    text = b[8:8 + l_text].rstrip(b"\0").decode()
    p = 8 + l_text
    n_ref, = struct.unpack_from("<I", b, p)
    p += 4
    refs = []
    for _ in range(n_ref):
        l, = struct.unpack_from("<I", b, p)
        name = b[p + 4:p + 4 + l - 1].decode()
        ln, = struct.unpack_from("<I", b, p + 4 + l)
        refs.append((name, ln))
        p += 8 + l
    recs = []
    while p < len(b):
        bs, = struct.unpack_from("<I", b, p)
        ref, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", b, p + 4)
        q = p + 36
        name = b[q:q + l_name - 1].decode()
        q += l_name
        cigar = []
        for i in range(n_cig):
            v, = struct.unpack_from("<I", b, q + 4 * i)
            cigar.append((CIGAR_OPS[v & 15], v >> 4))
        q += 4 * n_cig
        sq = b[q:q + (l_seq + 1) // 2]
        seq = "".join(SEQ_CODE[(sq[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        q += (l_seq + 1) // 2
        qual = b[q:q + l_seq]
        if l_seq and qual[0] == 0xff:
            qual = None
        recs.append(dict(name=name, flag=flag, ref=ref, pos=pos, mapq=mapq, cigar=cigar, next_ref=nref,
                         next_pos=npos, tlen=tlen, seq=seq, qual=qual))
        p += 4 + bs
    return text, refs, recs


def _reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


def end_pos(rec):
    """bam_endpos(): pos + reference-consuming CIGAR length, pos + 1 when there is none"""
    n = sum(l for op, l in rec["cigar"] if op in "MDN=X")
    return rec["pos"] + (n if n else 1)


def write_bam(path, header_text, refs, recs, block=0xff00):
    out = [b"BAM\x01", struct.pack("<I", len(header_text)), header_text.encode(), struct.pack("<I", len(refs))]
    for name, ln in refs:
        out.append(struct.pack("<I", len(name) + 1) + name.encode() + b"\0" + struct.pack("<I", ln))
    code = {c: i for i, c in enumerate(SEQ_CODE)}
    for r in recs:
        name = r["name"].encode() + b"\0"
        seq = r["seq"]
        l_seq = len(seq)
        packed = bytearray((l_seq + 1) // 2)
        for i, c in enumerate(seq):
            packed[i >> 1] |= code.get(c.upper(), 15) << (0 if i & 1 else 4)
        qual = r.get("qual")
        qual = bytes(qual) if qual is not None else b"\xff" * l_seq
        cig = b"".join(struct.pack("<I", (l << 4) | CIGAR_OPS.index(op)) for op, l in r["cigar"])
        body = struct.pack("<iiBBHHHIiii", r["ref"], r["pos"], len(name), r.get("mapq", 0),
                           _reg2bin(max(r["pos"], 0), max(end_pos(r), 1)) if r["ref"] >= 0 else 4680,
                           len(r["cigar"]), r["flag"], l_seq, r.get("next_ref", -1), r.get("next_pos", -1),
                           r.get("tlen", 0)) + name + cig + bytes(packed) + qual
        out.append(struct.pack("<I", len(body)) + body)
    with open(path, "wb") as f:
        f.write(bgzf_deflate(b"".join(out), block))


def read_bed(path):
    """{chrom: [(beg, end)]}; `samtools view -L` semantics (0-based half-open)"""
    out = {}
    for line in open(path):
        if line.startswith(("#", "track", "browser")):
            continue
        t = line.split()
        if len(t) >= 3:
            out.setdefault(t[0], []).append((int(t[1]), int(t[2])))
    return out


def in_regions(rec, refs, bed):
    if rec["ref"] < 0:
        return False
    e = end_pos(rec)
    return any(b < e and rec["pos"] < en for b, en in bed.get(refs[rec["ref"]][0], ()))


def sam_line(rec, refs):
    rn = refs[rec["ref"]][0] if rec["ref"] >= 0 else "*"
    nx = "*" if rec["next_ref"] < 0 else ("=" if rec["next_ref"] == rec["ref"] else refs[rec["next_ref"]][0])
    cg = "".join(f"{l}{op}" for op, l in rec["cigar"]) or "*"
    ql = "*" if rec["qual"] is None else "".join(chr(33 + v) for v in rec["qual"])
    return "\t".join([rec["name"], str(rec["flag"]), rn, str(rec["pos"] + 1), str(rec["mapq"]), cg, nx,
                      str(rec["next_pos"] + 1), str(rec["tlen"]), rec["seq"] or "*", ql or "*"])


def fastq_sets(records):
    """`samtools fastq -0 -1 -2` over (name, flag, seq, qual-string-or-None) tuples:
    -> {0: [...], 1: [...], 2: [...]} of (header, seq, qual); secondary / supplementary dropped,
    reverse-flagged records reverse-complemented, /1 /2 appended on READ1 / READ2, quality '"'
    (the default -v 1) when the record has none"""
    out = {0: [], 1: [], 2: []}
    for name, flag, seq, qual in records:
        if flag & 0x900:
            continue
        if qual is None:
            qual = '"' * len(seq)
        if flag & 0x10:
            seq, qual = revcomp(seq), qual[::-1]
        r1, r2 = bool(flag & 0x40), bool(flag & 0x80)
        which = 1 if (r1 and not r2) else 2 if (r2 and not r1) else 0
        out[which].append(("@" + name + ("/1" if which == 1 else "/2" if which == 2 else ""), seq, qual))
    return out
