"""Device BGZF/BAM decode (bkid_push_bgzf): host half on the CPU, device half against the host decoder on the GPU."""
import os
import struct
import zlib

import numpy as np
import pytest

from breakid_b200 import api, bamio, synth


def _bgzf_write(path, payload: bytes, block=0xff00, level=6, vary=False):
    rng = np.random.RandomState(9)
    with open(path, "wb") as f:
        o = 0
        while o < len(payload):
            n = int(rng.randint(1, block)) if vary else block
            f.write(bamio._bgzf_block(payload[o:o + n], level if not vary else int(rng.choice([0, 1, 6, 9]))))
            o += n
        f.write(bamio._BGZF_EOF)


def _raw_bam(records, targets):
    text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % t for t in targets)
    out = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(targets))
    for nm, ln in targets:
        out += struct.pack("<i", len(nm) + 1) + nm.encode() + b"\0" + struct.pack("<i", ln)
    first = len(out)
    for r in records:
        out += struct.pack("<i", len(r)) + r
    return out, first


def _record(tid, pos, name, flag, mapq, cigar, l_seq, mtid, mpos, isize, aux=b""):
    nm = name.encode() + b"\0"
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(nm), mapq, 4680, len(cigar), flag, l_seq, mtid, mpos, isize) + nm
    body += b"".join(struct.pack("<I", (n << 4) | op) for n, op in cigar)
    body += b"\x11" * ((l_seq + 1) // 2) + b"\x1e" * l_seq + aux
    return body


def _hostile_records(rng, n):
    """records with every aux type, OC tags, empty SA, long names, unmapped reads, zero-op cigars, big B arrays"""
    recs = []
    pos = 100
    for i in range(n):
        pos += int(rng.randint(0, 50))
        k = i % 12
        name = "read_%d_%s" % (i, "x" * int(rng.randint(0, 60)))
        flag = int(rng.choice([99, 147, 65, 129, 97, 145, 1089, 4, 77, 141, 1, 2]))
        cigar = [(int(rng.randint(1, 90)), 0)]
        aux = b"NMC\x01" + b"MDZ" + b"%d" % int(rng.randint(1, 150)) + b"\0" + b"ASi" + struct.pack("<i", 77)
        if k == 1: aux += b"SAZchr2,%d,+,40M60S,60,0;\0" % (pos + 7)
        if k == 2: aux += b"OCZ60M40S\0" + b"SAZchr1,%d,-,60S40M,13,1;chr2,5,+,10M,0,0;\0" % (pos + 999)
        if k == 3: aux += b"SAZ\0"                                     # empty SA: not an SA record
        if k == 4: aux += b"XBBs" + struct.pack("<i", 5) + b"\1\0\2\0\3\0\4\0\5\0" + b"SAZchrX,9,+,50M50S,60,0;\0"
        if k == 5: aux += b"XFf" + struct.pack("<f", 1.5) + b"XHH1A2B\0" + b"XAAq" + b"XSs" + struct.pack("<h", -3)
        if k == 6: cigar = [(10, 4), (30, 0), (5, 1), (20, 0), (1000, 3), (35, 0), (8, 2), (10, 7), (5, 8), (3, 5)]
        if k == 7: cigar = []
        if k == 8: flag |= 4
        if k == 9: aux = b""
        if k == 10: aux += b"SAZchr1,1,+,1S1M,0,0;\0" + b"SAZignored,2,+,1M,0,0;\0"   # first occurrence wins
        if k == 11: aux += b"XBBI" + struct.pack("<i", 300) + bytes(1200)
        tid = 0 if i < n * 2 // 3 else 1
        if i == n * 2 // 3: pos = 5
        l_seq = int(rng.choice([0, 1, 100, 151]))
        recs.append(_record(tid, pos, name, flag, int(rng.randint(0, 61)), cigar, l_seq, int(rng.choice([-1, 0, 1])), int(rng.randint(-1, 5000)), int(rng.randint(-900, 900)), aux))
    recs.append(_record(-1, -1, "unmapped_tail", 77, 0, [], 30, -1, -1, 0))
    return recs


def _write_hostile(tmp_path, n=4000, vary=True, seed=3):
    rng = np.random.RandomState(seed)
    payload, first = _raw_bam(_hostile_records(rng, n), [("chr1", 500000), ("chr2", 400000)])
    p = str(tmp_path / "hostile.bam")
    _bgzf_write(p, payload, vary=vary)
    return p, payload, first


def test_bgzf_open_block_table_and_header(tmp_path):
    p, payload, first = _write_hostile(tmp_path)
    f = api.BgzfFile(p)
    assert f.target_names == ["chr1", "chr2"] and f.target_len == [500000, 400000]
    assert f.first_record == first and f.usize == len(payload)
    bt = f.block_table()
    assert int(bt["usize"].sum()) == len(payload) and f.n_blocks == len(bt)
    raw = open(p, "rb").read()
    got = b"".join(zlib.decompress(raw[int(b["payload_off"]):int(b["payload_off"]) + int(b["payload_len"])], -15) for b in bt)
    assert got == payload
    f.close()
    with pytest.raises(IOError):
        api.BgzfFile(str(tmp_path / "missing.bam"))
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(b"not a bam file at all, just text" * 10)
    with pytest.raises(IOError):
        api.BgzfFile(bad)


_ALL = (("flag", np.uint16), ("mapq", np.uint8), ("tid", np.int32), ("pos", np.int32), ("isize", np.int32), ("endpos", np.int32)) + api._XCOLS + api._SIDE


def _assert_same_as_host_decoder(path, chunk_kb=None):
    hb = api.HostBatch.from_bam(path, threads=4)
    f = api.BgzfFile(path)
    if chunk_kb:
        os.environ["BKID_BGZF_CHUNK_KB"] = str(chunk_kb)
    try:
        c = api.Context(f.target_len, f.target_names, device=0)
        n = c.push_bgzf(f)
    finally:
        os.environ.pop("BKID_BGZF_CHUNK_KB", None)
    assert n == hb.n
    st = c.decode_stats()
    assert st["n_records"] == hb.n and st["uncompressed_bytes"] == f.usize
    for k, dt in _ALL:
        got = c.fetch_column(k, dt)
        exp = hb.cols[k] if k in hb.cols else hb.x[k] if k in hb.x else hb.side[k]
        assert np.array_equal(got, np.asarray(exp, dtype=dt).reshape(-1)), (k, chunk_kb)
    if hb.seq is not None:                       # read bases of the SA records
        for k, dt in api._SEQ:
            assert np.array_equal(c.fetch_column(k, dt), hb.seq[k]), k
    f.close()
    return c, hb, st


@pytest.mark.gpu
@pytest.mark.parametrize("chunk_kb", [None, 64, 300, 1024])
def test_device_decode_equals_host_decoder_hostile(tmp_path, chunk_kb):
    p, _, _ = _write_hostile(tmp_path, n=6000)
    c, hb, st = _assert_same_as_host_decoder(p, chunk_kb)
    if chunk_kb:
        assert st["n_chunks"] > 1
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunk_kb", [None, 2048])
def test_device_decode_whole_path(tmp_path, small_data, chunk_kb):
    """BAM file -> device decode -> whole hot path == oracle on the generator's own batch"""
    import oracle_py as O
    d, hb0, nibs = small_data
    p = str(tmp_path / "reads.bam")
    genome = [synth.nib_ascii(pl, l) for pl, l in nibs]
    seq, reads = synth.split_read_sequences(hb0, genome)
    bamio.write_bam(p, d, sa_seq=seq)
    c, hb, st = _assert_same_as_host_decoder(p, chunk_kb)
    assert hb.seq is not None and np.array_equal(hb.seq["seq4"], seq["seq4"]) and np.array_equal(hb.seq["seq_len"], seq["seq_len"])
    for t, (pl, l) in enumerate(nibs):
        c.set_nib(t, pl, l)
    mean, sd, dist, ncall = c.run()
    got = c.fetch_clusters()
    om, osd, od, exp = O.run(hb0, nibs, mode=0)
    assert (mean, sd, dist) == (om, osd, od)
    assert got.tobytes() == exp.tobytes()
    # a second file pushed after a reset reuses the decoder state
    c.reset()
    f = api.BgzfFile(p)
    assert c.push_bgzf(f) == hb.n
    assert c.run()[:3] == (om, osd, od)
    f.close()
    c.close()


@pytest.mark.gpu
def test_device_decode_rejects_corrupt_files(tmp_path):
    p, payload, first = _write_hostile(tmp_path, n=3000, vary=False)
    raw = bytearray(open(p, "rb").read())
    f0 = api.BgzfFile(p)
    bt = f0.block_table()
    f0.close()
    # damaged deflate payload (reserved block type in the first byte of block 3)
    bad = bytearray(raw)
    bad[int(bt["payload_off"][3])] |= 0x06
    q = str(tmp_path / "bad_deflate.bam")
    open(q, "wb").write(bad)
    f = api.BgzfFile(q)
    c = api.Context(f.target_len, f.target_names, device=0)
    with pytest.raises(api.BkidError, match="inflate failed"):
        c.push_bgzf(f)
    f.close()
    # a flipped bit in the stored CRC32, and a flipped literal that still inflates to the right size: both are CRC errors
    bad = bytearray(raw)
    bad[int(bt["payload_off"][5] + bt["payload_len"][5])] ^= 0x10
    q = str(tmp_path / "bad_crc.bam")
    open(q, "wb").write(bad)
    f = api.BgzfFile(q)
    with pytest.raises(api.BkidError, match="CRC32 mismatch in BGZF block 5"):
        c.push_bgzf(f)
    f.close()
    hits = 0
    for off in range(40, 400, 7):                      # somewhere in the payload of block 7: most flips change literals only
        bad = bytearray(raw)
        bad[int(bt["payload_off"][7]) + off] ^= 0x01
        open(q, "wb").write(bad)
        f = api.BgzfFile(q)
        try:
            c.reset()
            c.push_bgzf(f)
        except api.BkidError as e:
            hits += 1
            assert "BGZF block 7" in str(e), str(e)
        f.close()
    assert hits == len(range(40, 400, 7))               # no damaged block gets through
    # truncated last record: drop the tail of the uncompressed stream
    q = str(tmp_path / "trunc.bam")
    _bgzf_write(q, payload[:-10])
    f = api.BgzfFile(q)
    with pytest.raises(api.BkidError, match="truncated BAM record"):
        c.push_bgzf(f)
    f.close()
    # the context is still usable afterwards
    c.reset()
    f = api.BgzfFile(p)
    assert c.push_bgzf(f) > 3000
    f.close()
    c.close()


@pytest.mark.gpu
def test_device_decode_edge_files(tmp_path):
    """no records at all; a BAM header that spans several BGZF blocks (first record deep inside block 3); a single
    record; every record in its own BGZF block"""
    rng = np.random.RandomState(4)
    targets = [("chr1", 500000), ("chr2", 400000)]
    # 1. header only
    payload, first = _raw_bam([], targets)
    p = str(tmp_path / "empty.bam")
    _bgzf_write(p, payload)
    f = api.BgzfFile(p)
    c = api.Context(f.target_len, f.target_names, device=0)
    assert c.push_bgzf(f) == 0
    f.close()
    c.close()
    # 2. long header: 200 KB of @CO lines in front of the records
    recs = _hostile_records(rng, 500)
    text_pad = "".join("@CO\t%s\n" % ("x" * 90) for _ in range(2200))
    text = "@HD\tVN:1.4\tSO:coordinate\n" + text_pad + "".join("@SQ\tSN:%s\tLN:%d\n" % t for t in targets)
    out = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(targets))
    for nm, ln in targets:
        out += struct.pack("<i", len(nm) + 1) + nm.encode() + b"\0" + struct.pack("<i", ln)
    assert len(out) > 3 * 0xff00
    for r in recs:
        out += struct.pack("<i", len(r)) + r
    p = str(tmp_path / "longhdr.bam")
    _bgzf_write(p, out)
    cc, hb, st = _assert_same_as_host_decoder(p, None)
    cc.close()
    cc, hb, st = _assert_same_as_host_decoder(p, 64)
    cc.close()
    # 3. one record; 4. one BGZF block per record (tiny blocks)
    payload, first = _raw_bam(recs[:1], targets)
    p = str(tmp_path / "one.bam")
    _bgzf_write(p, payload)
    cc, hb, st = _assert_same_as_host_decoder(p, None)
    assert hb.n == 1
    cc.close()
    payload, first = _raw_bam(recs[:300], targets)
    p = str(tmp_path / "tiny_blocks.bam")
    with open(p, "wb") as fo:
        fo.write(bamio._bgzf_block(payload[:first], 6))
        o = first
        for r in recs[:300]:
            n = 4 + len(r)
            fo.write(bamio._bgzf_block(payload[o:o + n], 6))
            o += n
        fo.write(bamio._BGZF_EOF)
    cc, hb, st = _assert_same_as_host_decoder(p, None)
    assert hb.n == 300 and st["n_blocks"] == 301
    cc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nparts,chunk_kb", [(2, None), (5, None), (3, 256), (16, None)])
def test_block_range_decode_union_equals_whole_file(tmp_path, nparts, chunk_kb):
    """multi-GPU ingest: every rank decodes a block range of the file; the ranges stitch exactly (where one range lands
    is where the next one starts) and their union is the whole-file decode, column by column"""
    p, payload, first = _write_hostile(tmp_path, n=9000)
    f = api.BgzfFile(p)
    whole = api.Context(f.target_len, f.target_names, device=0)
    n_all = whole.push_bgzf(f)
    cuts = [f.n_blocks * i // nparts for i in range(nparts + 1)]
    if chunk_kb:
        os.environ["BKID_BGZF_CHUNK_KB"] = str(chunk_kb)
    try:
        parts, marks = [], []
        for r in range(nparts):
            c = api.Context(f.target_len, f.target_names, device=0)
            n, a, b = c.push_bgzf_range(f, cuts[r], cuts[r + 1])
            parts.append((c, n)); marks.append((a, b))
    finally:
        os.environ.pop("BKID_BGZF_CHUNK_KB", None)
    assert marks[0][0] == first and marks[-1][1] == f.usize
    for r in range(nparts - 1):
        assert marks[r][1] == marks[r + 1][0], (r, marks)             # the cross-rank stitch check
    assert sum(n for _, n in parts) == n_all
    base = 0
    for k, dt in (("flag", np.uint16), ("mapq", np.uint8), ("tid", np.int32), ("pos", np.int32), ("isize", np.int32), ("endpos", np.int32)):
        assert np.array_equal(np.concatenate([c.fetch_column(k, dt) for c, _ in parts]), whole.fetch_column(k, dt)), k
    # sparse / SA tables: record indices are range-local, everything else concatenates
    off = np.cumsum([0] + [n for _, n in parts])
    assert np.array_equal(np.concatenate([c.fetch_column("x_rec", np.uint32).astype(np.int64) + off[i] for i, (c, _) in enumerate(parts)]),
                          whole.fetch_column("x_rec", np.uint32).astype(np.int64))
    for k, dt in (("x_mtid", np.int32), ("x_mpos", np.int32), ("x_name_hash", np.uint64), ("cig_ops", np.uint32), ("sa_txt", np.uint8), ("oc_txt", np.uint8), ("seq4", np.uint8), ("seq_len", np.int32)):
        assert np.array_equal(np.concatenate([c.fetch_column(k, dt) for c, _ in parts]), whole.fetch_column(k, dt)), k
    for c, _ in parts:
        c.close()
    whole.close()
    f.close()


@pytest.mark.gpu
def test_false_seed_is_repaired_by_the_stitch(tmp_path):
    """a 600 KB aux array that spans segment boundaries and carries, right behind each boundary, three consecutive
    perfectly plausible fake records: the seed search must take the bait (they are the first plausible chain of the
    segment) and the stitch must throw it out again because the true chain does not land there"""
    rng = np.random.RandomState(12)
    targets = [("chr1", 500000), ("chr2", 400000)]
    recs = _hostile_records(rng, 800)
    payload0, first = _raw_bam(recs[:400], targets)
    big_off = len(payload0)                                   # stream offset where the big record starts
    name = "bigrec"
    head_len = 4 + 32 + len(name) + 1 + 4 + 0 + 0             # block_size + core + name + one cigar op, l_seq = 0
    aux_hdr = b"XBBc" + struct.pack("<i", 600000)
    arr0 = big_off + head_len + len(aux_hdr)                  # stream offset of the first array byte
    arr = bytearray(600000)
    fake = b"".join(struct.pack("<i", len(r)) + r for r in
                    [_record(0, 1000 + 7 * k, "fake%d" % k, 99, 60, [(100, 0)], 100, 0, 1200, 300, b"NMC\x01") for k in range(3)])
    SEG = 256 << 10
    planted = 0
    for b in range((arr0 // SEG + 1) * SEG, arr0 + len(arr) - len(fake) - 16, SEG):
        o = b - arr0 + 3                                      # 3 bytes behind the segment boundary
        arr[o:o + len(fake)] = fake
        planted += 1
    assert planted >= 2
    big = _record(0, 60000, name, 99, 60, [(50, 0)], 0, 0, 61000, 300, aux_hdr + bytes(arr))
    payload, _ = _raw_bam(recs[:400] + [big] + recs[400:], targets)
    # recs[400:] restart at low positions on chr1/chr2: coordinate order does not matter to the decoder
    p = str(tmp_path / "bait.bam")
    _bgzf_write(p, payload, vary=True)
    c, hb, st = _assert_same_as_host_decoder(p, None)
    assert hb.n == 802 and st["seed_repairs"] >= planted
    c.close()
    c, hb, st = _assert_same_as_host_decoder(p, 512)          # several chunks: the bait also sits in carried data
    c.close()


@pytest.mark.gpu
def test_false_seed_inside_record_straddling_range_end(tmp_path):
    """block-range ingest: the record that straddles the END of a range carries plausible fake record chains, so the
    straddling record jumps over false-positive seeds that lie before the range end.  Such a seed must be dropped
    (it is not on the chain) although the walk that proves it already landed behind the range end."""
    rng = np.random.RandomState(21)
    targets = [("chr1", 500000), ("chr2", 400000)]
    recs = _hostile_records(rng, 800)
    payload0, first = _raw_bam(recs[:400], targets)
    big_off = len(payload0)
    arr = bytearray(900000)
    fake = b"".join(struct.pack("<i", len(r)) + r for r in
                    [_record(0, 1000 + 7 * k, "fake%d" % k, 99, 60, [(100, 0)], 100, 0, 1200, 300, b"NMC\x01") for k in range(3)])
    for o in range(64, len(arr) - len(fake) - 64, 4096):      # bait in every segment, whatever the segment grid of a range is
        arr[o:o + len(fake)] = fake
    big = _record(0, 60000, "bigrec", 99, 60, [(50, 0)], 0, 0, 61000, 300, b"XBBc" + struct.pack("<i", len(arr)) + bytes(arr))
    payload, _ = _raw_bam(recs[:400] + [big] + recs[400:], targets)
    p = str(tmp_path / "bait_range.bam")
    _bgzf_write(p, payload, vary=True)
    f = api.BgzfFile(p)
    whole = api.Context(f.target_len, f.target_names, device=0)
    n_all = whole.push_bgzf(f)
    assert n_all == 802
    ustart = np.concatenate([[0], np.cumsum(f.block_table()["usize"].astype(np.int64))])
    big_end = big_off + 4 + len(big)
    for frac in (0.35, 0.6, 0.9):                              # range ends inside the big record, behind 1..3 segments of bait
        cut = int(np.searchsorted(ustart, big_off + int(frac * len(arr)), side="right") - 1)
        assert ustart[cut] > big_off + (256 << 10) and ustart[cut] < big_off + len(arr)
        # (the range that would START at `cut` begins inside the bait: a mid-stream start is verified by the cross-range
        # check next_record_uoff[r] == first_record_uoff[r+1], it cannot be repaired -- only the first range is decoded here)
        c = api.Context(f.target_len, f.target_names, device=0)
        n, lo, hi = c.push_bgzf_range(f, 0, cut)
        assert (n, lo, hi) == (401, first, big_end), (frac, n, lo, hi)   # 400 records + the straddling one, landing right behind it
        for k, dt in (("flag", np.uint16), ("pos", np.int32), ("isize", np.int32), ("endpos", np.int32)):
            assert np.array_equal(c.fetch_column(k, dt), whole.fetch_column(k, dt)[:401]), (frac, k)
        c.close()
    whole.close()
    f.close()


@pytest.mark.gpu
def test_device_decode_rejects_records_whose_fields_do_not_fit(tmp_path):
    """bam_read1 (htslib sam.c:427-429) rejects l_qseq < 0, l_qname < 1 and fixed fields longer than the record; so does
    the device decoder (BKID_ERR_IO) instead of indexing past the record.  An SA:Z value without NUL is corrupt too."""
    targets = [("chr1", 500000)]
    good = [_record(0, 100 + 10 * i, "r%d" % i, 99, 60, [(50, 0)], 50, 0, 400, 300, b"NMC\x01") for i in range(50)]

    def write(bad, name):
        payload, _ = _raw_bam(good[:25] + [bad] + good[25:], targets)
        p = str(tmp_path / name)
        _bgzf_write(p, payload, vary=False)
        return p

    r = bytearray(_record(0, 350, "liar", 99, 60, [(50, 0)], 50, 0, 400, 300, b"NMC\x01"))
    cases = []
    # body layout (no block_size prefix): refID 0, pos 4, l_read_name 8, mapq 9, bin 10, n_cigar_op 12, flag 14, l_seq 16
    b1 = bytearray(r); b1[12:14] = struct.pack("<H", 60000)                                            # 60000 cigar ops in a ~120-byte record
    cases.append(("ncig.bam", bytes(b1)))
    b2 = bytearray(r); b2[16:20] = struct.pack("<i", 1 << 20)                                          # l_seq = 1 Mi in a ~120-byte record
    cases.append(("lseq.bam", bytes(b2)))
    b3 = bytearray(r); b3[16:20] = struct.pack("<i", -5)
    cases.append(("lseq_neg.bam", bytes(b3)))
    b0 = bytearray(r); b0[8] = 0                                                                       # l_read_name = 0
    cases.append(("lname0.bam", bytes(b0)))
    b4 = _record(0, 350, "unterminated", 99, 60, [(50, 0)], 50, 0, 400, 300, b"NMC\x01SAZchr1,5,+,25M25S,60,0;")   # no NUL
    cases.append(("sa_nonul.bam", b4))
    for name, bad in cases:
        p = write(bad, name)
        f = api.BgzfFile(p)
        c = api.Context(f.target_len, f.target_names, device=0)
        with pytest.raises(api.BkidError) as e:
            c.push_bgzf(f)
        assert "corrupt BAM record" in str(e.value), (name, str(e.value))
        c.close(); f.close()
