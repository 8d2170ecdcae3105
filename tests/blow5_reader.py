"""Independent SLOW5/BLOW5 reader (test infrastructure): parses a file from the published SLOW5 v0.2.0 layout with
``struct`` only, sharing no code with csrc/blow5_writer.cpp."""
import struct
import zlib


def svb_zd_decode(blob):
    """slow5lib's "svb-zd" signal compression, decoded independently of csrc/blow5_writer.cpp: uint32 count, StreamVByte
    control bytes (2 bits per value = bytes - 1, value i of a group of four in bits 2 (i % 4)), little-endian data
    bytes; zigzag + delta (first delta against 0) back to int16."""
    (n,) = struct.unpack_from("<I", blob, 0)
    n_ctrl = (n + 3) // 4
    ctrl, pos = blob[4:4 + n_ctrl], 4 + n_ctrl
    out, prev = [], 0
    for i in range(n):
        nb = ((ctrl[i >> 2] >> (2 * (i & 3))) & 3) + 1
        z = int.from_bytes(blob[pos:pos + nb], "little")
        pos += nb
        d = (z >> 1) ^ -(z & 1)
        prev += d
        assert -32768 <= prev <= 32767
        out.append(prev)
    assert pos == len(blob), "svb-zd blob size"
    return out


def read_blow5(path):
    data = open(path, "rb").read()
    assert data[:6] == b"BLOW5\x01", "magic"
    version = tuple(data[6:9])
    rec_comp = data[9]
    (n_groups,) = struct.unpack_from("<I", data, 10)
    sig_comp = data[14]
    assert data[15:64] == b"\0" * 49, "padding"
    (hsize,) = struct.unpack_from("<I", data, 64)
    ascii_hdr = data[68:68 + hsize].decode()
    lines = ascii_hdr.split("\n")
    assert lines[-1] == ""
    attrs = dict(l[1:].split("\t", 1) for l in lines if l.startswith("@"))
    types = [l for l in lines if l.startswith("#")][0][1:].split("\t")
    names = [l for l in lines if l.startswith("#")][1][1:].split("\t")
    assert len(types) == len(names)
    pos, records = 68 + hsize, []
    assert data[-5:] == b"5WOLB", "eof marker"
    while pos < len(data) - 5:
        (size,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        body = data[pos:pos + size]
        pos += size
        if rec_comp == 1:
            body = zlib.decompress(body)
        q = 0
        (idl,) = struct.unpack_from("<H", body, q); q += 2
        rid = body[q:q + idl].decode(); q += idl
        (group,) = struct.unpack_from("<I", body, q); q += 4
        dig, off, rng, rate = struct.unpack_from("<dddd", body, q); q += 32
        (n,) = struct.unpack_from("<Q", body, q); q += 8
        if sig_comp == 1:
            (cb,) = struct.unpack_from("<Q", body, q); q += 8
            sig = svb_zd_decode(body[q:q + cb]); q += cb
            assert len(sig) == n
        else:
            assert sig_comp == 0
            sig = struct.unpack_from(f"<{n}h", body, q); q += 2 * n
        (cl,) = struct.unpack_from("<Q", body, q); q += 8
        chan = body[q:q + cl].decode(); q += cl
        (med,) = struct.unpack_from("<d", body, q); q += 8
        (rnum,) = struct.unpack_from("<i", body, q); q += 4
        (mux,) = struct.unpack_from("<B", body, q); q += 1
        (stime,) = struct.unpack_from("<Q", body, q); q += 8
        assert q == len(body), "record size"
        records.append(dict(read_id=rid, read_group=group, digitisation=dig, offset=off, range=rng, sampling_rate=rate,
                            len_raw_signal=n, signal=list(sig), channel_number=chan, median_before=med,
                            read_number=rnum, start_mux=mux, start_time=stime))
    assert pos == len(data) - 5
    return dict(version=version, record_compression=rec_comp, signal_compression=sig_comp, num_read_groups=n_groups,
                attrs=attrs, types=types, names=names, records=records)


def record_span(path):
    """(first byte of the first record, first byte of the end marker) of a BLOW5 file."""
    data = open(path, "rb").read()
    assert data[:6] == b"BLOW5\x01" and data[-5:] == b"5WOLB"
    (hsize,) = struct.unpack_from("<I", data, 64)
    return 68 + hsize, len(data) - 5


def read_slow5(path):
    lines = open(path).read().split("\n")
    assert lines[0] == "#slow5_version\t0.2.0" and lines[1] == "#num_read_groups\t1"
    attrs = dict(l[1:].split("\t", 1) for l in lines if l.startswith("@"))
    hdr = [l for l in lines[2:] if l.startswith("#")]
    names = hdr[1][1:].split("\t")
    recs = []
    for l in lines:
        if not l or l[0] in "#@":
            continue
        f = dict(zip(names, l.split("\t")))
        f["signal"] = [int(x) for x in f.pop("raw_signal").split(",")]
        recs.append(f)
    return dict(attrs=attrs, names=names, records=recs)
