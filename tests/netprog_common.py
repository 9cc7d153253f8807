"""Pure-Python reader of the network program format (csrc/net.cu header comment) -- test infrastructure."""
import struct


def parse_program(blob):
    assert blob[:8] == b"DANETPRG"
    (version, B, C, H, W, n_buf, n_const, n_out, n_step, prec, _r0, _r1) = struct.unpack_from("<12I", blob, 8)
    steps_off, steps_bytes, payload_off, payload_bytes = struct.unpack_from("<4Q", blob, 56)
    off = 88
    bufs = list(struct.unpack_from("<%dQ" % n_buf, blob, off)); off += 8 * n_buf
    consts = [struct.unpack_from("<QQ", blob, off + 16 * k) for k in range(n_const)]; off += 16 * n_const
    outs = []
    for k in range(n_out):
        name = blob[off:off + 32].split(b"\0")[0].decode()
        kind, rid, roff = struct.unpack_from("<IIQ", blob, off + 32)
        eb, ndim, d0, d1, d2, d3 = struct.unpack_from("<Ii4i", blob, off + 48)
        outs.append(dict(name=name, ref=(kind, rid, roff), elem_bytes=eb, dims=[d0, d1, d2, d3][:ndim]))
        off += 72
    assert off <= steps_off
    steps = []
    p = steps_off
    for _ in range(n_step):
        op, ni, nf, nr = struct.unpack_from("<4I", blob, p); p += 16
        ints = list(struct.unpack_from("<%di" % ni, blob, p)); p += 4 * ni
        floats = list(struct.unpack_from("<%df" % nf, blob, p)); p += 4 * nf
        refs = [struct.unpack_from("<IIQ", blob, p + 16 * r) for r in range(nr)]; p += 16 * nr
        steps.append(dict(op=op, ints=ints, floats=floats, refs=refs))
    assert p == steps_off + steps_bytes
    assert payload_off + payload_bytes <= len(blob)
    return dict(version=version, batch=B, chw=(C, H, W), bufs=bufs, consts=consts, outs=outs, steps=steps, precision=prec,
                payload=(payload_off, payload_bytes))
