"""Reader for the "ADMP1" dump files written by oracle/ref_harness.cpp (test infrastructure)."""
import struct

import numpy as np


def read_dump(path):
    """Return {name: ndarray} for every record in an ADMP1 file."""
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:5] == b"ADMP1", "not an ADMP1 dump"
    pos = 5
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl
        dtype = chr(buf[pos])
        pos += 1
        (nd,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        dims = struct.unpack_from("<%dq" % nd, buf, pos)
        pos += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        if dtype == "d":
            arr = np.frombuffer(buf, dtype="<f8", count=n, offset=pos).reshape(dims)
            pos += 8 * n
        else:
            arr = np.frombuffer(buf, dtype="<i4", count=n, offset=pos).reshape(dims)
            pos += 4 * n
        out[name] = arr
    return out
