"""Shared helpers for the tests."""
import ctypes
import math

import numpy as np

from rtc_b200._types import OBJECT_DTYPE, RtcParams


def params_from_bytes(b):
    p = RtcParams()
    raw = np.asarray(b, np.uint8).tobytes()
    ctypes.memmove(ctypes.byref(p), raw, ctypes.sizeof(p))
    return p


def objs_from_bytes(b):
    return np.frombuffer(np.asarray(b, np.uint8).tobytes(), OBJECT_DTYPE).copy()


PI32 = np.float32(math.pi)


def parse_stream(stream, x, y, mode):
    """Independent decoder of a minimised stream -> (colour keys, glyphs, full flags).
    Used for size-independent property checks (decode o encode == identity)."""
    cs = 12 if mode in (0, 1) else 20
    W = x - 1
    s = np.asarray(stream, np.uint8)
    keys = np.zeros((y * W, 3 if cs == 20 else 1), np.uint8)
    glyphs = np.zeros(y * W, np.uint8)
    full = np.zeros(y * W, np.uint8)
    i = 0
    cur = None
    for cell in range(y * W):
        if s[i] == 0x1B and i + cs <= len(s) and s[i + 1] == ord("[") and s[i + 3] == ord("8"):
            body = s[i:i + cs]

            def dec(b3):
                v = 0
                for ch in b3:
                    v = v * 10 + (0 if ch == 0 else int(ch) - 48)
                return v
            if cs == 20:
                cur = (dec(body[7:10]), dec(body[11:14]), dec(body[15:18]))
            else:
                cur = (dec(body[7:10]),)
            glyphs[cell] = body[cs - 1]
            full[cell] = 1
            i += cs
        else:
            glyphs[cell] = s[i]
            i += 1
        keys[cell] = cur
        if cell % W == W - 1:
            assert s[i] == 10, "missing newline at end of row"
            i += 1
    assert i == len(s), "trailing bytes in stream"
    return keys.reshape(-1), glyphs, full


def bind_stream(ctx):
    """Run torch and the rtc context on ONE explicit stream.  (rtc_set_stream(NULL) means "the context's own non-blocking
    stream", never the legacy default stream -- handing it torch's default stream would leave the rtc kernels unordered
    against the torch kernels that produce their inputs.)"""
    import torch
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx.set_stream(s.cuda_stream)
    return s


def unbind_stream(ctx):
    import torch
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    ctx.set_stream(0)
