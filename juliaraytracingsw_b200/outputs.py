"""Roll-over arithmetic of utils/SequencedOutputs.jl and utils/Collated.jl through the C ABI
(integer logic only; the Julia side keeps doing the JLD2 writes)."""
from __future__ import annotations

import ctypes as C

from ._lib import SeqOut, check, lib


class SequencedOutput:
    """SequencedOutput(filename_function, max_writes): `out[key] = val` counts one write and the
    roll-over check runs after EVERY key (utils/SequencedOutputs.jl:37-44,58-63)."""

    def __init__(self, base_filename, max_writes):
        self.base = base_filename
        self._s = SeqOut()
        check(lib().swrt_seqout_init(C.byref(self._s), int(max_writes)))
        self.files = {}

    def filename(self, idx):
        buf = C.create_string_buffer(1024)
        check(lib().swrt_seqout_filename(self.base.encode(), idx, buf, 1024))
        return buf.value.decode()

    def __setitem__(self, key, val):
        idx = C.c_longlong()
        check(lib().swrt_seqout_write(C.byref(self._s), 1, C.byref(idx)))
        self.files.setdefault(self.filename(idx.value), []).append(key)

    @property
    def file_index(self):
        return self._s.file_index

    @property
    def current_writes(self):
        return self._s.current_writes


def collated_filename(base, idx):
    buf = C.create_string_buffer(1024)
    check(lib().swrt_collated_filename(base.encode(), idx, buf, 1024))
    return buf.value.decode()
