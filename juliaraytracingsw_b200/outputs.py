"""Output side of the drivers: roll-over arithmetic of utils/SequencedOutputs.jl and utils/Collated.jl through the C ABI
(integer logic only) plus a writer that actually stores the frames under the reference's keys.

The reference writes JLD2 (HDF5) files; neither JLD2 nor h5py exists in this image, so the Python host stores each output file
as an ordered key -> array archive (`KeyedFile`, one `.npz` per file with the key order kept) -- same file names, same keys,
same partition of keys over files (a frame can straddle two files: the check runs after EVERY key,
utils/SequencedOutputs.jl:37-44,58-63).  A Julia host keeps using JLD2 with `swrt_seqout_*` for the arithmetic.  The
reader that accepts these files is the oracle's restatement of analysis/load_file.jl:89-160 (tests only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import SeqOut, check, lib


class KeyedFile:
    """One output file: `f[key] = value` in write order; `keys(group)` lists the members of a group like JLD2's
    `keys(file["p/t"])`; `close()` stores it as `<name>.npz` when a directory was given."""

    def __init__(self, name, directory=None):
        self.name, self.directory = name, directory
        self.data = {}

    def __setitem__(self, key, val):
        if key in self.data:
            raise KeyError(f"{key} already written to {self.name}")        # JLD2 refuses to overwrite a dataset too
        self.data[key] = None if val is None else np.array(val)

    def __getitem__(self, key):
        return self.data[key]

    def __contains__(self, key):
        return key in self.data

    def keys(self, group):
        pre = group.rstrip("/") + "/"
        return [k[len(pre):] for k in self.data if k.startswith(pre)]

    def close(self):
        if self.directory is not None:
            os.makedirs(self.directory, exist_ok=True)
            order = np.array(list(self.data), dtype=object)
            arrays = {f"a{i}": (np.array(np.nan) if v is None else v) for i, v in enumerate(self.data.values())}
            np.savez(os.path.join(self.directory, os.path.basename(self.name) + ".npz"), __order__=order, **arrays)

    @classmethod
    def load(cls, path):
        z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=True)
        f = cls(os.path.basename(path))
        for i, k in enumerate(z["__order__"]):
            f.data[str(k)] = z[f"a{i}"]
        return f


class SequencedOutput:
    """SequencedOutput(filename_function, max_writes): `out[key] = val` counts one write and the
    roll-over check runs after EVERY key (utils/SequencedOutputs.jl:37-44,58-63)."""

    def __init__(self, base_filename, max_writes, store=False, directory=None):
        """store=True keeps the values (KeyedFile per output file; written to `directory` as .npz when given) -- the writer;
        store=False only records which key went to which file."""
        self.base = base_filename
        self._s = SeqOut()
        check(lib().swrt_seqout_init(C.byref(self._s), int(max_writes)))
        self.files = {}
        self.store, self.directory = bool(store), directory
        self.archives = {}

    def filename(self, idx):
        buf = C.create_string_buffer(1024)
        check(lib().swrt_seqout_filename(self.base.encode(), idx, buf, 1024))
        return buf.value.decode()

    def __setitem__(self, key, val):
        idx = C.c_longlong()
        before = self._s.file_index
        check(lib().swrt_seqout_write(C.byref(self._s), 1, C.byref(idx)))
        name = self.filename(idx.value)
        self.files.setdefault(name, []).append(key)
        if self.store:
            if name not in self.archives:
                self.archives[name] = KeyedFile(name, self.directory)
            self.archives[name][key] = val
            if self._s.file_index != before:                        # check_writes(): close the full file, open the next
                self.archives[name].close()

    def close(self):
        """close(output): flush the file that is still open."""
        if self.store:
            name = self.filename(self._s.file_index)
            if name in self.archives:
                self.archives[name].close()

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
