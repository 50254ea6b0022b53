"""Roll-over arithmetic of the reference's output writers, restated (oracle; tests only).

Pure integer/string logic -- the bit-exact part of the contract.  Files are modelled as
ordered lists of keys; nothing is written to disk.

  SequencedOutput  utils/SequencedOutputs.jl:7-71  (check after EVERY key write :37-44,58-63)
  CollatedOutput   utils/Collated.jl:13-64         (K12)
  frame writer     raytracing/RaytracingDriver.jl:87-108 (params/* then p/t,x,k,u[,g] per frame)
  per-frame roller raytracing/TwoLayerRaytracing.jl:95-104,149-157 (older drivers)
"""
from __future__ import annotations


class SequencedOutput:
    """JLD2 flavour: `out[key] = val` counts one write, then check_writes()."""

    def __init__(self, filename_function, max_writes):
        self.max_writes = int(max_writes)
        self.current_writes = 0
        self.file_index = 0
        self.get_filename = filename_function
        self.files = {filename_function(0): []}
        self.current = filename_function(0)

    def _check_writes(self):
        if self.current_writes >= self.max_writes:
            self.current_writes = 0
            self.file_index += 1
            self.current = self.get_filename(self.file_index)
            self.files[self.current] = []

    def __setitem__(self, key, _val):
        self.files[self.current].append(key)
        self.current_writes += 1
        self._check_writes()

    # FourierFlows.Output flavour (:46-56): saveproblem counts 1, saveoutput counts len(fields)
    def saveproblem(self):
        self.files[self.current].append("<problem>")
        self.current_writes += 1
        self._check_writes()

    def saveoutput(self, step, nfields=1):
        self.files[self.current].append(f"snapshots/t/{step}")
        self.current_writes += nfields
        self._check_writes()


def packet_filename(base, idx):
    """raytracing/RaytracingDriver.jl:173-174: @sprintf("%s.%06d.jld2", base, idx)."""
    return "%s.%06d.jld2" % (base, idx)


def savepacketproblem(out):
    """RaytracingDriver.jl:87-94: six params/* keys, in this order."""
    for k in ("f0", "Cg", "dt", "N", "k0", "ωsign"):
        out["params/" + k] = None


def write_packets(out, step, write_gradients):
    """RaytracingDriver.jl:96-108: p/t, p/x, p/k, p/u[, p/g] in this order."""
    for k in ("t", "x", "k", "u") + (("g",) if write_gradients else ()):
        out[f"p/{k}/{step}"] = None


class CollatedOutput:
    """utils/Collated.jl:40-60: name "%s_%08d.out", roll when line_index >= line_limit."""

    def __init__(self, filename, line_limit):
        self.filename = filename
        self.line_limit = int(line_limit)
        self.line_index = 0
        self.file_index = 0
        self.files = {self.get_filename(): []}

    def get_filename(self, idx=None):
        return "%s_%08d.out" % (self.filename, self.file_index if idx is None else idx)

    def write(self, key, _val=None):
        self.files[self.get_filename()].append(key)
        self.line_index += 1
        if self.line_index >= self.line_limit:
            self.line_index = 0
            self.file_index += 1
            self.files[self.get_filename()] = []


class FrameRoller:
    """Older drivers (raytracing/TwoLayerRaytracing.jl:95-104,149-157): counter starts at 1
    (the initial frame), +1 per loop frame, roll to "%s.%08d" idx+1 and reset to 0 when it
    reaches max_writes => every file holds exactly max_writes frames."""

    def __init__(self, base, max_writes):
        self.base, self.max_writes = base, int(max_writes)
        self.file_index, self.current_writes = 0, 1
        self.files = {self.name(): []}

    def name(self):
        return "%s.%08d" % (self.base, self.file_index)

    def initial_frame(self, step):
        self.files[self.name()].append(step)

    def loop_frame(self, step):
        self.files[self.name()].append(step)
        self.current_writes += 1
        if self.current_writes >= self.max_writes:
            self.current_writes = 0
            self.file_index += 1
            self.files[self.name()] = []


def load_packet_analysis_files_collated(open_file, indices, packet_idxs=None, load_velocity=False):
    """analysis/load_file.jl:89-160 restated: reassemble (times, x, k[, u]) from a sequence of packet files whose last frame
    may be split over two files (`open_file(idx)` returns an object with `keys(group)`, `__getitem__`, `__contains__`).

    Quirks kept on purpose (it is the reader the writer has to satisfy): the frame count of a file is the number of `p/t`
    keys in it; the LAST `p/t` key of every file gets its x, k, u from this file or, for what is missing, from the next
    one (:131-148) -- and its time is never stored (`times` stays 0 there, the loop :121 runs over `[1:end-1]`)."""
    import numpy as np
    grp = "p/"
    total_N, Npackets = 0, 0
    for idx in indices:
        f = open_file(idx)
        total_N += len(f.keys(grp + "t"))
        first_key = f.keys(grp + "x")[0]
        Npackets = f[grp + "x/" + first_key].shape[0]
    sel = slice(None) if packet_idxs is None else np.asarray(packet_idxs)
    if packet_idxs is not None:
        Npackets = len(packet_idxs)
    times = np.zeros(total_N)
    x, k, u = (np.zeros((total_N, Npackets, 2)) for _ in range(3))
    base = 0
    for idx in indices:
        f = open_file(idx)
        nxt = open_file(idx + 1) if idx < indices[-1] else None
        tkeys = f.keys(grp + "t")
        N = len(tkeys)
        index = 0
        for ts in tkeys[:-1]:
            times[base + index] = f[grp + f"t/{ts}"]
            x[base + index] = f[grp + f"x/{ts}"][sel]
            k[base + index] = f[grp + f"k/{ts}"][sel]
            if load_velocity:
                u[base + index] = f[grp + f"u/{ts}"][sel]
            index += 1
        ts = tkeys[-1]
        src = {name: (f if (grp + f"{name}/{ts}") in f else nxt) for name in ("x", "k", "u")}
        if (grp + f"u/{ts}") in f:
            pass
        elif (grp + f"k/{ts}") in f:
            src["u"] = nxt
        elif (grp + f"x/{ts}") in f:
            src["k"] = src["u"] = nxt
        else:
            src["x"] = src["k"] = src["u"] = nxt
        x[base + index] = src["x"][grp + f"x/{ts}"][sel]
        k[base + index] = src["k"][grp + f"k/{ts}"][sel]
        u[base + index] = src["u"][grp + f"u/{ts}"][sel]
        base += N
    return (times, x, k, u) if load_velocity else (times, x, k)
